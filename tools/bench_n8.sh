p=29520
for w in sha256 sha256_2p30 gemm; do
  p=$((p+1)); st=200
  [ $w = sha256_2p30 ] && st=10
  [ $w = gemm ] && st=50
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 8 --workload $w --steps $st --warmup 5 2>gpurun_out/n8_$w.err | tail -1 > gpurun_out/bench_r01_${w}_n8.json
  python tools/show_bench.py gpurun_out/bench_r01_${w}_n8.json gpurun_out/n8_$w.err
done
