#!/bin/bash
# r02 GPU call 2b: AES with lane-distributed Philox, GEMM with shared-memory operands again, hybrid / zero-copy host paths, compact ncu summaries
set -u
out=gpurun_out/r02c2b
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py tests/test_chstone_aes.py tests/test_board_b200_flow.py -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -6 "$out/pytest_gpu.log"
{
for nc in 1 2 3; do
  python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time
  python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --inject 0.0009765625
done
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1 --inject 0.0009765625
for nc in 1 2 3; do python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 20 --time; done
} > "$out/timings.txt" 2>&1
cat "$out/timings.txt"
for hp in hybrid zerocopy; do
  for wl in sha256 aes crc16; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --workload $wl --host-path $hp > "$out/bench_${wl}_${hp}.json" 2> "$out/bench_${wl}_${hp}.err"; echo "bench $wl $hp rc=$?" | tee -a "$out/summary.txt"
  done
done
prof() {  # name, kernel regex, profile_target args...
  local name=$1 k=$2; shift 2
  timeout 600 ncu --set full --clock-control none -k regex:$k -c 1 -o "$out/$name" python tools/profile_target.py "$@" > "$out/ncu_$name.log" 2>&1
  echo "ncu $name rc=$?" | tee -a "$out/summary.txt"
  python tools/ncu_summary.py "$out/$name.ncu-rep" "$out/$name.json" > /dev/null 2>&1
  rm -f "$out/$name.ncu-rep"
}
prof aes_enc_nc2_inj1 xmr_aes128_enc_nc2_inj1 --kernel aes --nc 2 --log2n 24 --iters 2 --inject 0.0009765625
prof aes_enc_nc2_inj0 xmr_aes128_enc_nc2_inj0 --kernel aes --nc 2 --log2n 24 --iters 2
prof aes_dec_nc2_inj0 xmr_aes128_dec_nc2_inj0 --kernel aes --nc 2 --log2n 24 --iters 2 --aes-mode 1
prof gemm_nc3 xmr_gemm_tf32_nc3_inj0 --kernel gemm --nc 3 --side 4096 --iters 2
prof gemm_nc1 xmr_gemm_tf32_nc1_inj0 --kernel gemm --nc 1 --side 4096 --iters 2
{
for cfg in "16 0" "32 0" "16 1" "32 1"; do
  set -- $cfg
  echo "== GROUP_M=$1 L2_HINTS=$2"
  COAST_GEMM_GROUP_M=$1 COAST_GEMM_L2_HINTS=$2 python tools/profile_target.py --kernel gemm --nc 3 --side 4096 --iters 20 --time
  COAST_GEMM_GROUP_M=$1 COAST_GEMM_L2_HINTS=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:xmr_gemm_tf32_nc3_inj0 -c 1 --csv python tools/profile_target.py --kernel gemm --nc 3 --side 4096 --iters 2 2>/dev/null | grep -E "dram__bytes|gpu__time" | cut -d, -f10- 
done
} > "$out/gemm_l2_sweep.txt" 2>&1
cat "$out/gemm_l2_sweep.txt"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c2b/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsed',e); continue
    e=d.get('e2e',{})
    print(f, d.get('value'), d.get('ms_per_step'), 'e2e', e.get('value'), e.get('ms_per_step'), e.get('path'), e.get('frac_of_bound'), e.get('pcie_pinned_copy_gbs'))
for f in sorted(glob.glob('gpurun_out/r02c2b/*.json')):
    if 'bench_' in f: continue
    try:
        d=json.load(open(f))[0]
        print(f, {k:d.get(k) for k in ('kernel','duration','dram_read','dram_write','pipe_alu_pct','pipe_lsu_pct','pipe_fma_pct','pipe_tensor_pct','warp_insts','issue_per_cycle_per_smsp','registers')})
    except Exception as e: print(f, e)
PY
du -sh gpurun_out
