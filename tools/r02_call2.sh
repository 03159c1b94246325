#!/bin/bash
# r02 GPU call 2: new AES kernels (parity + timings + ncu), hybrid host path, pipelined matmul host call
set -u
out=gpurun_out/r02c2
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -8 "$out/pytest_gpu.log"
{
for nc in 1 2 3; do
  python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time
  python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --inject 0.0009765625
done
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1 --inject 0.0009765625
python tools/profile_target.py --kernel aes --nc 3 --log2n 24 --iters 10 --time --aes-mode 1
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 2
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 3
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 7
} > "$out/aes_timings.txt" 2>&1
cat "$out/aes_timings.txt"
{
for nc in 1 2 3; do python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 20 --time; done
python tools/profile_target.py --kernel gemm --nc 3 --side 4096 --iters 20 --time --inject 0.001
for g in 8 32; do COAST_GEMM_GROUP_M=$g python tools/profile_target.py --kernel gemm --nc 3 --side 4096 --iters 20 --time; done
python tools/profile_target.py --kernel gemm --nc 1 --side 8192 --iters 10 --time
python tools/profile_target.py --kernel gemm --nc 3 --side 8192 --iters 10 --time
} > "$out/gemm_timings.txt" 2>&1
cat "$out/gemm_timings.txt"
for hp in staged hybrid; do
  for wl in sha256 aes crc16; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --workload $wl --host-path $hp > "$out/bench_${wl}_${hp}.json" 2> "$out/bench_${wl}_${hp}.err"; echo "bench $wl $hp rc=$?" | tee -a "$out/summary.txt"
  done
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --workload gemm > "$out/bench_gemm.json" 2> "$out/bench_gemm.err"; echo "bench gemm rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xmr_aes128_enc_nc2_inj1 -c 1 -o "$out/aes_enc_nc2_inj1" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 --inject 0.0009765625 > "$out/ncu1.log" 2>&1; echo "ncu1 rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xmr_aes128_enc_nc2_inj0 -c 1 -o "$out/aes_enc_nc2_inj0" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 > "$out/ncu2.log" 2>&1; echo "ncu2 rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xmr_aes128_dec_nc2_inj0 -c 1 -o "$out/aes_dec_nc2_inj0" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 --aes-mode 1 > "$out/ncu3.log" 2>&1; echo "ncu3 rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xmr_gemm_tf32_nc3_inj0 -c 1 -o "$out/gemm_nc3" python tools/profile_target.py --kernel gemm --nc 3 --side 4096 --iters 2 > "$out/ncu4.log" 2>&1; echo "ncu4 rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xmr_gemm_tf32_nc1_inj0 -c 1 -o "$out/gemm_nc1" python tools/profile_target.py --kernel gemm --nc 1 --side 4096 --iters 2 > "$out/ncu5.log" 2>&1; echo "ncu5 rc=$?" | tee -a "$out/summary.txt"
COAST_GEMM_GROUP_M=32 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:xmr_gemm_tf32_nc3_inj0 -c 1 --csv --log-file "$out/gemm_nc3_g32_traffic.csv" python tools/profile_target.py --kernel gemm --nc 3 --side 4096 --iters 2 > "$out/ncu6.log" 2>&1; echo "ncu6 rc=$?" | tee -a "$out/summary.txt"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c2/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsed',e); continue
    e=d.get('e2e',{})
    print(f, d.get('value'), d.get('ms_per_step'), 'e2e', e.get('value'), e.get('ms_per_step'), e.get('path'), e.get('frac_of_bound'), e.get('pcie_pinned_copy_gbs'))
PY
