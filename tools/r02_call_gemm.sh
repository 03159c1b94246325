#!/bin/bash
# r02 one-GPU call: GEMM tests + timings after the tail-split / 8-warp epilogue change
set -u
out=gpurun_out/r02gemm
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > "$out/pytest.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -5 "$out/pytest.txt"
for nc in 1 2 3; do timeout 300 python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 30 --time 2>&1 | tail -2 | tee -a "$out/gemm_timings.txt"; done
COAST_GEMM_TAIL_SPLIT=0 timeout 300 python tools/profile_target.py --kernel gemm --nc 1 --side 4096 --iters 30 --time 2>&1 | tail -2 | tee -a "$out/gemm_timings.txt"
timeout 300 python tools/profile_target.py --kernel gemm --nc 1 --side 8192 --iters 10 --time 2>&1 | tail -2 | tee -a "$out/gemm_timings.txt"
timeout 300 python bench.py --workload gemm --steps 40 --warmup 5 --no-cpu-baseline > "$out/bench_gemm.json" 2> "$out/bench_gemm.err"; echo "bench gemm rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:xmr_gemm_tf32_nc1 -c 1 -o "$out/gemm_nc1" python tools/profile_target.py --kernel gemm --nc 1 --side 4096 --iters 1 > "$out/ncu1.log" 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py "$out/gemm_nc1.ncu-rep" > "$out/gemm_nc1.json" 2>> "$out/ncu1.log"; rm -f "$out"/*.ncu-rep
