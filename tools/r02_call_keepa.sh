#!/bin/bash
set -u
out=gpurun_out/r02keepa
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > "$out/pytest.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -5 "$out/pytest.txt"
for rep in 1 2; do
for k in 0 1; do for p in 0 1; do for nc in 2 3; do echo "KEEP_A=$k PAIR=$p nc=$nc" | tee -a "$out/timings.txt"; COAST_GEMM_KEEP_A=$k COAST_GEMM_PAIR=$p timeout 120 python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 30 --time 2>&1 | tail -2 | head -1 | cut -c1-90 | tee -a "$out/timings.txt"; done; done; done
done
