#!/bin/bash
# r02 one-GPU call: bring-up of the CTA-pair GEMM kernels (every step under its own timeout; the kernels trap instead of spinning)
set -u
out=gpurun_out/r02pair
mkdir -p "$out"
timeout 120 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "pair and 256-256-32" > "$out/pytest_small.txt" 2>&1; echo "small rc=$?" | tee -a "$out/summary.txt"
tail -15 "$out/pytest_small.txt"
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > "$out/pytest.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -15 "$out/pytest.txt"
for p in 0 1; do for nc in 1 2 3; do echo "PAIR=$p nc=$nc" | tee -a "$out/gemm_timings.txt"; COAST_GEMM_PAIR=$p timeout 120 python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 30 --time 2>&1 | tail -2 | tee -a "$out/gemm_timings.txt"; done; done
COAST_GEMM_PAIR=1 timeout 120 python tools/profile_target.py --kernel gemm --nc 1 --side 8192 --iters 10 --time 2>&1 | tail -2 | tee -a "$out/gemm_timings.txt"
nvidia-smi --query-gpu=name,clocks.sm,temperature.gpu --format=csv | tee -a "$out/summary.txt"
