#!/bin/bash
set -u
out=gpurun_out/r02mmkeep
mkdir -p "$out"
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mm_tensor_core or mm_full_size" > "$out/pytest.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest.txt"
for k in 1 0; do echo "MM_KEEP_A=$k nc=3" | tee -a "$out/timings.txt"; COAST_MM_KEEP_A=$k timeout 60 python tools/profile_target.py --kernel mm --nc 3 --side 4096 --iters 5 --time 2>&1 | tail -1 | cut -c1-100 | tee -a "$out/timings.txt"; done
