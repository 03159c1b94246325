#!/bin/bash
# r02 two-GPU call: the peer-counter-block test, then the bench at N=2 exactly as the driver launches it, then N=1 on the same box
set -u
out=gpurun_out/r02peer
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_peer_counters.py tests/test_gpu_gemm.py -x -q -m gpu > "$out/pytest.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -5 "$out/pytest.txt"
for wl in gemm; do
  for nc in 1 2 3; do timeout 300 python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 30 --time 2>&1 | tail -2 | tee -a "$out/gemm_timings.txt"; done
done
bash tools/r02_call_n.sh 2
