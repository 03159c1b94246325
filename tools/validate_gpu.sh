#!/bin/bash
# One gpurun call = the whole round-end check (what the driver runs on a fresh B200), so a round spends one box
# acquisition on it instead of several:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/validate_gpu.sh'
# Outputs land in gpurun_out/validate/ (merged back): pytest log, smoke log, the default bench line, the extra bench lines
# and the ncu launch list of the bench command (gpu__time_duration per launch; never a bench number).
set -u
out=gpurun_out/validate
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_gpu.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py > "$out/bench_sha256_n1.json" 2> "$out/bench_sha256_n1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
for wl in crc16 aes gemm; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 > "$out/bench_${wl}_n1.json" 2> "$out/bench_${wl}_n1.err"; echo "bench $wl rc=$?" | tee -a "$out/summary.txt"
done
if [ "${1:-}" = "--ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" \
      python bench.py --steps 5 --warmup 3 > "$out/bench_under_ncu.log" 2>&1; echo "ncu launch list rc=$?" | tee -a "$out/summary.txt"
fi
cut -c1-300 "$out/bench_sha256_n1.json"
