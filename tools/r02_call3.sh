#!/bin/bash
# r02 GPU call 3 (1 GPU): the round-end check on the final tree -- full parity suite, smoke, the default bench line as the driver runs it,
# the ncu launch list of that command, a sanitizer pass over the kernels that changed in r02
set -u
out=gpurun_out/r02c3
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -6 "$out/pytest_gpu.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"
{
for nc in 1 2 3; do
  python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time
  python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --inject 0.0009765625
done
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1
python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1 --inject 0.0009765625
for nc in 1 2 3; do python tools/profile_target.py --kernel gemm --nc $nc --side 4096 --iters 20 --time; done
} > "$out/timings.txt" 2>&1
cat "$out/timings.txt"
timeout 600 ncu --set full --clock-control none -k regex:xmr_aes128_enc_nc2_inj1 -c 1 -o "$out/aes_nc2_inj1" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 --inject 0.0009765625 > "$out/ncu_aes.log" 2>&1
python tools/ncu_summary.py "$out/aes_nc2_inj1.ncu-rep" "$out/aes_nc2_inj1.json" > /dev/null 2>&1; rm -f "$out/aes_nc2_inj1.ncu-rep"
for cfg in "3 0 gemm_nc3" "1 1 gemm_nc1_pair" "1 0 gemm_nc1_single" "2 1 gemm_nc2_pair"; do
  set -- $cfg
  COAST_GEMM_PAIR=$2 timeout 600 ncu --set full --clock-control none -k regex:xmr_gemm_tf32 -c 1 -o "$out/$3" python tools/profile_target.py --kernel gemm --nc $1 --side 4096 --iters 1 > "$out/ncu_$3.log" 2>&1
  python tools/ncu_summary.py "$out/$3.ncu-rep" "$out/$3.json" > /dev/null 2>&1; rm -f "$out/$3.ncu-rep"
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/bench_default.json" 2> "$out/bench_default.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 --ref-budget-s 25 > "$out/bench_reference.json" 2> "$out/bench_reference.err"; echo "ref rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file "$out/launches_bench.csv" \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench_under_ncu.log" 2>&1; echo "ncu launch list rc=$?" | tee -a "$out/summary.txt"
{
run() { echo "=== $*"; timeout 300 compute-sanitizer --tool $1 --error-exitcode 9 python tools/profile_target.py "${@:2}" 2>&1 | tail -4; echo "--- exit $?"; }
for tool in memcheck racecheck; do
  run $tool --kernel aes --nc 2 --log2n 13 --iters 1 --inject 0.1
  run $tool --kernel aes --nc 3 --log2n 13 --iters 1 --aes-mode 1 --inject 0.1
  run $tool --kernel aes --nc 2 --log2n 13 --iters 1 --aes-mode 7
  run $tool --kernel gemm --nc 3 --side 256 --iters 1 --inject 0.01
  run $tool --kernel gemm --nc 1 --side 256 --iters 1
  run $tool --kernel crc16 --nc 3 --log2n 12 --iters 1 --inject 0.1 --flags 0x200
  run $tool --kernel gemm --nc 2 --side 512 --iters 1 --inject 0.01
  COAST_GEMM_PAIR=1 run $tool --kernel gemm --nc 3 --side 512 --iters 1 --inject 0.01
  run $tool --kernel sha256 --nc 3 --log2n 12 --iters 1 --inject 0.1
done
} > "$out/sanitizer.log" 2>&1
grep -E "^===|--- exit|ERROR SUMMARY|RACECHECK SUMMARY" "$out/sanitizer.log" | head -60
python tools/show_bench.py "$out/bench_default.json" "$out/bench_default.err" 2>/dev/null | head -40
python -c "import json;d=json.load(open('$out/aes_nc2_inj1.json'))[0];print({k:d.get(k) for k in ('kernel','duration','pipe_alu_pct','pipe_lsu_pct','warp_insts','issue_per_cycle_per_smsp','registers')})"
du -sh gpurun_out
