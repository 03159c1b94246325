#!/bin/bash
# compute-sanitizer pass over every kernel family (small sizes); run under gpurun.  Output: gpurun_out/sanitizer.log
set -u
out=gpurun_out/sanitizer.log; : > $out
run() { echo "=== $*" >> $out; timeout 300 compute-sanitizer --tool $1 --error-exitcode 9 python tools/profile_target.py "${@:2}" >> $out 2>&1; echo "--- exit $?" >> $out; }
for tool in memcheck racecheck; do
  run $tool --kernel sha256 --nc 3 --log2n 12 --iters 1
  run $tool --kernel sha256 --nc 3 --log2n 12 --iters 1 --flags 0x8 --inject 0.1
  run $tool --kernel sha256 --nc 2 --log2n 12 --iters 1
  run $tool --kernel aes --nc 2 --log2n 13 --iters 1 --inject 0.1
  run $tool --kernel aes --nc 3 --log2n 13 --iters 1
  run $tool --kernel crc16 --nc 3 --log2n 12 --iters 1 --inject 0.1
  run $tool --kernel mm --nc 3 --side 256 --iters 1 --inject 0.01
  COAST_MM_PATH=tiled run $tool --kernel mm --nc 3 --side 256 --iters 1 --inject 0.01
  run $tool --kernel gemm --nc 3 --side 256 --iters 1 --inject 0.01
done
grep -E "^===|--- exit|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" $out | head -80
