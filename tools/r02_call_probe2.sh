#!/bin/bash
set -u
out=gpurun_out/r02probe2
mkdir -p "$out"
for i in 1 2; do
for p in 1 0; do
  COAST_BENCH_CLOCK_PROBE=$p timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > "$out/bench_p${p}_$i.json" 2> "$out/bench_p${p}_$i.err"
  python - "$out/bench_p${p}_$i.json" "$p" <<'PY'
import json,sys
a=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('probe',sys.argv[2],'sha',a['ms_per_step'],{k:(v.get('ms_per_step'),v.get('roofline',{}).get('kernel_ms_single_launch_events'),v.get('clocks',{}).get('sm_mhz_in_timed_region')) for k,v in a['also'].items() if k in('aes','gemm','crc16')})
PY
done
done
