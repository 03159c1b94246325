#!/bin/bash
# r02 one-GPU call: the default bench line with the clock probes, plus the gemm workload on its own
set -u
out=gpurun_out/r02probe
mkdir -p "$out"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/bench_default.json" 2> "$out/bench_default.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
python tools/show_bench.py "$out/bench_default.json" 2>/dev/null | head -20
python - <<'PY'
import json
a=json.loads([l for l in open('gpurun_out/r02probe/bench_default.json') if l.startswith('{')][-1])
print('main clocks',a['clocks'])
for k,v in a['also'].items():
    if 'clocks' in v: print(k,{x:v['clocks'].get(x) for x in ('sm_mhz','sm_mhz_in_timed_region','power_w_max','reasons')}, (v.get('roofline') or {}).get('frac_of_clock_scaled_hw_rate'))
PY
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -x -q -m gpu -k "gemm or fill or counters" 2>&1 | tail -3
