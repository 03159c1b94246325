#!/bin/bash
# r02 GPU call 1: existing + new parity tests, the new bench line (also: aes/gemm/crc16/2^30), staged-vs-zero-copy e2e, CPU arm stability
set -u
out=gpurun_out/r02c1
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c > "$out/numa_nodes.txt"
lscpu | head -30 > "$out/lscpu.txt"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -5 "$out/pytest_gpu.log"
timeout 600 python bench.py --steps 20 --warmup 5 > "$out/bench_default.json" 2> "$out/bench_default.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/bench_default.err"
timeout 300 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --host-path staged > "$out/bench_staged.json" 2> "$out/bench_staged.err"; echo "bench staged rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --workload aes --host-path staged > "$out/bench_aes_staged.json" 2> "$out/bench_aes_staged.err"; echo "bench aes staged rc=$?" | tee -a "$out/summary.txt"
COAST_NUMA_BIND=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline > "$out/bench_nonuma.json" 2> "$out/bench_nonuma.err"; echo "bench nonuma rc=$?" | tee -a "$out/summary.txt"
for i in 1 2; do
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 --ref-budget-s 25 > "$out/bench_ref_$i.json" 2> "$out/bench_ref_$i.err"; echo "ref $i rc=$?" | tee -a "$out/summary.txt"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c1/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsed',e); continue
    print(f, d.get('value'), d.get('ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_step'), d.get('e2e',{}).get('path'), d.get('e2e',{}).get('frac_of_bound'), d.get('e2e',{}).get('numa_node'), d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('spread'))
    for k,v in (d.get('also') or {}).items():
        print('   also',k, v.get('value'), v.get('ms_per_step'), 'e2e', v.get('e2e',{}).get('value'), v.get('e2e',{}).get('path'), v.get('e2e',{}).get('frac_of_bound'), (v.get('roofline') or {}).get('frac'), v.get('cpu_baseline',{}).get('value'), v.get('error'), v.get('wall_s'))
PY
