#!/bin/bash
# r02 scaling call on ONE 8-GPU box: the bench exactly as the driver launches it at N = 1, 2, 4, 8 (default steps), back to back
set -u
out=gpurun_out/r02scale
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-also --no-cpu-baseline > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench N=1 rc=$?" | tee -a "$out/summary.txt"
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 20 --warmup 5 \
      > "$out/bench_n$N.json" 2> "$out/bench_n$N.err"; echo "bench N=$N rc=$?" | tee -a "$out/summary.txt"
done
python - <<'PY' | tee -a gpurun_out/r02scale/summary.txt
import json
def last(f):
    try: return json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e: print(f,'unparsed',e); return None
b=last('gpurun_out/r02scale/bench_n1.json')
for N in (2,4,8):
    a=last(f'gpurun_out/r02scale/bench_n{N}.json')
    if a and b:
        print('N',N,'value',a['value'],'ms',a['ms_per_step'],'eff',round(a['value']/(N*b['value']),4),'fold',a['collectives']['counter_fold'],'e2e',a['e2e']['value'],'vs N=1 e2e x',round(a['e2e']['value']/b['e2e']['value'],2))
        for k,v in (a.get('also') or {}).items():
            print('   also',k,v.get('value'),v.get('ms_per_step'),(v.get('roofline') or {}).get('frac'),v.get('error'))
PY
