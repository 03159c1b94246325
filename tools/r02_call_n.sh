#!/bin/bash
# r02 multi-GPU call: bash tools/r02_call_n.sh N  -- the bench as the driver launches it at N GPUs (both arms), plus the gloo-free NCCL sanity of the also-workloads
set -u
N=${1:-2}
out=gpurun_out/r02n$N
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 \
    > "$out/bench_n$N.json" 2> "$out/bench_n$N.err"; echo "bench N=$N rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/bench_n$N.err"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 20 --warmup 5 --ref-budget-s 20 \
    > "$out/bench_ref_n$N.json" 2> "$out/bench_ref_n$N.err"; echo "ref N=$N rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-also --no-cpu-baseline > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench N=1 rc=$?" | tee -a "$out/summary.txt"
python - "$N" <<'PY'
import json,sys
N=sys.argv[1]
def last(f):
    try: return json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e: print(f,'unparsed',e); return None
a=last(f'gpurun_out/r02n{N}/bench_n{N}.json'); b=last(f'gpurun_out/r02n{N}/bench_n1.json')
if a and b:
    print('N',N,'value',a['value'],'ms',a['ms_per_step'],'N=1',b['value'],'eff',round(a['value']/(int(N)*b['value']),4),'collectives',a.get('collectives'))
    print('e2e',a['e2e']['value'],a['e2e']['ms_per_step'],'N=1 e2e',b['e2e']['value'])
    for k,v in (a.get('also') or {}).items():
        print('  also',k,v.get('value'),v.get('ms_per_step'),v.get('scaling'),(v.get('roofline') or {}).get('frac'),v.get('collectives',{}).get('in_value_ms'),v.get('error'))
PY
