"""Child of tests/test_gpu_peer_counters.py: a second process on ANOTHER GPU that maps the parent's counter block over
NVLink/PCIe peer memory and runs its shard.  Usage: python peer_counters_child.py <device> <handle-hex> <in.npy> <out.npy>
<unit_base> <nc> <seed> <threshold>"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import coast_b200 as cb
    device, handle, fin, fout = int(sys.argv[1]), bytes.fromhex(sys.argv[2]), sys.argv[3], sys.argv[4]
    unit_base, nc, seed, threshold = int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8])
    torch.cuda.set_device(device)
    rt = cb.Runtime(device)
    rt.counters_attach(handle)
    m = np.load(fin)
    n = m.size // 64
    plan = cb.FaultPlan(mode=cb.PLAN_BERNOULLI, seed=seed, threshold=threshold)
    out, st = rt.run(cb.K_SHA256, nc, torch.from_numpy(m).cuda(), n, flags=3, unit_bytes=64, plan=plan, unit_base=unit_base)
    np.save(fout, out.cpu().numpy())
    mine = st.as_dict()                                     # an attached process reports zeros: its tallies live in the owner's block
    rt.counters_detach()
    out2, st2 = rt.run(cb.K_SHA256, nc, torch.from_numpy(m).cuda(), n, flags=3, unit_bytes=64, plan=plan, unit_base=unit_base)
    print(json.dumps({"attached": mine, "detached": st2.as_dict(), "same_out": bool(torch.equal(out, out2))}))


if __name__ == "__main__":
    main()
