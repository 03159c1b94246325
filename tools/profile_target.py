#!/usr/bin/env python3
"""Small launch driver for ncu captures (never a bench number): runs one workload/protection combo
`--iters` times on device-resident Philox inputs.

  ncu --set full --clock-control none --import-source on -k regex:xmr_sha256 -o gpurun_out/prof \
      python tools/profile_target.py --kernel sha256 --nc 3 --iters 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", choices=["sha256", "aes", "crc16", "mm", "gemm", "qsort", "chsha"], default="sha256")
    ap.add_argument("--nc", type=int, default=3)
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--inject", type=float, default=0.0)
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--stream-bytes", type=int, default=16384, help="chsha: bytes per stream (the benchmark's is 2 x 8192)")
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0, help="extra COAST_F_* bits, e.g. 0x8 = -i, 0x10 = -s")
    ap.add_argument("--table-density", type=float, default=-1.0, help="use a TABLE plan with this fraction of units faulted (0 = all-zero table)")
    ap.add_argument("--threshold", type=int, default=-1, help="Bernoulli plan with this raw threshold (0 = the injector kernel never hits)")
    ap.add_argument("--aes-mode", type=lambda x: int(x, 0), default=0, help="aes: COAST_AES_* bits (1 decrypt, 2 per-unit keys, 4 key write-back)")
    ap.add_argument("--time", action="store_true", help="print CUDA-event ms per launch (outside any profiler)")
    a = ap.parse_args()
    import torch
    import coast_b200 as cb
    rt = cb.Runtime(0)
    n = 1 << a.log2n
    plan = cb.FaultPlan(mode=cb.PLAN_BERNOULLI, seed=7, p=a.inject) if a.inject > 0 else None
    if a.threshold >= 0:
        plan = cb.FaultPlan(mode=cb.PLAN_BERNOULLI, seed=7, threshold=a.threshold)
    if a.table_density >= 0:
        nn = 1 << a.log2n
        tab = torch.zeros(nn, dtype=torch.int32, device="cuda")
        if a.table_density > 0:
            g = torch.Generator(device="cuda"); g.manual_seed(3)
            hit = torch.rand(nn, device="cuda", generator=g) < a.table_density
            site = torch.randint(0, 176, (nn,), device="cuda", generator=g, dtype=torch.int32)
            rep = torch.randint(0, max(a.nc, 1), (nn,), device="cuda", generator=g, dtype=torch.int32)
            bit = torch.randint(0, 8, (nn,), device="cuda", generator=g, dtype=torch.int32)
            ent = (torch.full((nn,), -2 ** 31, dtype=torch.int32, device="cuda") | (rep << 29) | (site << 5) | bit)
            tab = torch.where(hit, ent, tab)
        plan = cb.FaultPlan(mode=cb.PLAN_TABLE, table=tab)
    flags = cb.F_COUNT_ERRORS | cb.F_COUNT_SYNCS | a.flags
    if a.kernel == "sha256":
        d_in = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); rt.fill_philox(d_in, 2)
        out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        d = rt.make_desc(cb.K_SHA256, a.nc, d_in, out, n, flags=flags, unit_bytes=64, plan=plan)
        alg = n * 96
    elif a.kernel == "crc16":
        d_in = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); rt.fill_philox(d_in, 1)
        out = torch.empty(n * 2, dtype=torch.uint8, device="cuda")
        d = rt.make_desc(cb.K_CRC16, a.nc, d_in, out, n, flags=flags, unit_bytes=64, plan=plan)
        alg = n * 66
    elif a.kernel == "aes":
        d_in = torch.empty(n * 16, dtype=torch.uint8, device="cuda"); rt.fill_philox(d_in, 3)
        out = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
        keys = None
        if a.aes_mode & 2:
            keys = torch.empty(n * 16, dtype=torch.uint8, device="cuda"); rt.fill_philox(keys, 5)
        d = rt.make_desc(cb.K_AES128, a.nc, d_in, out, n, flags=flags, key=bytes(16), plan=plan, mode=a.aes_mode, d_aux=keys)
        alg = n * (48 if keys is not None else 32)
    elif a.kernel == "chsha":
        ub = a.stream_bytes
        d_in = torch.empty(n * ub, dtype=torch.uint8, device="cuda"); rt.fill_philox(d_in, 6)
        out = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
        d = rt.make_desc(cb.K_CHSTONE_SHA, a.nc, d_in, out, n, flags=flags, unit_bytes=ub, plan=plan)
        alg = n * (ub + 20)
    elif a.kernel == "qsort":
        L = 580
        d_in = torch.empty(n * L, dtype=torch.int32, device="cuda"); rt.fill_philox(d_in, 9)
        out = torch.empty(n * L, dtype=torch.int32, device="cuda")
        d = rt.make_desc(cb.K_QSORT, a.nc, d_in, out, n, flags=flags, unit_bytes=4 * L, plan=plan)
        alg = n * L * 8
    elif a.kernel == "gemm":
        s = a.side
        A = torch.rand(s * s, dtype=torch.float32, device="cuda") * 2 - 1
        B = torch.rand(s * s, dtype=torch.float32, device="cuda") * 2 - 1
        out = torch.empty(s * s, dtype=torch.float32, device="cuda")
        d = rt.make_desc(cb.K_GEMM_TF32, a.nc, A, out, s * s, flags=flags, M=s, N=s, K=s, d_aux=B, plan=plan)
        alg = 3 * s * s * 4
        flops = 2.0 * s * s * s
    else:
        s = a.side
        A = torch.empty(s * s, dtype=torch.int32, device="cuda"); rt.fill_philox(A, 4)
        B = torch.empty(s * s, dtype=torch.int32, device="cuda"); rt.fill_philox(B, 44)
        out = torch.empty(s * s, dtype=torch.int32, device="cuda")
        d = rt.make_desc(cb.K_MM_U32, a.nc, A, out, s * s, flags=flags, M=s, N=s, K=s, d_aux=B, plan=plan)
        alg = 3 * s * s * 4
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rt.launch(d); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    st = rt.sync()
    if a.time:
        best = min(ms)
        if a.kernel in ("mm", "gemm"):
            n = a.side * a.side
        print(f"{a.kernel} nc={a.nc} n={n} inject={a.inject}: best {best:.4f} ms, median {sorted(ms)[len(ms)//2]:.4f} ms, "
              f"{alg / best / 1e6:.1f} GB/s algorithmic; stats={st.as_dict()}")
        if a.kernel == "gemm":
            print(f"   useful {flops / best / 1e9:.1f} TFLOP/s, issued {a.nc * flops / best / 1e9:.1f} TFLOP/s (tf32)")


if __name__ == "__main__":
    main()
