#!/usr/bin/env python3
"""bench.py -- headline measurement of the protected-region hot path.

Workload (BASELINE.json configs[1]): SHA-256 under TMR, 2^20 x 64-byte messages per GPU,
warp-shuffle select voter, -countErrors -countSyncs.  Metric: MB/s of VOTED OUTPUT
(32 digest bytes per message).  One "step" = one protected launch over the whole batch.

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference ...                      # the reference's own C sources (oracle/_ref,
                                                            # else the oracle port) under CPU TMR, all host threads
Under torchrun (N>1) every rank hashes its own 2^20-message shard (weak scaling, no data-path
collective); the only exchange is an NCCL all-reduce of the 4 fault counters per step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UNITS = 1 << 20          # messages per GPU
UNIT_BYTES = 64
OUT_BYTES = 32
ALG_BYTES_PER_UNIT = 96    # 64 in + 32 out
NSETS = 4                  # rotating in/out buffer sets: 4 x 96 MiB = 384 MiB > 126 MB L2
METRIC = "protected-kernel throughput (MB/s voted output), sha256 TMR"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons via NVML while the timed regions run."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period = 0.002
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self, period=0.002):
        """period: NVML polling interval.  2 ms inside the device-timed region (GPU-bound, launches are cheap); 25 ms inside
        the host-call region, where NVML queries contend with the copy/launch submissions on the driver lock and were
        measured to slow the pipeline by up to 1.8x."""
        self.period = period
        if self.nv:
            self._stop.clear()
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()

    def stop(self):
        if self._th:
            self._stop.set()
            self._th.join()
            self._th = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# one name per workload, shared by both arms so that `config.workload` is identical in the two JSON lines the driver compares
WORKLOAD_NAMES = {
    "sha256": "sha256 TMR, 2^20 x 64-byte messages per GPU (BASELINE configs[1])",
    "sha256_2p30": "batched sha256 TMR, 2^30 x 64-byte messages sharded over the GPUs (BASELINE configs[4])",
    "aes": "aes-128 ECB encrypt DWC, 2^24 x 16-byte blocks, Bernoulli(2^-10) single-bit flips (BASELINE configs[2])",
    "crc16": "crc16 TMR, 2^20 x 64-byte messages (SURVEY.md 8d config 1 timing shape)",
    "gemm": "matmul TMR 4096x4096x4096 fp32 on tcgen05 kind::tf32, three TMEM accumulator replicas + voter (BASELINE configs[3])",
}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sha256_hash() under the restated TMR wrapper (oracle/_ref), or the port
# ------------------------------------------------------------------------------------------------
def cpu_tmr_sha(n_units: int, threads: int, repeats: int = 1):
    """Returns (seconds per pass, kind).  Inputs are Philox(seed=2) bytes, same generator as the GPU arm."""
    import ctypes as C
    import numpy as np
    from oracle import pyoracle as po
    po.build()
    msgs = po.fill_philox(n_units * UNIT_BYTES // 4, 0, 2).view(np.uint8)
    out = np.zeros(n_units * OUT_BYTES, dtype=np.uint8)
    if po.ref_available():
        lib = po.ref("sha256")
        lib.ref_sha256_xmr_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                          C.c_int, C.POINTER(po.RefStats)]

        def one():
            st = po.RefStats()
            lib.ref_sha256_xmr_mt(msgs.ctypes.data, out.ctypes.data, n_units, UNIT_BYTES, 3, 1, 1, threads, C.byref(st))
        kind = "reference"
    else:
        def one():
            po.run(po.K_SHA256, 3, msgs, n_units, unit_bytes=UNIT_BYTES, flags=3, threads=threads)
        kind = "port"
    t0 = time.perf_counter()
    for _ in range(repeats):
        one()
    return (time.perf_counter() - t0) / repeats, kind


def cpu_baseline_block(budget_s: float = 12.0):
    threads = os.cpu_count() or 1
    cal_n = 1 << 14
    t_cal, kind = cpu_tmr_sha(cal_n, threads)
    rate = cal_n / max(t_cal, 1e-6)
    n = int(min(N_UNITS * 4, max(1 << 15, rate * budget_s)))
    n = 1 << (n.bit_length() - 1)
    t, kind = cpu_tmr_sha(n, threads)
    return {"value": round(n * OUT_BYTES / t / 1e6, 3), "unit": "MB/s", "cores": threads, "kind": kind,
            "sample": f"{n} x 64-byte messages, Philox(seed=2), TMR + countErrors, {threads} pthreads, {t:.2f} s"}


def cpu_xmr(workload: str, n_units: int, threads: int, repeats: int = 1):
    """CPU arm of any workload: (seconds per pass, kind, out_bytes_per_unit).  sha256 runs the reference's own
    sha256_hash() (oracle/_ref); the others run the oracle port (oracle/coast_oracle.c) with pthreads."""
    import numpy as np
    from oracle import pyoracle as po
    if workload.startswith("sha256"):
        t, kind = cpu_tmr_sha(n_units, threads, repeats)
        return t, kind, OUT_BYTES
    po.build()
    if workload == "crc16":
        inp = po.fill_philox(n_units * 16, 0, 2).view(np.uint8)
        if po.ref_available():                              # the reference's own crc16() under the restated TMR wrapper
            import ctypes as C
            lib = po.ref("crc16")
            lib.ref_crc16_xmr_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                             C.c_int, C.POINTER(po.RefStats)]
            out = np.zeros(n_units, dtype=np.uint16)
            t0 = time.perf_counter()
            for _ in range(repeats):
                st = po.RefStats()
                lib.ref_crc16_xmr_mt(inp.ctypes.data, out.ctypes.data, n_units, 64, 3, 1, 1, threads, C.byref(st))
            return (time.perf_counter() - t0) / repeats, "reference", 2
        kw = dict(kernel=po.K_CRC16, nc=3, flags=3, unit_bytes=64)
        ob = 2
    elif workload == "aes":
        inp = po.fill_philox(n_units * 4, 0, 2).view(np.uint8)
        if po.ref_available():
            # the reference's own aes_enc_dec() under the restated DWC wrapper.  Its flips can only go into a replica's private
            # copy of the INPUT (mid-round sites need edited sources): Bernoulli(2^-10) per block, as in the GPU arm's plan.
            import ctypes as C
            lib = po.ref("aes")
            lib.ref_aes_xmr_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                           C.c_int, C.c_void_p, C.c_int, C.POINTER(po.RefStats)]
            rng = np.random.default_rng(33)
            hit = rng.random(n_units) < 2.0 ** -10
            faults = np.zeros((n_units, 3), dtype=np.int32)           # ref_fault {replica, byte, bit}
            faults[:, 1] = -1
            k = int(hit.sum())
            faults[hit] = np.stack([rng.integers(0, 2, k), rng.integers(0, 16, k), rng.integers(0, 8, k)], axis=1)
            key = np.zeros(16, dtype=np.uint8)
            out = np.zeros(n_units * 16, dtype=np.uint8)
            t0 = time.perf_counter()
            for _ in range(repeats):
                st = po.RefStats()
                lib.ref_aes_xmr_mt(inp.ctypes.data, out.ctypes.data, n_units, key.ctypes.data, 0, 0, 2, 0, 0, faults.ctypes.data,
                                   threads, C.byref(st))
                assert st.dwc_detected == st.injected == k      # detect-rate parity holds on the CPU arm too
            return (time.perf_counter() - t0) / repeats, "reference", 16
        kw = dict(kernel=po.K_AES128, nc=2, flags=0, key=bytes(16), plan=po.make_plan(po.PLAN_BERNOULLI, seed=33, p=2.0 ** -10))
        ob = 16
    else:  # gemm: a row-block sample of the 4096^3 problem (n_units = rows * 4096)
        side = 4096
        rows = max(1, n_units // side)
        A = (po.fill_philox(rows * side, 0, 4).astype(np.float64) / 2 ** 31 - 1.0).astype(np.float32)
        B = (po.fill_philox(side * side, 0, 44).astype(np.float64) / 2 ** 31 - 1.0).astype(np.float32)
        kw = dict(kernel=po.K_GEMM_TF32, nc=3, flags=3, M=rows, N=side, K=side, aux=B)
        inp, n_units, ob = A, rows * side, 4
    kernel, nc, flags = kw.pop("kernel"), kw.pop("nc"), kw.pop("flags")
    t0 = time.perf_counter()
    for _ in range(repeats):
        po.run(kernel, nc, inp, n_units, flags=flags, threads=threads, **kw)
    return (time.perf_counter() - t0) / repeats, "port", ob


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.threads or (os.cpu_count() or 1)
    wl = args.workload
    cal_n = {"gemm": 4096 * 2, "aes": 1 << 16}.get(wl, 1 << 14)
    t_cal, kind, ob = cpu_xmr(wl, cal_n, threads)
    rate = cal_n / max(t_cal, 1e-6)
    total_budget = args.ref_budget_s                      # whole --steps/--warmup run stays within a few minutes
    full = {"sha256": N_UNITS, "sha256_2p30": N_UNITS, "crc16": 1 << 20, "aes": 1 << 24, "gemm": 4096 * 4096}[wl]
    n = int(max(cal_n, min(full, rate * total_budget / max(1, args.steps + args.warmup))))
    n = (n // 4096) * 4096 if wl == "gemm" else 1 << (n.bit_length() - 1)
    for _ in range(args.warmup):
        cpu_xmr(wl, n, threads)
    t, kind, ob = cpu_xmr(wl, n, threads, repeats=args.steps)
    val = n * ob / t / 1e6
    line = {
        "impl": "reference",
        "metric": METRIC if wl.startswith("sha256") else f"protected-kernel throughput (MB/s voted output), {wl}",
        "value": round(val, 3), "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True,
        "scaling": "strong" if wl == "gemm" else "weak", "vs_baseline": None,
        "dtype": "f32" if wl == "gemm" else ("u8" if wl == "aes" else "u32"), "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[wl], "units_per_step": n, "protection": {"aes": "-DWC + injector", }.get(wl, "-TMR -countErrors -countSyncs"),
                   "note": "reference C sources compiled in place (oracle/_ref) + restated xMR wrapper (sha256, crc16, aes; aes flips go "
                           "into a replica's input copy); oracle port for the fp32 matmul; the real opt -TMR binary needs LLVM 7 "
                           "(absent). A step is a bounded sample of the GPU arm's batch."},
        "cpu_baseline": {"value": round(val, 3), "unit": "MB/s", "cores": threads, "kind": kind,
                         "sample": f"{n} units per step, {threads} pthreads"},
        "e2e": {"value": round(val, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def workload_table(cb):
    """BASELINE.json configs as bench workloads.  `sha256` (configs[1]) is the default and the headline line; the others
    are extra lines for the remaining single-GPU configurations (python bench.py --workload aes|gemm|crc16)."""
    F = cb.F_COUNT_ERRORS | cb.F_COUNT_SYNCS
    return {
        "sha256": dict(kernel=cb.K_SHA256, nc=3, flags=F, n=1 << 20, unit_bytes=64, in_b=64, out_b=32, alg_b=96, plan=None,
                       name=WORKLOAD_NAMES["sha256"], kname="xmr_sha256_b64_seg_nc3_inj0",
                       protection="-TMR -countErrors -countSyncs", bound="hbm", sets=4),
        "sha256_2p30": dict(kernel=cb.K_SHA256, nc=3, flags=F, n=1 << 30, unit_bytes=64, in_b=64, out_b=32, alg_b=96, plan=None, strong=True,
                            name=WORKLOAD_NAMES["sha256_2p30"],
                            kname="xmr_sha256_b64_seg_nc3_inj0", protection="-TMR -countErrors -countSyncs", bound="hbm", sets=1),
        "aes": dict(kernel=cb.K_AES128, nc=2, flags=0, n=1 << 24, unit_bytes=0, in_b=16, out_b=16, alg_b=32,
                    plan=dict(seed=33, p=2.0 ** -10), key=bytes(16),
                    name=WORKLOAD_NAMES["aes"],
                    kname="xmr_aes128_enc_nc2_inj1", protection="-DWC + on-device injector", bound="hbm", sets=2),
        "crc16": dict(kernel=cb.K_CRC16, nc=3, flags=F, n=1 << 20, unit_bytes=64, in_b=64, out_b=2, alg_b=66, plan=None,
                      name=WORKLOAD_NAMES["crc16"], kname="xmr_crc16_b64_nc3_inj0",
                      protection="-TMR -countErrors -countSyncs", bound="hbm", sets=4),
        "gemm": dict(kernel=cb.K_GEMM_TF32, nc=3, flags=F, side=4096, plan=None,
                     name=WORKLOAD_NAMES["gemm"],
                     kname="xmr_gemm_tf32_nc3_inj0", protection="-TMR -countErrors -countSyncs", bound="tensor", sets=2),
    }


def run_ours(args):
    import torch
    import coast_b200 as cb
    from coast_b200.shard import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rt = cb.Runtime(local)
    dev = f"cuda:{local}"
    W = workload_table(cb)[args.workload]
    is_gemm = W["kernel"] == cb.K_GEMM_TF32
    plan = cb.FaultPlan(mode=cb.PLAN_BERNOULLI, **W["plan"]) if W.get("plan") else None
    nsets = W["sets"]

    if is_gemm:
        # strong scaling: C row-blocks of side/world rows per GPU, A row-block local, B replicated
        side = W["side"]
        r0, r1 = shard_range(side // 128, rank, world)
        rows = (r1 - r0) * 128
        n = rows * side
        unit_base = r0 * 128 * side
        scaling = "strong"
        out_b, in_b = 4, 0
        alg_bytes = (rows * side + side * side + rows * side) * 4
        flops_issued = 3 * 2.0 * rows * side * side
        ins, outs, auxs = [], [], []
        for i in range(nsets):
            A = torch.empty(max(rows, 1) * side, dtype=torch.float32, device=dev)
            B = torch.empty(side * side, dtype=torch.float32, device=dev)
            rt.fill_philox(A, seed=4, word_base=unit_base + i * 0x20000000)
            rt.fill_philox(B, seed=44, word_base=i * 0x20000000)
            # Philox words -> uniform(-1,1) fp32
            A = (A.view(torch.int32).to(torch.float64) / 2 ** 31).to(torch.float32).contiguous()
            B = (B.view(torch.int32).to(torch.float64) / 2 ** 31).to(torch.float32).contiguous()
            ins.append(A); auxs.append(B); outs.append(torch.empty(max(n, 1), dtype=torch.float32, device=dev))
        descs = [rt.make_desc(W["kernel"], W["nc"], ins[i], outs[i], n, flags=W["flags"], M=rows, N=side, K=side, d_aux=auxs[i],
                              unit_base=unit_base, plan=plan) for i in range(nsets)] if n else []
        total_out_bytes = side * side * 4
    else:
        if W.get("strong"):                                  # config 5: a fixed 2^30-message batch, contiguous shards
            lo, hi = shard_range(W["n"], rank, world)
            n, unit_base, scaling = hi - lo, lo, "strong"
            if n * 96 > 150 * (1 << 30):
                raise SystemExit(f"{args.workload}: {n} messages per GPU do not fit 180 GB; use more GPUs")
        else:
            n = W["n"]                                       # weak scaling: the same shard size on every GPU
            unit_base = rank * n
            scaling = "weak"
        in_b, out_b = W["in_b"], W["out_b"]
        alg_bytes = n * W["alg_b"]
        ins = [torch.empty(n * in_b, dtype=torch.uint8, device=dev) for _ in range(nsets)]
        outs = [torch.empty(n * out_b, dtype=torch.uint8, device=dev) for _ in range(nsets)]
        for i, t in enumerate(ins):
            rt.fill_philox(t, seed=2, word_base=(unit_base * in_b // 4) + i * 0x10000000)
        descs = [rt.make_desc(W["kernel"], W["nc"], ins[i], outs[i], n, flags=W["flags"], unit_bytes=W["unit_bytes"], key=W.get("key"),
                              unit_base=unit_base, plan=plan) for i in range(nsets)]
        total_out_bytes = (W["n"] if W.get("strong") else world * n) * out_b
    NSTAT = 8                                              # counter-exchange buffers in flight
    d_stats = [torch.zeros(5, dtype=torch.int64, device=dev) for _ in range(NSTAT)]
    pending = [None] * NSTAT
    launches = 0

    def step(i):
        nonlocal launches
        if descs:
            rt.launch(descs[i % nsets])                    # ONE kernel: replicas + voter + counters (+ injector)
            launches += 1
        if dist is not None:
            # the only exchange step: 32 bytes of counters over NVLink.  Issued asynchronously on NCCL's stream (it waits for
            # the snapshot, the compute stream does not wait for it), so it overlaps the next step's kernel; every exchange
            # is completed inside the timed region (drain() before the stop event).
            k = i % NSTAT
            if pending[k] is not None:
                pending[k].wait()
            rt.stats_snapshot(d_stats[k])                  # D2D copy of the counters
            pending[k] = dist.all_reduce(d_stats[k][:4], async_op=True)

    def drain():
        for k in range(NSTAT):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def fence():
        drain()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    for i in range(args.warmup):
        step(i)
    fence()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for i in range(args.steps):
        step(i)
    drain()                                                # all counter exchanges complete before the clock stops
    e1.record()
    fence()
    sampler.stop()
    ms = e0.elapsed_time(e1)
    timed_launches = launches
    st = rt.sync()                                         # fold counters once (outside the timed region)
    if args.workload.startswith("sha256"):
        assert st.errors_corrected == 0 and st.syncs == 32 * n * (args.steps + args.warmup), st
    if args.workload == "aes":
        assert st.dwc_detected == st.injected > 0, st      # detect-rate parity: every state flip is detected

    # kernel-only duration for the roofline: per-launch events on the launching stream
    kms = []
    for i in range(max(3, min(args.steps, 20))):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if descs:
            rt.launch(descs[i % nsets])
        b.record()
        b.synchronize()
        kms.append(a.elapsed_time(b))
    rt.sync()
    k_ms = statistics.median(kms)

    # end to end through the reference-facing host call: pinned HOST buffers, H2D + kernel + D2H every step
    e2e_steps = max(3, min(args.steps, 20))
    if is_gemm:
        h_in = ins[0].cpu().pin_memory(); h_aux = auxs[0].cpu().pin_memory()
        h_out = torch.empty(max(n, 1), dtype=torch.float32).pin_memory()
        call = lambda: rt.run_host(W["kernel"], W["nc"], h_in, h_out, n, flags=W["flags"], M=rows, N=side, K=side, h_aux=h_aux,
                                   unit_base=unit_base, plan=plan)
        h2d, d2h = (rows * side + side * side) * 4, rows * side * 4 + 40
    else:
        ne = min(n, 1 << 24)                                 # e2e batch: at most 2^24 units of the shard through host memory
        h_in = torch.empty(ne * in_b, dtype=torch.uint8).pin_memory()
        h_in.copy_(ins[0][: ne * in_b].cpu())
        h_out = torch.empty(ne * out_b, dtype=torch.uint8).pin_memory()
        call = lambda: rt.run_host(W["kernel"], W["nc"], h_in, h_out, ne, unit_bytes=W["unit_bytes"], flags=W["flags"],
                                   key=W.get("key"), unit_base=unit_base, plan=plan)
        h2d, d2h = ne * in_b, ne * out_b + 40
    for _ in range(min(3, args.warmup)):
        call()
    fence()
    sampler.start(period=0.025)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.stop()
    assert torch.equal(h_out.view(torch.uint8), outs[0].view(torch.uint8)[: h_out.numel() * h_out.element_size()].cpu())
    e2e_units = (h_out.numel() * h_out.element_size()) // out_b      # units per e2e call on this rank

    # context for the e2e number: what a bare pinned copy of the same buffers achieves on this box
    def copy_gbs(dst, src, reps=5):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return src.numel() * src.element_size() * reps / (time.perf_counter() - t) / 1e9
    pcie_h2d = copy_gbs(ins[1 % nsets].view(torch.uint8)[: h_in.numel() * h_in.element_size()], h_in.view(torch.uint8))
    pcie_d2h = copy_gbs(h_out.view(torch.uint8), outs[0].view(torch.uint8)[: h_out.numel() * h_out.element_size()])

    # SURVEY.md 8e "measured separately": the OPTIONAL data-path collectives a caller may add around the sharded launch --
    # all-gather of the voted outputs; for the matmul also the one distribution step (B from rank 0).  Never part of
    # `value`: the path itself exchanges only the 5 counters.  Device-timed, max over ranks.
    coll_ms = [0.0, 0.0]
    coll_note = None
    if dist is not None and world > 1:
        def time_coll(fn, reps=3):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); a.record()
            for _ in range(reps):
                fn()
            b.record(); b.synchronize()
            return a.elapsed_time(b) / reps
        try:
            mine = outs[0].view(torch.uint8)
            per = torch.tensor([mine.numel()], dtype=torch.int64, device=dev)
            lo_hi = torch.stack([per, -per]).flatten()
            dist.all_reduce(lo_hi, op=dist.ReduceOp.MIN)
            equal = int(lo_hi[0]) == -int(lo_hi[1])                       # all_gather_into_tensor needs equal shards
            if equal and world * mine.numel() <= (2 << 30):
                gathered = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
                coll_ms[0] = time_coll(lambda: dist.all_gather_into_tensor(gathered, mine))
                assert torch.equal(gathered[rank * mine.numel():(rank + 1) * mine.numel()], mine)
                del gathered
            else:
                coll_note = "all-gather skipped: unequal shards or more than 2 GiB of outputs"
            if is_gemm:
                Bt = auxs[0]
                coll_ms[1] = time_coll(lambda: dist.broadcast(Bt, src=0))
        except Exception as exc:                                          # optional measurement: never lose the bench line
            coll_note = f"collective timing failed: {exc!r}"[:200]

    t = torch.tensor([ms, e2e_s, coll_ms[0], coll_ms[1]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max over ranks
    ms, e2e_s, coll_ms = float(t[0]), float(t[1]), [float(t[2]), float(t[3])]
    if rank == 0:
        ms_per_step = ms / args.steps
        value = total_out_bytes / (ms_per_step * 1e-3) / 1e6
        # the dominant kernel's average launch duration over the TIMED REGION (one launch per step, back to back on the
        # launching stream, CUDA events); the per-launch event pairs measured after it (k_ms) are kept as a cross-check
        k_reg = ms / max(1, timed_launches)
        if W["bound"] == "hbm":
            peak, peak_src = peaks()
            achieved, unit = alg_bytes / (k_reg * 1e-3) / 1e9, "GB/s"
            rl_extra = {"algorithmic_bytes_per_launch": alg_bytes,
                        "note": "integer-issue / shared-memory bound, not HBM bound: see DESIGN.md section 5"}
        else:
            pj = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            peak = float(pj.get("bf16_tflops", 1590.0)) / 2.0
            peak_src = "measured bf16 cuBLAS burst / 2 (TF32 runs at half the bf16 rate)" if pj else "fallback 1590/2"
            achieved, unit = flops_issued / (k_reg * 1e-3) / 1e12, "TFLOP/s"
            rl_extra = {"issued_flops_per_launch": flops_issued, "useful_flops_per_launch": flops_issued / 3,
                        "note": "issued = 3 replicas x 2MNK; useful = one replica"}
        traffic = None
        tp = os.path.join(ROOT, "profiles", f"r01_{args.workload}_traffic.json")
        if args.workload.startswith("sha256"):
            tp = os.path.join(ROOT, "profiles", "r01_sha256_tmr_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC if args.workload.startswith("sha256") else f"protected-kernel throughput (MB/s voted output), {args.workload}",
            "value": round(value, 1), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32(tf32 mma)" if is_gemm else ("u8" if args.workload == "aes" else "u32"), "data": "synthetic",
            "config": {"workload": W["name"], "units_per_gpu": n, "protection": W["protection"],
                       "voter": "select (r0==r1?r0:r2), one vote per stored element",
                       "layout": "-s: replicas on adjacent warps (sha256 TMR default); adjacent lanes otherwise; GEMM: 3 TMEM accumulators",
                       "l2": f"{nsets} rotating in/out buffer sets = {nsets * alg_bytes >> 20} MiB > 126 MB L2",
                       "parallelism": f"shard{world}" if world > 1 else "1gpu"},
            "roofline": dict({"bound": W["bound"], "achieved": round(achieved, 2), "peak": peak, "unit": unit,
                              "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                              "kernel": W["kname"], "kernel_ms": round(k_reg, 5),
                              "kernel_ms_source": "timed region / launches (CUDA events on the launching stream)",
                              "kernel_ms_single_launch_events": round(k_ms, 5)}, **rl_extra),
            "e2e": {"value": round(world * e2e_units * out_b / e2e_s / 1e6, 1), "unit": "MB/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_s * 1e3, 4), "timer": "host clock around the blocking C-ABI call coast_run_host",
                    "pcie_pinned_copy_gbs": {"h2d": round(pcie_h2d, 1), "d2h": round(pcie_d2h, 1)}},
            "gpu_launches": timed_launches,
            "clocks": sampler.summary(),
            "stats_last_sync": st.as_dict(),
        }
        if world > 1:
            line["collectives"] = {"in_value": "all-reduce of the 5 counters per step (inside the timed region)",
                                   "allgather_outputs_ms": round(coll_ms[0], 4) if coll_ms[0] else None,
                                   "allgather_bytes": world * outs[0].numel() * outs[0].element_size() if coll_ms[0] else None,
                                   "broadcast_B_ms": round(coll_ms[1], 4) if coll_ms[1] else None,
                                   "note": coll_note or "measured separately, not part of value (SURVEY.md 8e)"}
        if world == 1 and not args.no_cpu_baseline and args.workload == "sha256":
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=90.0, help="--impl reference: CPU seconds the whole run may take")
    ap.add_argument("--threads", type=int, default=0, help="--impl reference: host threads (default: all; config 1 is 1 thread)")
    ap.add_argument("--workload", choices=["sha256", "sha256_2p30", "aes", "crc16", "gemm"], default="sha256",
                    help="default sha256 = BASELINE configs[1], the headline line; the others are extra lines")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
