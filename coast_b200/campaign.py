"""On-device fault-injection campaigns, in the log format of the reference's campaign tooling.

The reference runs one QEMU+GDB session per injection (simulation/platform/supervisor.py:400-509): random
time, random location, one bit flip (resources/injector.py:202-207), then classifies the UART line
`C:.. E:.. F:.. T:..` (resources/decoder.py:66-86) and appends one InjectionLog dict per run to a JSON file
(resources/supportClasses.py:338-356; file layout supervisor.py:436, read back by jsonParser.py:120-145).

Here one *unit* of a protected launch is one "run": a Bernoulli(p=1) fault plan flips exactly one bit of one
live replica value per unit, the kernel reports per unit how many SoR-exit votes disagreed (d_status, the "F:"
field) and the voted output is compared with a fault-free launch (the "E:" field).  5 000 injections take one
kernel launch instead of hours; the JSON written here loads in the reference's jsonParser.py unchanged.

    python -m coast_b200.campaign --workload crc16 --passes="-TMR -countErrors" -t 5000 -o crc_tmr.json
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
from dataclasses import dataclass

import numpy as np

from . import runtime as R

WORKLOADS = {"crc16": R.K_CRC16, "sha256": R.K_SHA256, "aes": R.K_AES128, "mm": R.K_MM_U32, "qsort": R.K_QSORT,
             "chsha": R.K_CHSTONE_SHA, "chaes": R.K_CHSTONE_AES}


# --------------------------------------------------------------------------------------------------------
# host copy of the fault-plan arithmetic (include/coast_rt.h "Fault plan"): Philox4x32-10, vectorised
# --------------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        c0, c1, c2, c3 = n0 & MASK, p1 & MASK, n2 & MASK, p0 & MASK
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def plan_faults(kernel, num_clones, unit_bytes, K, n_units, seed, threshold, unit_base=0):
    """(active, replica, site, bit) arrays for units [unit_base, unit_base+n_units) -- what the kernel will do."""
    L = R.load_library()
    g = np.arange(unit_base, unit_base + n_units, dtype=np.uint64)
    z = np.zeros(n_units, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), z, z, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ns = int(L.coast_fault_sites(kernel, unit_bytes, K))
    active = x0 < np.uint32(threshold)
    replica = x1 % np.uint32(num_clones)
    site = x2 % np.uint32(ns)
    if kernel == R.K_CRC16:
        width = np.where(site < unit_bytes, 16, 8).astype(np.uint32)
    elif kernel in (R.K_AES128, R.K_CHSTONE_AES):
        width = np.full(n_units, 8, dtype=np.uint32)
    else:
        width = np.full(n_units, 32, dtype=np.uint32)
    return active, replica, site, x3 % width


def site_name(kernel, unit_bytes, site):
    """Human-readable name of an enumerated fault site -- the `name`/`address` of the log."""
    site = int(site)
    if kernel == R.K_CRC16:
        return f"crc16.crc@byte{site}" if site < unit_bytes else f"crc16.data[{site - unit_bytes}]"
    if kernel == R.K_SHA256:
        blk, s = divmod(site, 536)
        if s < 16:
            return f"sha256.blk{blk}.m[{s}]"
        if s < 528:
            return f"sha256.blk{blk}.round{(s - 16) // 8}.{'abcdefgh'[(s - 16) % 8]}"
        return f"sha256.blk{blk}.ctx_state[{s - 528}]"
    if kernel == R.K_AES128:
        return f"aes.state_in[{site}]" if site < 16 else f"aes.round{(site - 16) // 16}.state[{(site - 16) % 16}]"
    if kernel == R.K_CHSTONE_AES:
        return f"chaes.statemt_in[{site}]" if site < 16 else f"chaes.keyadd{(site - 16) // 16 + 1}.statemt[{(site - 16) % 16}]"
    if kernel == R.K_MM_U32:
        return f"mm.sum@k{site}"
    if kernel == R.K_QSORT:
        L = unit_bytes // 4
        return f"qsort.cmp_operand@event{site}" if site < 32 * L else f"qsort.array[{site - 32 * L}]"
    if kernel == R.K_CHSTONE_SHA:
        blk, s = divmod(site, 421)
        if s < 16:
            return f"chsha.blk{blk}.W[{s}]"
        if s < 416:
            return f"chsha.blk{blk}.round{(s - 16) // 5}.{'ABCDE'[(s - 16) % 5]}"
        return f"chsha.blk{blk}.sha_info_digest[{s - 416}]"
    return f"site{site}"


# --------------------------------------------------------------------------------------------------------
# log records (same keys as supportClasses.py InjectionLog.getDict / RunResult.getDict / AbortResult.getDict)
# --------------------------------------------------------------------------------------------------------
def _now():
    return datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f")


def injection_record(number, section, address, old, new, name, result, cycles=0):
    return {"timestamp": _now(), "number": int(number), "section": section, "oldValue": int(old), "newValue": int(new),
            "address": address, "sleepTime": 0, "cycles": int(cycles), "PC": 0, "name": name, "result": result,
            "cacheInfo": None}


def run_result(errors, faults, runtime_s):
    return {"timestamp": _now(), "core": 0, "runtime": float(runtime_s), "errors": int(errors), "faults": int(faults)}


def abort_result(message):
    return {"type": "Data", "message": message, "timestamp": _now(), "errors": 0}


def write_log(path, exec_path, records):
    """First line = path of the executable used (jsonParser.py:124-129 checks it exists), then the JSON list."""
    with open(path, "w", encoding="utf-8") as f:
        f.write(f"{exec_path}\n")
        json.dump(records, f, indent=1)
        f.write("\n")


@dataclass
class Summary:
    name: str
    injections: int
    success: int = 0
    errors: int = 0       # SDC: the voted output is wrong        (jsonParser.py:173-174)
    faults: int = 0       # TMR: corrected, output right, F > 0    (:175-176)
    detected: int = 0     # DWC: FAULT_DETECTED_DWC -> abort; the reference books these under timeouts/aborts (:166-169)

    def as_dict(self):
        n = max(1, self.injections)
        return {"name": self.name, "injections": self.injections, "success": self.success, "errors": self.errors,
                "faults": self.faults, "dwc_detected": self.detected,
                "coverage_pct": round(100.0 * (self.injections - self.errors) / n, 3)}

    def row(self):
        """`OK / Err / DWC-detected`, the cell format of docs/images/msp430/fault_injection_results2.png."""
        ok = self.success + self.faults
        return f"{ok} / {self.errors} / {self.detected if self.detected else '-'}"


def run_campaign(rt, workload: str, opt_passes: str, n_injections: int, seed: int = 1, *, unit_bytes: int | None = None,
                 data_seed: int = 1, log_path: str | None = None):
    """One launch = n_injections single-bit-flip runs.  Returns (Summary, records)."""
    import torch
    kernel = WORKLOADS[workload]
    nc, flags = R.parse_opt_passes(opt_passes)
    flags |= R.F_COUNT_ERRORS if nc == 3 else 0
    dev = f"cuda:{rt.device}"
    kw, K = {}, 0
    n = n_injections
    if kernel == R.K_MM_U32:
        side = 1
        while side * side < n_injections:
            side += 1
        side = max(side, 9)
        n, K = side * side, side
        A = torch.empty(n, dtype=torch.int32, device=dev)
        B = torch.empty(n, dtype=torch.int32, device=dev)
        rt.fill_philox(A, data_seed)
        rt.fill_philox(B, data_seed + 40)
        inp, kw = A, dict(M=side, N=side, K=side, aux=B)
        ub = 0
    else:
        ub = {R.K_CRC16: 64, R.K_SHA256: 64, R.K_AES128: 16, R.K_QSORT: 4 * 580, R.K_CHSTONE_SHA: 1024, R.K_CHSTONE_AES: 64}[kernel] if unit_bytes is None else unit_bytes
        nbytes = (n * ub + 3) // 4 * 4
        inp = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rt.fill_philox(inp, data_seed)
        if kernel == R.K_CHSTONE_AES:                          # one byte per int (aes.c:83)
            inp = (inp.view(torch.int32) & 0xFF).view(torch.uint8)
            kw = dict(key=bytes(range(16)))
        elif kernel == R.K_AES128:
            kw = dict(key=bytes(16))
        else:
            kw = dict(unit_bytes=ub)
    golden, _ = rt.run(kernel, 1, inp, n, **kw)
    status = torch.zeros(n, dtype=torch.uint8, device=dev)
    plan = R.FaultPlan(mode=R.PLAN_BERNOULLI, seed=seed, threshold=0xFFFFFFFF)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    out, st = rt.run(kernel, nc, inp, n, flags=flags, plan=plan, status=status, **kw)
    t1.record()
    t1.synchronize()
    per_run_s = t0.elapsed_time(t1) * 1e-3 / n
    ob = R.out_bytes(kernel, ub)
    wrong = (out.view(n, ob) != golden.view(n, ob)).any(dim=1).cpu().numpy()[:n_injections]
    stat = status.cpu().numpy()[:n_injections]
    active, replica, site, bit = plan_faults(kernel, nc, ub, K, n_injections, seed, 0xFFFFFFFF)
    summ = Summary(f"{workload} [{opt_passes.strip() or 'unmitigated'}]", n_injections)
    records = []
    for u in range(n_injections):
        name = site_name(kernel, ub, site[u])
        section = "memory" if (".data[" in name or ".m[" in name or ".W[" in name or "state_in" in name or "statemt_in" in name or ".array[" in name) else "registers"
        if nc == 2 and stat[u]:
            res = abort_result("FAULT_DETECTED_DWC")
            summ.detected += 1
        else:
            res = run_result(int(wrong[u]), int(stat[u]) if nc == 3 else 0, per_run_s)
            if wrong[u]:
                summ.errors += 1
            elif nc == 3 and stat[u]:
                summ.faults += 1
            else:
                summ.success += 1
        records.append(injection_record(u, section, f"replica{int(replica[u])}:{name}", 0, 1 << int(bit[u]), name, res,
                                        cycles=int(site[u])))
    if log_path:
        write_log(log_path, R.lib_path(), records)
    # cross-check against the device counters of the same launch
    if nc == 3 and n == n_injections and int(stat.max(initial=0)) < 255:
        assert int(stat.astype(np.int64).sum()) == st.errors_corrected, (int(stat.sum()), st.errors_corrected)
    if nc == 2 and n == n_injections:
        assert summ.detected == st.dwc_detected, (summ.detected, st.dwc_detected)
    return summ, records


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="crc16")
    ap.add_argument("--passes", default="-TMR -countErrors", help="OPT_PASSES string, e.g. '', '-DWC', '-TMR -countErrors'")
    ap.add_argument("-t", type=int, default=5000, help="number of injections (supervisor.py -t)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("-o", default=None, help="JSON log path (jsonParser.py compatible)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    rt = R.Runtime(a.device)
    summ, _ = run_campaign(rt, a.workload, a.passes, a.t, a.seed, log_path=a.o)
    print(json.dumps(summ.as_dict()))
    print("OK / Err / DWC-detected:", summ.row())


if __name__ == "__main__":
    main()
