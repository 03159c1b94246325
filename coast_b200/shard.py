"""Multi-GPU host logic: the protected workloads shard by independent units.

One process per GPU; rank r owns the contiguous GLOBAL unit range shard_range(n, r, world) and
passes its start as ``unit_base`` so the Philox fault plan -- keyed by the global unit index --
is identical to the single-GPU run.  No data-path collective exists: the only exchange is the
reduction of the five counters (SUM for the four counts, MIN for first_fault_unit).
"""
from __future__ import annotations

NO_FAULT_UNIT = 0xFFFFFFFFFFFFFFFF
_I64_MAX = 0x7FFFFFFFFFFFFFFF


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced: the first (n % world) ranks get one extra unit."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stats_to_tensor(stats: dict, torch, device="cpu"):
    f = stats["first_fault_unit"]
    return torch.tensor([stats["errors_corrected"], stats["dwc_detected"], stats["syncs"], stats["injected"],
                         _I64_MAX if f == NO_FAULT_UNIT else f], dtype=torch.int64, device=device)


def tensor_to_stats(t) -> dict:
    v = [int(x) for x in t.tolist()]
    return dict(errors_corrected=v[0], dwc_detected=v[1], syncs=v[2], injected=v[3],
                first_fault_unit=NO_FAULT_UNIT if v[4] == _I64_MAX else v[4])


def allreduce_stats(t, dist, group=None):
    """In place: SUM over the four counters, MIN over first_fault_unit (works for nccl and gloo)."""
    dist.all_reduce(t[:4], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t[4:], op=dist.ReduceOp.MIN, group=group)
    return t


def device_counters_to_stats_tensor(raw):
    """raw: int64 view of the 5 u64 device counters (coast_stats_snapshot); maps ~0 -> INT64_MAX for MIN."""
    out = raw.clone()
    out[4] = _I64_MAX if int(raw[4]) == -1 else raw[4]
    return out
