"""CPU, world_size 2 over gloo: the N>1 host logic (coast_b200/shard.py).  Each rank runs its shard with
unit_base = shard start; counters are all-reduced (SUM / MIN).  The compute stand-in is the oracle -- the
point is the sharding arithmetic and the collective, which are identical under nccl."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from coast_b200.shard import shard_range
    for n in (0, 1, 7, 10, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from coast_b200.shard import allreduce_stats, shard_range, stats_to_tensor, tensor_to_stats
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    msgs = po.fill_philox(n * 16, 0, 2).view(np.uint8)
    plan = po.make_plan(po.PLAN_BERNOULLI, seed=5, p=0.2)
    out, st = po.run(po.K_SHA256, 3, msgs[64 * lo: 64 * hi], hi - lo, unit_bytes=64, flags=3, plan=plan, unit_base=lo)
    t = allreduce_stats(stats_to_tensor(st, torch), dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, out.tobytes())
    if rank == 0:
        q.put((tensor_to_stats(t), b"".join(gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_equals_single_rank():
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    n = 1001
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    msgs = po.fill_philox(n * 16, 0, 2).view(np.uint8)
    ref_out, ref_st = po.run(po.K_SHA256, 3, msgs, n, unit_bytes=64, flags=3, plan=po.make_plan(po.PLAN_BERNOULLI, seed=5, p=0.2))
    assert out == ref_out.tobytes()
    assert stats == ref_st


def _negotiate_worker(rank, world, port, scenario, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from coast_b200.shard import negotiate_peer_counter_block
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def export():
        calls.append("export")
        if scenario == "export_fails":
            raise RuntimeError("cuIpcGetMemHandle: not supported")
        return bytes(range(64))

    def probe(handle):
        calls.append(("probe", handle))
        if scenario == "no_peer_access" and rank == 1:
            raise RuntimeError("cuIpcOpenMemHandle: peer access unsupported")

    got = negotiate_peer_counter_block(dist, rank, torch, "cpu", export, probe)
    q.put((rank, got, calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scenario", ["ok", "export_fails", "no_peer_access"])
def test_peer_counter_block_negotiation_is_all_or_nothing(scenario):
    """bench.py's N>1 set-up of the NVLink counter fold (coast_b200/shard.py negotiate_peer_counter_block): every rank gets the same
    answer -- the owner's handle, or None when ANY rank cannot export / map it -- and nobody hangs in a collective"""
    import torch.multiprocessing as mp
    world = 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_negotiate_worker, args=(r, world, port, scenario, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict((r, (got, calls)) for r, got, calls in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes(range(64)) if scenario == "ok" else None
    assert all(res[r][0] == want for r in range(world)), res
    assert res[0][1] == ["export"]                                       # the owner never maps its own block
    if scenario == "export_fails":
        assert res[1][1] == [] and res[2][1] == []                       # nothing to probe
    else:
        assert res[1][1] == [("probe", bytes(range(64)))] == res[2][1]
