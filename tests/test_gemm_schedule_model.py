"""CPU: a model of the TF32 GEMM tile schedules (xmr_gemm_tf32.cuh `tile_coords` + `decode`, xmr_gemm_tf32_pair.cuh) -- every C element
is produced exactly once for any shape and grid, the halved tiles of a short last round included, and the three warp roles of a CTA
(producer, MMA issuer, epilogue) walk the same sequence.  The kernels' results on the device are pinned in tests/test_gpu_gemm.py;
this covers the shapes and SM counts a single box cannot."""
import itertools

import pytest


def tile_coords(tile, tiles_m, tiles_n, group_m):
    per_group = group_m * tiles_n
    g, w = divmod(tile, per_group)
    rows = min(group_m, tiles_m - g * group_m)
    return g * group_m + w % rows, w // rows


def schedule(M, N, bm, bn, workers, group_m, split):
    """-> {worker: [(m0, n0, rows, cols), ...]} as the kernels' `for (tile = id; tile < n_virtual; tile += workers)` loops produce it"""
    tiles_m, tiles_n = M // bm, N // bn
    n_tiles = tiles_m * tiles_n
    sched_full = n_virtual = n_tiles
    if split:
        whole = (n_tiles // workers) * workers
        rem = n_tiles - whole
        if rem and 2 * rem <= workers:
            sched_full, n_virtual = whole, whole + 2 * rem
    out = {}
    for w in range(workers):
        seq = []
        for v in range(w, n_virtual, workers):
            t, h, cols = v, 0, bn
            if v >= sched_full:
                t, h, cols = sched_full + (v - sched_full) // 2, (v - sched_full) & 1, bn // 2
            tm, tn = tile_coords(t, tiles_m, tiles_n, group_m)
            seq.append((tm * bm, tn * bn + h * (bn // 2), bm, cols))
        out[w] = seq
    return out


def covered_once(M, N, sched):
    seen = {}
    for seq in sched.values():
        for m0, n0, rows, cols in seq:
            assert m0 + rows <= M and n0 + cols <= N and cols % 64 == 0
            for bi in range(m0 // 128, (m0 + rows) // 128):
                for bj in range(n0 // 64, (n0 + cols) // 64):
                    assert (bi, bj) not in seen, (bi, bj)
                    seen[(bi, bj)] = 1
    assert len(seen) == (M // 128) * (N // 64)


SHAPES = [(128, 256), (256, 256), (512, 768), (2432, 2048), (2560, 4096), (4096, 4096), (1024, 8192), (8192, 512), (3840, 1280)]


@pytest.mark.parametrize("M,N", SHAPES)
@pytest.mark.parametrize("sms", [148, 132, 16, 3])
@pytest.mark.parametrize("group_m", [16, 8, 1, 32])
def test_single_cta_schedules_cover_c_exactly_once(M, N, sms, group_m):
    for bn, split in ((256, True), (256, False), (128, False)):      # unprotected wide (with / without the tail split), protected / narrow
        if N % bn:
            continue
        tiles = (M // 128) * (N // bn)
        grid = min(tiles, sms)
        covered_once(M, N, schedule(M, N, 128, bn, grid, group_m, split))


@pytest.mark.parametrize("M,N", [s for s in SHAPES if s[0] % 256 == 0])
@pytest.mark.parametrize("sms", [148, 132, 16, 3])
@pytest.mark.parametrize("group_m", [16, 8, 2])
def test_cta_pair_schedules_cover_c_exactly_once(M, N, sms, group_m):
    for bn, split in ((256, True), (128, False)):                    # unprotected pairs (tail split), DWC / TMR pairs
        if N % bn:
            continue
        grid = min((M // 128) * (N // bn), sms) & ~1
        if grid < 2:
            continue
        pairs = grid // 2
        covered_once(M, N, schedule(M, N, 256, bn, pairs, max(1, group_m // 2), split))


def test_the_tail_split_is_taken_exactly_when_it_saves_time():
    # 4096^2 unprotected on 148 CTAs: 512 tiles = 3 whole rounds + 68 <= 74 -> 136 half tiles in ONE more (half-length) round
    s = schedule(4096, 4096, 128, 256, 148, 16, True)
    lens = sorted({len(v) for v in s.values()})
    assert lens == [3, 4] and sum(1 for v in s.values() if len(v) == 4) == 136
    assert all(seq[-1][3] == 128 for seq in s.values() if len(seq) == 4) and all(t[3] == 256 for seq in s.values() for t in seq[:3])
    # 75 left-over tiles > 74: two half rounds would cost what one whole round costs -> whole tiles stay
    s = schedule(128 * 223, 256, 128, 256, 148, 16, True)           # 223 tiles = 148 + 75
    assert all(t[3] == 256 for seq in s.values() for t in seq)
    # consecutive halves of one tile go to neighbouring CTAs (they share the A row block in L2)
    s = schedule(2432, 2048, 128, 256, 148, 16, True)               # 152 tiles: 4 left over -> 8 halves on CTAs 0..7
    tails = [s[w][-1] for w in range(8)]
    assert all(tails[2 * i][0] == tails[2 * i + 1][0] and tails[2 * i][1] + 128 == tails[2 * i + 1][1] for i in range(4))
