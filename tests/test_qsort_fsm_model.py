"""CPU model of the per-unit state machine of coast_b200/csrc/xmr_qsort.cuh (SCAN_I / SCAN_J / POP, left part entered
directly, right part on the explicit stack) checked against the oracle's nested-loop formulation of
tests/quicksort/quicksort.c:121-136: same sorted output, same number of sync points, same error count, same fault-event
numbering -- the scheduling is a rewrite of the control flow, not of the algorithm."""
import numpy as np


def fsm_unit(arr, nc, fault=None, count_errors=True):
    L = len(arr)
    rep = [list(map(int, arr)) for _ in range(nc)]
    fr, fsite, fmask = (fault[0], fault[1], 1 << fault[2]) if fault else (-1, -1, 0)
    if fsite >= 32 * L:
        rep[fr][fsite - 32 * L] = _i32(rep[fr][fsite - 32 * L] ^ fmask)
    syncs = errors = ev = 0
    stack = [(0, L)]
    phase, off, ln, pivot, i, j = "POP", 0, 0, [0] * nc, 0, 0

    def vote(c):
        nonlocal errors
        if nc == 1:
            return c[0]
        if nc == 2:
            errors += c[0] != c[1]
            return c[0]
        c01, c02 = c[0] == c[1], c[0] == c[2]
        errors += not (c01 and c02)
        return c[0] if c01 else c[2]

    def enter(o, n):
        nonlocal off, ln, pivot, i, j, phase, syncs
        off, ln = o, n
        syncs += 1                                           # if (len < 2) return;   :122
        if n < 2:
            phase = "POP"
        else:
            pivot = [rep[r][o + n // 2] for r in range(nc)]
            i, j, phase = 0, n - 1, "I"

    while True:
        if phase == "POP":
            if not stack:
                break
            enter(*stack.pop())
            continue
        idx = i if phase == "I" else j
        c = []
        for r in range(nc):
            v = rep[r][off + idx]
            if r == fr and fsite == ev:
                v = _i32(v ^ fmask)
            cr = (v < pivot[r]) if phase == "I" else (v > pivot[r])
            if (phase == "I" and i >= ln - 1) or (phase == "J" and j <= 0):
                cr = False
            c.append(cr)
        ev += 1
        syncs += 1
        voted = vote(c)
        if phase == "I":
            if voted:
                i += 1
            else:
                phase = "J"
        else:
            if voted:
                j -= 1
            else:
                syncs += 1                                   # if (i >= j) break;   :128
                if i < j:
                    for r in range(nc):
                        rep[r][off + i], rep[r][off + j] = rep[r][off + j], rep[r][off + i]
                    i, j, phase = i + 1, j - 1, "I"
                else:
                    i = min(max(i, 1), ln - 1)
                    stack.append((off + i, ln - i))          # :135, later
                    enter(off, i)                            # :134, now
    out = []
    for e in range(L):                                       # SoR exit votes
        vals = [rep[r][e] for r in range(nc)]
        if nc == 3:
            c01, c02 = vals[0] == vals[1], vals[0] == vals[2]
            errors += not (c01 and c02)
            out.append(vals[0] if c01 else vals[2])
        else:
            if nc == 2:
                errors += vals[0] != vals[1]
            out.append(vals[0])
    return out, syncs + L, errors


def _i32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def test_state_machine_equals_the_nested_loops(oracle):
    rng = np.random.default_rng(11)
    for L in (1, 2, 3, 7, 33, 100):
        for trial in range(6):
            arr = rng.integers(-50, 50, L, dtype=np.int32) if trial % 2 else rng.integers(-2 ** 31, 2 ** 31 - 1, L, dtype=np.int32)
            for nc in (1, 2, 3):
                ns = oracle.fault_sites(oracle.K_QSORT, 4 * L)
                faults = [None] + [(int(rng.integers(0, nc)), int(rng.integers(0, ns)), int(rng.integers(0, 32))) for _ in range(4)]
                for f in faults:
                    plan = None
                    if f is not None:
                        plan = oracle.make_plan(oracle.PLAN_TABLE, table=np.array([oracle.fault_entry(*f)], dtype=np.uint32))
                    want, st = oracle.run(oracle.K_QSORT, nc, arr, 1, unit_bytes=4 * L, plan=plan,
                                          flags=oracle.F_COUNT_ERRORS | oracle.F_COUNT_SYNCS)
                    out, syncs, errors = fsm_unit(arr, nc, f)
                    assert out == [int(v) for v in want.view(np.int32)], (L, nc, f)
                    if nc == 3:
                        assert syncs == st["syncs"] and errors == st["errors_corrected"], (L, nc, f, syncs, errors, st)
                    if nc == 2:
                        assert (errors > 0) == (st["dwc_detected"] == 1), (L, nc, f)
