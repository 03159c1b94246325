"""CPU: the protected-region semantics the oracle restates from
projects/dataflowProtection/synchronization.cpp."""
import hashlib

import numpy as np


def _msgs(oracle, n, nbytes, seed):
    return oracle.fill_philox(n * nbytes // 4, 0, seed).view(np.uint8)


def test_zero_fault_equals_unprotected(oracle):
    m = _msgs(oracle, 50, 64, 2)
    outs = [oracle.run(oracle.K_SHA256, nc, m, 50, unit_bytes=64, flags=3)[0] for nc in (1, 2, 3)]
    assert (outs[0] == outs[1]).all() and (outs[0] == outs[2]).all()
    for u in range(50):
        assert outs[0][32 * u: 32 * u + 32].tobytes() == hashlib.sha256(m[64 * u: 64 * u + 64].tobytes()).digest()


def test_tmr_single_fault_is_corrected_and_counted_per_byte(oracle):
    n = 200
    m = _msgs(oracle, n, 64, 3)
    clean, _ = oracle.run(oracle.K_SHA256, 1, m, n, unit_bytes=64)
    plan = oracle.make_plan(oracle.PLAN_BERNOULLI, seed=99, p=0.5)
    out, st = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=64, plan=plan, flags=oracle.F_COUNT_ERRORS | oracle.F_COUNT_SYNCS)
    assert (out == clean).all()                    # one faulty replica is always out-voted
    assert 60 < st["injected"] < 140
    assert st["syncs"] == 32 * n                   # 32 u8 votes per message (sha256_common_tmr.c:169-178)
    # count = number of digest BYTES in which the faulty replica differs (one vote per stored u8)
    expect = 0
    for u in range(n):
        f = oracle.fault_for_unit(plan, oracle.K_SHA256, 3, 64, 0, u)
        if f is None:
            continue
        tab = np.zeros(1, dtype=np.uint32)
        tab[0] = oracle.fault_entry(0, f[1], f[2])
        bad, _ = oracle.run(oracle.K_SHA256, 1, m[64 * u: 64 * u + 64], 1, unit_bytes=64,
                            plan=oracle.make_plan(oracle.PLAN_TABLE, table=tab))
        expect += int((bad != clean[32 * u: 32 * u + 32]).sum())
    assert st["errors_corrected"] == expect
    # without -countErrors nothing is counted, the vote still happens
    out2, st2 = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=64, plan=plan)
    assert (out2 == clean).all() and st2["errors_corrected"] == 0 and st2["syncs"] == 0


def test_select_voter_is_not_a_majority(oracle):
    """vote = (r0==r1) ? r0 : r2 (synchronization.cpp:512-522): a fault in r2 is ignored, a fault in r0
    or r1 selects r2 -- and a fault in r1 ALONE already makes the voter take r2."""
    n = 3
    m = _msgs(oracle, n, 13, 5)
    clean, _ = oracle.run(oracle.K_CRC16, 1, m, n, unit_bytes=13)
    for rep in range(3):
        tab = np.array([oracle.fault_entry(rep, 4, 7)] * n, dtype=np.uint32)
        out, st = oracle.run(oracle.K_CRC16, 3, m, n, unit_bytes=13, flags=1, plan=oracle.make_plan(oracle.PLAN_TABLE, table=tab))
        assert (out == clean).all() and st["errors_corrected"] == n and st["first_fault_unit"] == 0


def test_dwc_detects_and_keeps_r0(oracle):
    n = 64
    blocks = _msgs(oracle, n, 16, 3)
    key = bytes(16)
    clean, _ = oracle.run(oracle.K_AES128, 1, blocks, n, key=key)
    plan = oracle.make_plan(oracle.PLAN_BERNOULLI, seed=33, p=0.25)
    out, st = oracle.run(oracle.K_AES128, 2, blocks, n, key=key, plan=plan)
    # AES rounds are bijections: every state flip reaches the output -> detected == injected
    assert st["dwc_detected"] == st["injected"] > 0
    for u in range(n):
        f = oracle.fault_for_unit(plan, oracle.K_AES128, 2, 0, 0, u)
        same = (out[16 * u: 16 * u + 16] == clean[16 * u: 16 * u + 16]).all()
        assert same == (f is None or f[0] == 1)            # r0's (possibly faulty) value is what gets stored


def test_unit_base_makes_sharding_invisible(oracle):
    n = 96
    m = _msgs(oracle, n, 64, 2)
    plan = oracle.make_plan(oracle.PLAN_BERNOULLI, seed=5, p=0.3)
    whole, st = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=64, plan=plan, flags=3)
    parts, tot = [], dict(errors_corrected=0, syncs=0, injected=0)
    for r in range(4):
        lo, hi = r * 24, r * 24 + 24
        o, s = oracle.run(oracle.K_SHA256, 3, m[64 * lo: 64 * hi], 24, unit_bytes=64, plan=plan, flags=3, unit_base=lo)
        parts.append(o)
        for k in tot:
            tot[k] += s[k]
    assert (np.concatenate(parts) == whole).all()
    assert all(tot[k] == st[k] for k in tot)


def test_fault_site_geometry(oracle):
    assert oracle.fault_sites(oracle.K_SHA256, 64) == 2 * 536
    assert oracle.fault_sites(oracle.K_SHA256, 10) == 536
    assert oracle.fault_sites(oracle.K_SHA256, 55) == 536 and oracle.fault_sites(oracle.K_SHA256, 56) == 1072
    assert oracle.fault_sites(oracle.K_AES128) == 176
    assert oracle.fault_sites(oracle.K_CRC16, 13) == 26
    assert oracle.fault_site_bits(oracle.K_CRC16, 13, 0, 12) == 16 and oracle.fault_site_bits(oracle.K_CRC16, 13, 0, 13) == 8
    assert oracle.fault_sites(oracle.K_MM_U32, 0, 9) == 9


def test_mt_matches_single_thread(oracle):
    n = 1000
    m = _msgs(oracle, n, 64, 2)
    plan = oracle.make_plan(oracle.PLAN_BERNOULLI, seed=1, p=0.1)
    a, sa = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=64, plan=plan, flags=3)
    b, sb = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=64, plan=plan, flags=3, threads=4)
    assert (a == b).all() and sa == sb


def test_quicksort_branch_sync_points(oracle):
    """quick_sort (tests/quicksort/quicksort.c:121-136): the data-dependent branch conditions are the sync points"""
    L, n = 580, 40                                           # array_elements = 580 (quicksort.c:83)
    a = oracle.fill_philox(n * L, 0, 9).view(np.int32)
    want = np.sort(a.reshape(n, L), axis=1)
    for nc in (1, 2, 3):
        out, st = oracle.run(oracle.K_QSORT, nc, a, n, unit_bytes=4 * L, flags=3)
        assert (out.view(np.int32).reshape(n, L) == want).all()
        assert st["errors_corrected"] == 0 and st["dwc_detected"] == 0
    assert st["syncs"] > n * L * 8                            # many more sync points than SoR-exit stores
    plan = oracle.make_plan(oracle.PLAN_BERNOULLI, seed=3, threshold=0xFFFFFFFF)
    out3, st3 = oracle.run(oracle.K_QSORT, 3, a, n, unit_bytes=4 * L, flags=3, plan=plan)
    assert (out3.view(np.int32).reshape(n, L) == want).all()  # every voted branch follows the majority -> always sorted
    out2, st2 = oracle.run(oracle.K_QSORT, 2, a, n, unit_bytes=4 * L, plan=plan)
    out1, _ = oracle.run(oracle.K_QSORT, 1, a, n, unit_bytes=4 * L, plan=plan)
    wrong1 = (out1.view(np.int32).reshape(n, L) != want).any(axis=1).sum()
    assert st3["injected"] == n and 0 < st3["errors_corrected"] and st2["dwc_detected"] > 0
    assert wrong1 <= st2["dwc_detected"] + 2                  # what corrupts the unprotected sort is what DWC flags (replica choice aside)
    # an input-copy flip (site >= 32L) in one TMR replica: exactly one element vote disagrees at the SoR exit,
    # plus the branch votes that element took part in
    tab = np.zeros(n, dtype=np.uint32)
    tab[5] = oracle.fault_entry(1, 32 * L + 17, 30)
    o, s = oracle.run(oracle.K_QSORT, 3, a, n, unit_bytes=4 * L, flags=3, plan=oracle.make_plan(oracle.PLAN_TABLE, table=tab))
    assert (o.view(np.int32).reshape(n, L) == want).all() and s["injected"] == 1 and s["errors_corrected"] >= 1 and s["first_fault_unit"] == 5
