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


# ---------------------------------------------------------------- 8f-1: in-loop store votes
def test_store_votes_change_the_sync_count_not_the_result(oracle):
    """-storeDataSync / -noMemReplication (synchronization.cpp:205-215, syncStoreInst :476-560): every assignment to a data
    variable is voted -- crc16: 3 per byte + the exit, matrix_multiply: K + the exit -- and -noStoreDataSync removes them again"""
    import numpy as np
    n, L = 50, 13
    msg = oracle.fill_philox((n * L + 3) // 4, 0, 1).view(np.uint8)[: n * L].copy()
    base, st0 = oracle.run(oracle.K_CRC16, 3, msg, n, unit_bytes=L, flags=3)
    assert st0["syncs"] == n
    for extra in (oracle.F_STORE_DATA_SYNC, oracle.F_NO_MEM_REPLICATION, oracle.F_NO_MEM_REPLICATION | oracle.F_NO_LOAD_SYNC):
        out, st = oracle.run(oracle.K_CRC16, 3, msg, n, unit_bytes=L, flags=3 | extra)
        assert out.tobytes() == base.tobytes() and st["syncs"] == n * (3 * L + 1) and st["errors_corrected"] == 0
    out, st = oracle.run(oracle.K_CRC16, 3, msg, n, unit_bytes=L, flags=3 | oracle.F_NO_MEM_REPLICATION | oracle.F_NO_STORE_DATA_SYNC)
    assert st["syncs"] == n                                     # C4 switched off again: only the forced SoR-exit votes remain
    M = N = K = 9
    A, B = oracle.fill_philox(M * K, 0, 4), oracle.fill_philox(K * N, 0, 44)
    base, st0 = oracle.run(oracle.K_MM_U32, 3, A, M * N, flags=3, M=M, N=N, K=K, aux=B)
    out, st = oracle.run(oracle.K_MM_U32, 3, A, M * N, flags=3 | oracle.F_NO_MEM_REPLICATION, M=M, N=N, K=K, aux=B)
    assert out.tobytes() == base.tobytes() and st0["syncs"] == M * N and st["syncs"] == M * N * (K + 1)


def test_store_votes_correct_a_flip_at_the_next_assignment(oracle):
    """every site, one flipped replica: the output is always the fault-free one; TMR counts exactly one corrected vote per flip
    (the next vote on a value computed from the flipped copy); DWC flags the unit"""
    import numpy as np
    L = 7
    sites = list(range(2 * L))
    n = len(sites)
    msg = oracle.fill_philox((n * L + 3) // 4, 0, 1).view(np.uint8)[: n * L].copy()
    clean, _ = oracle.run(oracle.K_CRC16, 1, msg, n, unit_bytes=L)
    for rep in (0, 1, 2):
        table = np.array([oracle.fault_entry(rep, s, (s * 5) % (16 if s < L else 8)) for s in sites], dtype=np.uint32)
        out, st = oracle.run(oracle.K_CRC16, 3, msg, n, unit_bytes=L, flags=3 | oracle.F_STORE_DATA_SYNC, plan=oracle.make_plan(oracle.PLAN_TABLE, table=table))
        assert out.tobytes() == clean.tobytes() and st["errors_corrected"] == n and st["injected"] == n
        if rep < 2:
            out, st = oracle.run(oracle.K_CRC16, 2, msg, n, unit_bytes=L, flags=oracle.F_STORE_DATA_SYNC, plan=oracle.make_plan(oracle.PLAN_TABLE, table=table))
            assert st["dwc_detected"] == n


def test_sha256_store_votes_count_and_correct_every_site(oracle):
    """sha256 under -noMemReplication / -storeDataSync: len + 720 per compression + 32 votes per message; every one of the 536 sites of
    both compressions of a 64-byte message is corrected before the digest; a flipped working variable can be voted twice before it is
    overwritten, so errors_corrected >= injected"""
    import hashlib
    import numpy as np
    for L in (0, 1, 55, 56, 64, 100):
        n = 6
        m = oracle.fill_philox((n * L + 3) // 4 + 1, 0, 2).view(np.uint8)[: n * L].copy() if L else np.zeros(4, dtype=np.uint8)
        out, st = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=L, flags=3 | oracle.F_NO_MEM_REPLICATION)
        assert st["syncs"] == n * (L + 720 * ((L + 8) // 64 + 1) + 32) and st["errors_corrected"] == 0
        for u in range(n):
            assert out[32 * u: 32 * u + 32].tobytes() == hashlib.sha256(m[L * u: L * u + L].tobytes()).digest()
    L, n = 64, 1072
    m = oracle.fill_philox(n * L // 4, 0, 2).view(np.uint8)
    clean, _ = oracle.run(oracle.K_SHA256, 1, m, n, unit_bytes=L)
    tab = np.array([oracle.fault_entry(u % 3, u, (u * 7) % 32) for u in range(n)], dtype=np.uint32)
    out, st = oracle.run(oracle.K_SHA256, 3, m, n, unit_bytes=L, flags=3 | oracle.F_STORE_DATA_SYNC, plan=oracle.make_plan(oracle.PLAN_TABLE, table=tab))
    assert out.tobytes() == clean.tobytes() and st["injected"] == n and n <= st["errors_corrected"] <= 3 * n
