#!/usr/bin/env python3
"""Extended pin of the CPU oracle (oracle/coast_oracle.c) against the reference's own code compiled in place
(oracle/_ref): a large randomized differential run, beyond the committed golden fixture.  CPU only; needs the
reference checkout (or a prebuilt oracle/_ref).  Writes profiles/r01_oracle_pin_extended.md."""
import ctypes as C
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402


def main():
    po.build()
    assert po.ref_available(), "oracle/_ref is not built"
    rng = np.random.default_rng(2026)
    rows = []
    t0 = time.time()

    # crc16: every length 1..255, 40 random messages each
    rc = po.ref("crc16"); rc.ref_crc16.restype = C.c_ushort; rc.ref_crc16.argtypes = [C.c_char_p, C.c_ubyte]
    n = 0
    for ln in range(1, 256):
        for _ in range(40):
            m = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
            assert po.crc16(m) == rc.ref_crc16(m, ln), ln
            n += 1
    rows.append(("crc16()", "every length 1..255 x 40 random messages", n))

    # sha256: random lengths 0..1000 (every padding branch), also against hashlib
    rs = po.ref("sha256"); rs.ref_sha256.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p]
    n = 0
    for _ in range(20000):
        ln = int(rng.integers(0, 1001))
        m = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
        out = np.zeros(32, dtype=np.uint8)
        rs.ref_sha256(m if ln else b"\0", ln, out.ctypes.data)
        d = po.sha256(m)
        assert d == out.tobytes() == hashlib.sha256(m).digest(), ln
        n += 1
    rows.append(("sha256_hash()", "random lengths 0..1000, also == hashlib.sha256", n))

    # aes: random state/key, both directions, state AND the mutated key[]
    ra = po.ref("aes"); ra.ref_aes_enc_dec.argtypes = [C.c_void_p, C.c_void_p, C.c_ubyte]
    n = 0
    for _ in range(50000):
        st = rng.integers(0, 256, 16, dtype=np.uint8); key = rng.integers(0, 256, 16, dtype=np.uint8)
        for direction in (0, 1):
            s2, k2 = st.copy(), key.copy()
            ra.ref_aes_enc_dec(s2.ctypes.data, k2.ctypes.data, direction)
            so, ko = po.aes128(st.tobytes(), key.tobytes(), direction)
            assert so == s2.tobytes() and ko == k2.tobytes()
            n += 1
    rows.append(("aes_enc_dec()", "random state and key, both directions, state and mutated key[]", n))

    # chstone sha: random whole-block streams
    rh = po.ref("chsha"); rh.ref_chsha.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    n = 0
    for _ in range(5000):
        ln = 64 * int(rng.integers(1, 65))
        m = rng.integers(0, 256, ln, dtype=np.uint8)
        dg = np.zeros(5, dtype=np.uint32)
        rh.ref_chsha(m.ctypes.data, ln, dg.ctypes.data)
        assert po.chstone_sha(m.tobytes()) == [int(x) for x in dg], ln
        n += 1
    rows.append(("chstone sha_init/update/final", "random streams of 1..64 blocks", n))

    # mm: the 9x9 u32 kernel on random operands (mod 2^32)
    rm = po.ref("mm"); rm.ref_mm_multiply.argtypes = [C.c_void_p] * 3
    side = int(rm.ref_mm_side())
    n = 0
    for _ in range(2000):
        A = rng.integers(0, 2 ** 32, side * side, dtype=np.uint32); B = rng.integers(0, 2 ** 32, side * side, dtype=np.uint32)
        R = np.zeros(side * side, dtype=np.uint32)
        rm.ref_mm_multiply(A.ctypes.data, B.ctypes.data, R.ctypes.data)
        out, _ = po.run(po.K_MM_U32, 1, A, side * side, M=side, N=side, K=side, aux=B)
        assert out.tobytes() == R.tobytes()
        n += 1
    rows.append(("matrix_multiply() (mm_common_tmr.c)", f"random {side}x{side} uint32 operands", n))

    # quicksort: random arrays of random lengths (duplicates included)
    rq = po.ref("qsort")
    n = 0
    for _ in range(3000):
        ln = int(rng.integers(1, 1025))
        a = rng.integers(-2 ** 31, 2 ** 31 - 1, ln, dtype=np.int64).astype(np.int32)
        if n % 3 == 0:
            a = (a % 11).astype(np.int32)
        b = a.copy()
        rq.ref_quick_sort(b.ctypes.data, ln)
        out, _ = po.run(po.K_QSORT, 3, a, 1, unit_bytes=4 * ln)
        assert out.tobytes() == b.tobytes(), ln
        n += 1
    rows.append(("quick_sort() (quicksort.c)", "random arrays of 1..1024 ints, a third with many duplicates, under the TMR wrapper", n))

    # chstone aes: random blocks and keys through the reference's encrypt()/decrypt() (one byte per int, key never modified)
    rc_ = po.ref("chaes"); rc_.ref_chaes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    n = 0
    for _ in range(20000):
        blk = rng.integers(0, 256, 16, dtype=np.int64).astype(np.int32); key = rng.integers(0, 256, 16, dtype=np.int64).astype(np.int32)
        for direction in (0, 1):
            a_ = np.zeros(32, dtype=np.int32); k_ = np.zeros(32, dtype=np.int32)
            a_[:16] = blk; k_[:16] = key
            rc_.ref_chaes(a_.ctypes.data, k_.ctypes.data, direction)
            out, _ = po.run(po.K_CHSTONE_AES, 1, blk, 1, mode=2 | direction, aux=key)
            assert list(out.view(np.int32)) == [int(v) for v in a_[:16]] and (k_[:16] == key).all()
            n += 1
    rows.append(("chstone encrypt()/decrypt() (aes_enc.c, aes_dec.c)", "random block and key, both directions", n))

    dt = time.time() - t0
    lines = ["# Extended oracle pin (r02) -- `python tools/oracle_pin_extended.py`", "",
             "Randomized differential run of `oracle/coast_oracle.c` against the reference's own functions compiled in place",
             f"(`oracle/_ref`, byuccl/coast @ 397a26e), on the CPU box; {dt:.0f} s, seed 2026.  Zero mismatches.", "",
             "| reference function | inputs | cases |", "|---|---|---|"]
    lines += [f"| `{a}` | {b} | {c} |" for a, b, c in rows]
    text = "\n".join(lines) + "\n"
    with open(os.path.join(ROOT, "profiles", "r02_oracle_pin_extended.md"), "w") as f:
        f.write(text)
    print(text)


if __name__ == "__main__":
    main()
