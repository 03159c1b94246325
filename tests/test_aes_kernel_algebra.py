"""CPU model of the identities xmr_aes128.cuh relies on, checked against the oracle (which is pinned on the reference's own
aes_enc_dec() and the 568 NIST vectors): the column/T-table form of both directions, the decrypt loop that carries
InvMixColumns(state) between iterations (v' = TD rows(v) ^ InvMixColumns(round key)), the packed-word GF(2^8) column
arithmetic, the on-the-fly forward/inverse key schedule with what it leaves in key[], the byte-permute selectors, and the
fault hooks (a flip at the bottom of iteration r lands on the fused value; for decrypt through InvMixColumns).
Mirrors the device code statement by statement; no GPU."""
import random

import numpy as np
import pytest

M32 = 0xFFFFFFFF


def byte_perm(a, b, sel):                                    # PRMT, default mode
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    return sum(src[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))


def xtime4(w):
    return (((w & 0x7F7F7F7F) << 1) ^ (((w >> 7) & 0x01010101) * 0x1B)) & M32


def mix_column(w):
    r1, r2, r3 = byte_perm(w, 0, 0x0321), byte_perm(w, 0, 0x1032), byte_perm(w, 0, 0x2103)
    return xtime4(w ^ r1) ^ r1 ^ r2 ^ r3


def inv_mix_column(w):
    v = w ^ byte_perm(w, 0, 0x1032)
    return mix_column(w ^ xtime4(xtime4(v)))


@pytest.fixture(scope="module")
def tabs(oracle):
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "coast_b200", "csrc", "aes_tables.inc")).read()   # generated from FIPS-197
    def arr(name):
        body = src[src.index(name):]
        body = body[body.index("{") + 1: body.index("}")]
        return [int(t.rstrip("u"), 16) for t in re.findall(r"0x[0-9a-fA-F]+u?", body)]
    S, IS, TE0, RCON = arr("XMR_AES_SBOX"), arr("XMR_AES_RSBOX"), arr("XMR_AES_TE0"), arr("XMR_AES_RCON")
    assert len(S) == len(IS) == len(TE0) == 256 and len(RCON) >= 10
    rot = lambda v, k: ((v << (8 * k)) | (v >> (32 - 8 * k))) & M32 if k else v
    T = {"S": S, "IS": IS, "RCON": RCON}
    T["TE"] = [[rot(TE0[x], k) for x in range(256)] for k in range(4)]
    T["TD"] = [[rot(inv_mix_column(IS[x]), k) for x in range(256)] for k in range(4)]
    T["SIS"] = [IS[x] | (S[x] << 8) | (IS[x] << 16) | (S[x] << 24) for x in range(256)]
    return T


def B(w, k):
    return (w >> (8 * k)) & 0xFF


class Model:
    def __init__(self, T):
        self.T = T

    def sub_rot_word(self, dec, w):
        T = self.T
        if dec:                                              # S in byte 1 of the (InvS, S, InvS, S) rows, gathered with two PRMT levels
            a, b, c, d = T["SIS"][B(w, 1)], T["SIS"][B(w, 2)], T["SIS"][B(w, 3)], T["SIS"][B(w, 0)]
            return byte_perm(byte_perm(a, b, 0x0051), byte_perm(c, d, 0x0051), 0x5410)
        TE = T["TE"]
        return (TE[2][B(w, 1)] & 0xFF) | (TE[0][B(w, 2)] & 0xFF00) | (TE[0][B(w, 3)] & 0xFF0000) | (TE[1][B(w, 0)] & 0xFF000000)

    def key_next(self, dec, k, rd):
        k[0] ^= self.sub_rot_word(dec, k[3]) ^ self.T["RCON"][rd]
        k[1] ^= k[0]; k[2] ^= k[1]; k[3] ^= k[2]

    def key_prev(self, dec, k, rd):
        k[3] ^= k[2]; k[2] ^= k[1]; k[1] ^= k[0]
        k[0] ^= self.sub_rot_word(dec, k[3]) ^ self.T["RCON"][rd]

    def round_cols(self, dec, last, t):
        T = self.T
        n = [0] * 4
        for c in range(4):
            if not dec and not last:
                n[c] = T["TE"][0][B(t[c], 0)] ^ T["TE"][1][B(t[(c + 1) & 3], 1)] ^ T["TE"][2][B(t[(c + 2) & 3], 2)] ^ T["TE"][3][B(t[(c + 3) & 3], 3)]
            elif not dec:
                n[c] = (T["TE"][2][B(t[c], 0)] & 0xFF) | (T["TE"][0][B(t[(c + 1) & 3], 1)] & 0xFF00) | \
                       (T["TE"][0][B(t[(c + 2) & 3], 2)] & 0xFF0000) | (T["TE"][1][B(t[(c + 3) & 3], 3)] & 0xFF000000)
            elif not last:
                n[c] = T["TD"][0][B(t[c], 0)] ^ T["TD"][1][B(t[(c + 3) & 3], 1)] ^ T["TD"][2][B(t[(c + 2) & 3], 2)] ^ T["TD"][3][B(t[(c + 1) & 3], 3)]
            else:
                a, b = T["SIS"][B(t[c], 0)], T["SIS"][B(t[(c + 3) & 3], 1)]
                d, e = T["SIS"][B(t[(c + 2) & 3], 2)], T["SIS"][B(t[(c + 1) & 3], 3)]
                n[c] = byte_perm(byte_perm(a, b, 0x0040), byte_perm(d, e, 0x0040), 0x5410)
        return n

    def expand(self, dec, key_words):
        """the one-key register file of aes128_body"""
        rk = [0] * 44
        k = list(key_words)
        if not dec:
            rk[0:4] = k
            for rd in range(10):
                self.key_next(False, k, rd)
                rk[4 * (rd + 1): 4 * (rd + 1) + 4] = k
        else:
            rk[40:44] = k
            for rd in range(10):
                self.key_next(True, k, rd)
                for c in range(4):
                    rk[(4 * (9 - rd) + c) if rd < 9 else c] = inv_mix_column(k[c]) if rd < 9 else k[c]
        return rk

    def block(self, dec, perkey, state_words, key_words, fault=None):
        """aes128_body for one block of one replica; fault = (site, bit) or None.  Returns (out words, key words left)."""
        s = list(state_words)
        k = list(key_words)
        rk = None if perkey else self.expand(dec, key_words)
        frd, fcol, fbit = -2, 0, 0
        if fault:
            site, bit = fault
            i = site if site < 16 else (site - 16) & 15
            frd = -1 if site < 16 else (site - 16) >> 4
            fcol, fbit = i >> 2, ((1 << bit) << (8 * (i & 3))) & M32
            if frd < 0:
                s[fcol] ^= fbit
        if perkey and dec:
            for rd in range(10):
                self.key_next(True, k, rd)
        for c in range(4):
            s[c] ^= k[c] if perkey else rk[c]
        for rd in range(10):
            n = self.round_cols(dec, rd == 9, s)
            hit = fbit if frd == rd else 0
            if dec and rd < 9:
                hit = inv_mix_column(hit)
            n[fcol] ^= hit
            if perkey:
                if not dec:
                    self.key_next(False, k, rd)
                    s = [n[c] ^ k[c] for c in range(4)]
                else:
                    self.key_prev(True, k, 9 - rd)
                    s = [n[c] ^ (inv_mix_column(k[c]) if rd < 9 else k[c]) for c in range(4)]
            else:
                s = [n[c] ^ rk[4 * (rd + 1) + c] for c in range(4)]
        return s, k


def words(b):
    return [int.from_bytes(b[4 * i: 4 * i + 4], "little") for i in range(4)]


def unwords(w):
    return b"".join(x.to_bytes(4, "little") for x in w)


def test_packed_column_arithmetic_matches_the_table_form(tabs):
    TE0 = tabs["TE"][0]
    for x in range(256):
        assert mix_column(tabs["S"][x]) == TE0[x]                           # MixColumns of the column (S[x], 0, 0, 0)
    rnd = random.Random(5)
    for _ in range(2000):
        w = rnd.getrandbits(32)
        assert inv_mix_column(mix_column(w)) == w and mix_column(inv_mix_column(w)) == w
        a, b = rnd.getrandbits(32), rnd.getrandbits(32)
        assert inv_mix_column(a ^ b) == inv_mix_column(a) ^ inv_mix_column(b)   # the linearity the decrypt fusion and its fault hook rely on


def test_half_row_address_trick():
    """addr = PRMT(t, 2*lanebase) >> 1 for the 128-byte-row table at 0x30000"""
    for lane in range(32):
        lb2x = 2 * (0x30000 + 4 * lane)
        for k in range(4):
            for x in (0, 1, 0x7F, 0x80, 0xFF):
                t = x << (8 * k)
                assert byte_perm(t, lb2x, 0x7604 | (k << 4)) >> 1 == 0x30000 + 128 * x + 4 * lane
        for k in range(4):
            lb = 0x10000 + 128 + 4 * lane
            assert byte_perm(0xAB << (8 * k), lb, 0x7604 | (k << 4)) == 0x10000 + 256 * 0xAB + 128 + 4 * lane


@pytest.mark.parametrize("dec", [False, True])
@pytest.mark.parametrize("perkey", [False, True])
def test_table_form_equals_the_oracle_both_directions_and_key_left_behind(oracle, tabs, dec, perkey):
    m = Model(tabs)
    rnd = random.Random(11 + 2 * dec + perkey)
    for _ in range(300):
        st, key = bytes(rnd.getrandbits(8) for _ in range(16)), bytes(rnd.getrandbits(8) for _ in range(16))
        want, key_after = oracle.aes128(st, key, 1 if dec else 0)
        got, k = m.block(dec, perkey, words(st), words(key))
        assert unwords(got) == want
        if perkey:                                           # what aes_enc_dec() leaves in key[] (TI_aes_128.c:214-221 / :133-141)
            assert unwords(k) == key_after
    if dec:                                                  # decrypt(encrypt(x)) == x through the two table forms
        st, key = bytes(range(16)), bytes(range(16, 32))
        ct, _ = m.block(False, perkey, words(st), words(key))
        pt, _ = m.block(True, perkey, ct, words(key))
        assert unwords(pt) == st


@pytest.mark.parametrize("dec", [False, True])
@pytest.mark.parametrize("perkey", [False, True])
def test_every_fault_site_lands_where_the_oracle_puts_it(oracle, tabs, dec, perkey):
    """all 176 sites x a few bits through the oracle's run() with a TABLE plan (one replica flipped, nc=1 so the flipped value is the output)"""
    m = Model(tabs)
    rnd = random.Random(3)
    sites = list(range(176))
    n = len(sites)
    blocks = np.frombuffer(bytes(rnd.getrandbits(8) for _ in range(16 * n)), dtype=np.uint8).copy()
    keys = np.frombuffer(bytes(rnd.getrandbits(8) for _ in range(16 * n)), dtype=np.uint8).copy()
    bits = [rnd.randrange(8) for _ in sites]
    table = np.array([oracle.fault_entry(0, s, b) for s, b in zip(sites, bits)], dtype=np.uint32)
    mode = (1 if dec else 0) | 2
    out, st = oracle.run(oracle.K_AES128, 1, blocks, n, mode=mode, aux=keys, plan=oracle.make_plan(oracle.PLAN_TABLE, table=table))
    assert st["injected"] == n
    for u, (s, b) in enumerate(zip(sites, bits)):
        got, _ = m.block(dec, perkey, words(blocks[16 * u: 16 * u + 16].tobytes()), words(keys[16 * u: 16 * u + 16].tobytes()), fault=(s, b))
        assert unwords(got) == out[16 * u: 16 * u + 16].tobytes(), (dec, perkey, s, b)
