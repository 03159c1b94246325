"""CPU model of the two identities the 64-byte CRC16 kernel (coast_b200/csrc/xmr_crc16.cuh) relies on, checked against the
oracle for every fault site and bit -- so the device-side tricks are pinned without a GPU:

1. the byte-position-rotating table form  x_(i+1) = t_(i-1).lo ^ t_i.hi ^ b_(i+1)  equals crc16.c:21-31;
2. a flip of `crc` after byte s is a flip of data byte s+1 (bits 8..15) or s+2 (bits 0..7), or lands in the result.
"""
import numpy as np


def step0(x):                       # crc16_step(0, x) of the kernel = crc16.c:26-28 with crc == 0
    x &= 0xFF
    x ^= x >> 4
    return ((x << 12) ^ (x << 5) ^ x) & 0xFFFF


def crc_rotating_tables(msg: bytes) -> int:
    """the device loop: two table words W1 (even steps) / W2 (odd steps), state = the last two table words"""
    W1, W2 = [], []
    for x in range(256):
        t = step0(x)
        h, l = t >> 8, t & 0xFF
        W1.append(l | (h << 8) | (l << 16) | (h << 24))
        W2.append(h | (l << 8) | (h << 16) | (l << 24))
    words = np.frombuffer(msg, dtype="<u4")
    t1, t2 = 0xFFFF, 0
    for i in range(len(msg)):
        v = t2 ^ t1 ^ int(words[i >> 2])
        x = (v >> (8 * (i & 3))) & 0xFF                      # the PRMT picks byte i & 3
        t2, t1 = t1, (W2 if i & 1 else W1)[x]
    return (((t2 ^ t1) & 0xFF) << 8) | ((t1 >> 8) & 0xFF)


def folded_fault(msg: bytes, site: int, bit: int) -> int:
    """the injector's form: one XOR into one message word (or into the result) instead of a hook at every site"""
    L = len(msg)
    data, high = site >= L, bit >= 8
    pos = site - L if data else site + (1 if high else 2)
    m8 = 1 << (bit - 8 if high else bit)
    buf = bytearray(msg)
    ffin = 0
    if pos < L:
        buf[pos] ^= m8
    else:
        ffin = (1 << bit) if high else ((m8 << 8) if pos == L else m8)
    return crc_rotating_tables(bytes(buf)) ^ ffin


def test_rotating_table_form_equals_the_reference_byte_step(oracle, golden):
    rng = np.random.default_rng(5)
    for _ in range(200):
        msg = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        assert crc_rotating_tables(msg) == oracle.crc16(msg)
    # shorter multiples of 4 exercise the same recurrences; the shipped 13-byte message goes through the general kernel
    for n in (4, 8, 12, 60):
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert crc_rotating_tables(msg) == oracle.crc16(msg)


def test_folded_fault_equals_the_hook_at_every_site_and_bit(oracle):
    rng = np.random.default_rng(6)
    for trial in range(3):
        msg = rng.integers(0, 256, 64, dtype=np.uint8)
        for site in range(128):
            for bit in range(oracle.fault_site_bits(oracle.K_CRC16, 64, 0, site)):
                tab = np.array([oracle.fault_entry(0, site, bit)], dtype=np.uint32)
                out, st = oracle.run(oracle.K_CRC16, 1, msg, 1, unit_bytes=64, plan=oracle.make_plan(oracle.PLAN_TABLE, table=tab))
                assert st["injected"] == 1
                assert int(out.view(np.uint16)[0]) == folded_fault(msg.tobytes(), site, bit), (site, bit)
