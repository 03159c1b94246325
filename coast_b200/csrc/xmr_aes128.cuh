// xmr_aes128.cuh -- protected AES-128 single-block (tests/aes/TI_aes_128.c:107-231 of byuccl/coast)
//
// Unit = one 16-byte block -> 16 bytes.  SoR exit = the 16 u8 state bytes (:224-229): 16 votes.
// Two kernels, same results:
//   xmr_aes128_enc  : encrypt, one ECB key (BASELINE config 3).  Column/T-table formulation of the
//                     same rounds: AddRoundKey-then-SubBytes (:143-146), ShiftRows (:147-166),
//                     MixColumns (:169-184), forward key schedule (:214-221), last AddRoundKey (:224-229).
//                     TE0..TE3 are replicated 32x in shared memory (row x = 32 lanes of TE_k[x]) so every lookup
//                     is bank-conflict free whatever the data and its address is ONE byte-permute; blocks
//                     arrive through the TMA tile ring.
//   xmr_aes128_gen  : byte-wise restatement that follows the TI control flow literally; handles
//                     decrypt (dir=1, :112-129,133-141,187-212) and per-unit keys (the 568 KATs).
// Fault sites (identical in oracle/): 0..15 = state byte as loaded; 16+16r+i = state[i] at the
// bottom of main-loop iteration r (:132-223).  A flip there is XOR-linear through the following
// AddRoundKey, which is why the T-table kernel can apply it after its fused key add.
#pragma once
#include "xmr_common.cuh"
#include "aes_tables.inc"

namespace xmr {

// ---- T-table kernel geometry --------------------------------------------------------------------
// 512-thread CTAs, one per SM.  Shared-memory WINDOW layout (absolute shared::cta addresses):
//   [dyn base .. 0x10000)  TMA tile ring (2 stages)
//   [0x10000 .. 0x20000)   row x (256 B): TE0[x] replicated over 32 lanes | TE1[x] = rotl8  replicated over 32 lanes
//   [0x20000 .. 0x30000)   row x (256 B): TE2[x] = rotl16 x 32 lanes     | TE3[x] = rotl24 x 32 lanes
// Row stride 256 B + 64 KiB alignment make the lookup address ONE byte-permute:
//   addr = PRMT(t, lanebase) = lanebase.b3 : lanebase.b2 : byte_k(t) : lanebase.b0,  lanebase = table | half | lane*4
// and bank = lane for every lane whatever the data -> no shared-memory bank conflicts, no rotates, no LEA.
// (r01 first version: one TE0 copy + PRMT rotates + SHF/LOP3/LEA per lookup = 975 instructions per block; ncu showed
//  the ALU pipe at 95 %.)
constexpr int AES_THREADS = 512, AES_WARPS = 16;
constexpr uint32_t AES_TAB01 = 0x10000u, AES_TAB23 = 0x20000u, AES_WINDOW_END = 0x30000u;
template <int NC> struct AesGeom { static constexpr int J = NC == 1 ? 2 : 4; static constexpr int TROWS = AES_WARPS * Lanes<NC>::kUnitsPerWarp * J; };

__device__ __forceinline__ uint32_t lds32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
// table row of byte k of t: splice that byte into byte 1 of the lane's base address
__device__ __forceinline__ uint32_t tab_b0(uint32_t lb, uint32_t t) { return lds32(__byte_perm(t, lb, 0x7604u)); }
__device__ __forceinline__ uint32_t tab_b1(uint32_t lb, uint32_t t) { return lds32(__byte_perm(t, lb, 0x7614u)); }
__device__ __forceinline__ uint32_t tab_b2(uint32_t lb, uint32_t t) { return lds32(__byte_perm(t, lb, 0x7624u)); }
__device__ __forceinline__ uint32_t tab_b3(uint32_t lb, uint32_t t) { return lds32(__byte_perm(t, lb, 0x7634u)); }

template <int NC>
__device__ __forceinline__ void aes_vote_store(const uint32_t (&c)[4], uint8_t* out, unsigned long long local,
                                               unsigned long long gunit, bool valid, int lane, uint32_t flags, Tally& tally) {
    const bool majority = flags & COAST_F_MAJORITY_D;
    uint32_t o[4], bad = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { Voted v = vote_u32<NC, 1>(c[i], majority); o[i] = v.vote; bad += v.bad; }
    if (valid && Lanes<NC>::voter(lane)) {
        *reinterpret_cast<uint4*>(out + local * 16ull) = make_uint4(o[0], o[1], o[2], o[3]);
        tally.unit_exit<NC>(bad, 16u, flags, gunit);
    }
}

template <int NC, bool INJECT>
__device__ __forceinline__ void aes128_enc_body(const xmr_args& a, const CUtensorMap* tmap) {
    constexpr int UPW = Lanes<NC>::kUnitsPerWarp;
    constexpr int J = AesGeom<NC>::J;
    constexpr int TROWS = AesGeom<NC>::TROWS;
    using Ring = TileRing<TROWS, 16>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t win = smem_u32(smem_raw);                    // shared-window address of the dynamic region
    uint8_t* ring_mem = smem_raw + ((1024u - (win & 1023u)) & 1023u);
    Ring ring;
    ring.init(ring_mem, tmap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // build the two 64 KiB tables
        uint32_t* t01 = reinterpret_cast<uint32_t*>(smem_raw + (AES_TAB01 - win));
        uint32_t* t23 = reinterpret_cast<uint32_t*>(smem_raw + (AES_TAB23 - win));
        for (int i = tid; i < 256 * 64; i += AES_THREADS) {
            const uint32_t v = XMR_AES_TE0[i >> 6];
            const bool hi = (i & 32) != 0;
            t01[i] = hi ? __byte_perm(v, 0u, 0x2103u) : v;                               // TE1 = rotl8
            t23[i] = hi ? __byte_perm(v, 0u, 0x0321u) : __byte_perm(v, 0u, 0x1032u);     // TE3 = rotl24 : TE2 = rotl16
        }
    }
    __syncthreads();
    const uint32_t lb0 = AES_TAB01 + 4u * lane, lb1 = AES_TAB01 + 128u + 4u * lane;
    const uint32_t lb2 = AES_TAB23 + 4u * lane, lb3 = AES_TAB23 + 128u + 4u * lane;
    const int r = Lanes<NC>::replica(lane);
    const int u = Lanes<NC>::unit(lane);

    // every replica lane expands ITS OWN copy of the key (cloneGlobals: key[] is per-replica memory)
    uint32_t rk[44];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        rk[i] = (uint32_t)a.key[4 * i] | ((uint32_t)a.key[4 * i + 1] << 8) | ((uint32_t)a.key[4 * i + 2] << 16) | ((uint32_t)a.key[4 * i + 3] << 24);
#pragma unroll
    for (int rd = 0; rd < 10; ++rd) {                          // :214-221; S-box byte = byte 1 of TE0 = byte 2 of TE1 ...
        const uint32_t w = rk[4 * rd + 3];
        // SubWord(RotWord(w)): bytes (S[w.b1], S[w.b2], S[w.b3], S[w.b0])
        const uint32_t sw = (tab_b1(lb2, w) & 0x000000FFu) | (tab_b2(lb0, w) & 0x0000FF00u) |
                            (tab_b3(lb0, w) & 0x00FF0000u) | (tab_b0(lb1, w) & 0xFF000000u);
        rk[4 * rd + 4] = rk[4 * rd] ^ sw ^ (uint32_t)XMR_AES_RCON[rd];
        rk[4 * rd + 5] = rk[4 * rd + 1] ^ rk[4 * rd + 4];
        rk[4 * rd + 6] = rk[4 * rd + 2] ^ rk[4 * rd + 5];
        rk[4 * rd + 7] = rk[4 * rd + 3] ^ rk[4 * rd + 6];
    }

    const uint32_t n_tiles = a.n_tiles;
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) ring.issue(0, tile);
    Tally tally(a);
    uint32_t it = 0;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t next = tile + gridDim.x;
        if (next < n_tiles) ring.issue((it + 1u) & 1u, next);
        const uint8_t* base = ring.wait(it);
        uint32_t s[J][4];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            uint4 q = *reinterpret_cast<const uint4*>(base + ((warp * J + j) * UPW + u) * 16);
            s[j][0] = q.x; s[j][1] = q.y; s[j][2] = q.z; s[j][3] = q.w;
        }
        __syncthreads();

        unsigned long long local[J];
        bool valid[J];
        // fault of block j as branch-free masks: column word fcol[j], shifted bit fbit[j], applied when frd[j] == round
        // (frd = -1: the replica's input copy, before round 0)
        uint32_t fbit[J]; int frd[J], fcol[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            local[j] = (unsigned long long)tile * TROWS + (unsigned)((warp * J + j) * UPW + u);
            valid[j] = local[j] < a.n_units;
            fbit[j] = 0u; frd[j] = -2; fcol[j] = 0;
            if (INJECT) {
                Fault f = fault_for_unit(a, NC, valid[j] ? local[j] : 0ull, [](uint32_t) { return 8u; });
                if (f.active && valid[j]) {
                    if (Lanes<NC>::voter(lane)) tally.injected++;
                    if ((int)f.replica == r) {
                        const uint32_t i = f.site < 16u ? f.site : ((f.site - 16u) & 15u);
                        frd[j] = f.site < 16u ? -1 : (int)((f.site - 16u) >> 4);
                        fcol[j] = (int)(i >> 2);
                        fbit[j] = (1u << f.bit) << (8u * (i & 3u));
                    }
                }
                const uint32_t hit = frd[j] == -1 ? fbit[j] : 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c) s[j][c] ^= fcol[j] == c ? hit : 0u;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) s[j][c] ^= rk[c];       // first half of :143-146 (state ^ key)
        }
#pragma unroll
        for (int rd = 0; rd < 10; ++rd) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t t0 = s[j][0], t1 = s[j][1], t2 = s[j][2], t3 = s[j][3];
                uint32_t n[4];
                if (rd < 9) {                                   // SubBytes + ShiftRows + MixColumns = 4 table rows XORed
                    n[0] = tab_b0(lb0, t0) ^ tab_b1(lb1, t1) ^ tab_b2(lb2, t2) ^ tab_b3(lb3, t3);
                    n[1] = tab_b0(lb0, t1) ^ tab_b1(lb1, t2) ^ tab_b2(lb2, t3) ^ tab_b3(lb3, t0);
                    n[2] = tab_b0(lb0, t2) ^ tab_b1(lb1, t3) ^ tab_b2(lb2, t0) ^ tab_b3(lb3, t1);
                    n[3] = tab_b0(lb0, t3) ^ tab_b1(lb1, t0) ^ tab_b2(lb2, t1) ^ tab_b3(lb3, t2);
                } else {                                        // round 9: no MixColumns (:168); S[x] sits in byte p of the table picked per position
                    n[0] = (tab_b0(lb2, t0) & 0x000000FFu) | (tab_b1(lb0, t1) & 0x0000FF00u) | (tab_b2(lb0, t2) & 0x00FF0000u) | (tab_b3(lb1, t3) & 0xFF000000u);
                    n[1] = (tab_b0(lb2, t1) & 0x000000FFu) | (tab_b1(lb0, t2) & 0x0000FF00u) | (tab_b2(lb0, t3) & 0x00FF0000u) | (tab_b3(lb1, t0) & 0xFF000000u);
                    n[2] = (tab_b0(lb2, t2) & 0x000000FFu) | (tab_b1(lb0, t3) & 0x0000FF00u) | (tab_b2(lb0, t0) & 0x00FF0000u) | (tab_b3(lb1, t1) & 0xFF000000u);
                    n[3] = (tab_b0(lb2, t3) & 0x000000FFu) | (tab_b1(lb0, t0) & 0x0000FF00u) | (tab_b2(lb0, t1) & 0x00FF0000u) | (tab_b3(lb1, t2) & 0xFF000000u);
                }
                if (INJECT) {
                    const uint32_t hit = frd[j] == rd ? fbit[j] : 0u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) n[c] ^= fcol[j] == c ? hit : 0u;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) s[j][c] = n[c] ^ rk[4 * (rd + 1) + c];   // next round's / last AddRoundKey
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j)
            aes_vote_store<NC>(s[j], static_cast<uint8_t*>(a.out), local[j], a.unit_base + local[j], valid[j], lane, a.flags, tally);
    }
    tally.flush(a.counters);
}

// ---------------------------------------------------------------------------------------------
// General path: literal byte-wise control flow of aes_enc_dec(), both directions, optional
// per-unit keys.  S-boxes live in shared memory (512 B).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t xtime8(uint8_t v) { return (uint8_t)((v << 1) ^ ((v & 0x80) ? 0x1b : 0)); }   // galois_mul2 :88-99

template <int NC, bool INJECT>
__device__ __forceinline__ void aes128_gen_body(const xmr_args& a) {
    __shared__ uint8_t S[256], IS[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { S[i] = XMR_AES_SBOX[i]; IS[i] = XMR_AES_RSBOX[i]; }
    __syncthreads();
    constexpr int UPW = Lanes<NC>::kUnitsPerWarp;
    const int lane = threadIdx.x & 31;
    const int r = Lanes<NC>::replica(lane);
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long n_wtiles = (a.n_units + UPW - 1) / UPW;
    const bool dir = a.mode & 1u, per_unit = a.mode & 2u;
    Tally tally(a);
    for (unsigned long long wt = gwarp; wt < n_wtiles; wt += nwarps) {
        const unsigned long long local = wt * UPW + Lanes<NC>::unit(lane);
        const bool valid = local < a.n_units;
        const unsigned long long ld = valid ? local : 0ull;
        uint8_t s[16], k[16];
        {
            uint4 q = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(a.in) + ld * 16ull);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
            if (per_unit) {
                uint4 kq = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(a.aux) + ld * 16ull);
                uint32_t kw[4] = {kq.x, kq.y, kq.z, kq.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) k[i] = (uint8_t)(kw[i >> 2] >> (8 * (i & 3)));
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) k[i] = a.key[i];
            }
        }
        uint32_t fsite = 0xFFFFFFFFu; uint8_t fmask = 0;
        if (INJECT) {
            Fault f = fault_for_unit(a, NC, ld, [](uint32_t) { return 8u; });
            if (f.active && valid) {
                if (Lanes<NC>::voter(lane)) tally.injected++;
                if ((int)f.replica == r) { fsite = f.site; fmask = (uint8_t)(1u << f.bit); }
            }
            if (fsite < 16u) {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (fsite == (uint32_t)i) s[i] ^= fmask;
            }
        }
        if (dir) {                                              // :112-129
            for (int rd = 0; rd < 10; ++rd) {
                k[0] ^= S[k[13]] ^ XMR_AES_RCON[rd]; k[1] ^= S[k[14]]; k[2] ^= S[k[15]]; k[3] ^= S[k[12]];
#pragma unroll
                for (int i = 4; i < 16; ++i) k[i] ^= k[i - 4];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] ^= k[i];
        }
        for (int rd = 0; rd < 10; ++rd) {                       // :132
            if (dir) {
#pragma unroll
                for (int i = 15; i > 3; --i) k[i] ^= k[i - 4];  // :134-137
                k[0] ^= S[k[13]] ^ XMR_AES_RCON[9 - rd]; k[1] ^= S[k[14]]; k[2] ^= S[k[15]]; k[3] ^= S[k[12]];   // :138-141
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) s[i] = S[s[i] ^ k[i]];   // :143-146
                uint8_t t;                                      // :147-166 shift rows
                t = s[1]; s[1] = s[5]; s[5] = s[9]; s[9] = s[13]; s[13] = t;
                t = s[2]; s[2] = s[10]; s[10] = t; t = s[6]; s[6] = s[14]; s[14] = t;
                t = s[15]; s[15] = s[11]; s[11] = s[7]; s[7] = s[3]; s[3] = t;
            }
            if ((rd > 0 && dir) || (rd < 9 && !dir)) {          // :168-185
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint8_t* p = s + 4 * c;
                    if (dir) {                                  // :172-177 inverse pre-multiply
                        uint8_t b1 = xtime8(xtime8(p[0] ^ p[2])), b2 = xtime8(xtime8(p[1] ^ p[3]));
                        p[0] ^= b1; p[1] ^= b2; p[2] ^= b1; p[3] ^= b2;
                    }
                    uint8_t all = p[0] ^ p[1] ^ p[2] ^ p[3], first = p[0];
                    p[0] ^= xtime8(p[0] ^ p[1]) ^ all;
                    p[1] ^= xtime8(p[1] ^ p[2]) ^ all;
                    p[2] ^= xtime8(p[2] ^ p[3]) ^ all;
                    p[3] ^= xtime8(p[3] ^ first) ^ all;
                }
            }
            if (dir) {
                uint8_t t;                                      // :187-206 inverse shift rows
                t = s[13]; s[13] = s[9]; s[9] = s[5]; s[5] = s[1]; s[1] = t;
                t = s[10]; s[10] = s[2]; s[2] = t; t = s[14]; s[14] = s[6]; s[6] = t;
                t = s[3]; s[3] = s[7]; s[7] = s[11]; s[11] = s[15]; s[15] = t;
#pragma unroll
                for (int i = 0; i < 16; ++i) s[i] = IS[s[i]] ^ k[i];   // :208-211
            } else {
                k[0] ^= S[k[13]] ^ XMR_AES_RCON[rd]; k[1] ^= S[k[14]]; k[2] ^= S[k[15]]; k[3] ^= S[k[12]];   // :214-221
#pragma unroll
                for (int i = 4; i < 16; ++i) k[i] ^= k[i - 4];
            }
            if (INJECT && fsite >= 16u && (fsite - 16u) >> 4 == (uint32_t)rd) {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (((fsite - 16u) & 15u) == (uint32_t)i) s[i] ^= fmask;
            }
        }
        if (!dir) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] ^= k[i];          // :224-229
        }
        uint32_t c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            c[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
        aes_vote_store<NC>(c, static_cast<uint8_t*>(a.out), local, a.unit_base + local, valid, lane, a.flags, tally);
        if ((a.mode & 4u) && per_unit && valid && Lanes<NC>::voter(lane)) {      // COAST_AES_KEY_WRITEBACK: replica 0's mutated key[]
            uint32_t kw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                kw[i] = (uint32_t)k[4 * i] | ((uint32_t)k[4 * i + 1] << 8) | ((uint32_t)k[4 * i + 2] << 16) | ((uint32_t)k[4 * i + 3] << 24);
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(const_cast<void*>(a.aux)) + local * 16ull) = make_uint4(kw[0], kw[1], kw[2], kw[3]);
        }
    }
    tally.flush(a.counters);
}

}  // namespace xmr

#define XMR_AES_ENC_KERNEL(NC, INJ)                                                                      \
    extern "C" __global__ void __launch_bounds__(xmr::AES_THREADS, 1)                                    \
    xmr_aes128_enc_nc##NC##_inj##INJ(const __grid_constant__ xmr_args a, const __grid_constant__ CUtensorMap tmap) { \
        xmr::aes128_enc_body<NC, INJ != 0>(a, &tmap);                                                    \
    }
#define XMR_AES_GEN_KERNEL(NC, INJ)                                                                      \
    extern "C" __global__ void __launch_bounds__(XMR_CTA_THREADS)                                        \
    xmr_aes128_gen_nc##NC##_inj##INJ(const __grid_constant__ xmr_args a) {                               \
        xmr::aes128_gen_body<NC, INJ != 0>(a);                                                           \
    }
XMR_AES_ENC_KERNEL(1, 0) XMR_AES_ENC_KERNEL(2, 0) XMR_AES_ENC_KERNEL(3, 0)
XMR_AES_ENC_KERNEL(1, 1) XMR_AES_ENC_KERNEL(2, 1) XMR_AES_ENC_KERNEL(3, 1)
XMR_AES_GEN_KERNEL(1, 0) XMR_AES_GEN_KERNEL(2, 0) XMR_AES_GEN_KERNEL(3, 0)
XMR_AES_GEN_KERNEL(1, 1) XMR_AES_GEN_KERNEL(2, 1) XMR_AES_GEN_KERNEL(3, 1)
