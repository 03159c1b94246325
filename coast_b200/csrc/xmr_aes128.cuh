// xmr_aes128.cuh -- protected AES-128 single-block (tests/aes/TI_aes_128.c:107-231 of byuccl/coast)
//
// Unit = one 16-byte block -> 16 bytes.  SoR exit = the 16 u8 state bytes (:224-229): 16 votes.
// Two kernels, same results:
//   xmr_aes128_enc  : encrypt, one ECB key (BASELINE config 3).  Column/T-table formulation of the
//                     same rounds: AddRoundKey-then-SubBytes (:143-146), ShiftRows (:147-166),
//                     MixColumns (:169-184), forward key schedule (:214-221), last AddRoundKey (:224-229).
//                     TE0 is replicated 32x in shared memory as te[x][lane] so every lookup is
//                     bank-conflict free whatever the data; blocks arrive through the TMA tile ring.
//   xmr_aes128_gen  : byte-wise restatement that follows the TI control flow literally; handles
//                     decrypt (dir=1, :112-129,133-141,187-212) and per-unit keys (the 568 KATs).
// Fault sites (identical in oracle/): 0..15 = state byte as loaded; 16+16r+i = state[i] at the
// bottom of main-loop iteration r (:132-223).  A flip there is XOR-linear through the following
// AddRoundKey, which is why the T-table kernel can apply it after its fused key add.
#pragma once
#include "xmr_common.cuh"
#include "aes_tables.inc"

namespace xmr {

constexpr int AES_J = 4;   // blocks per lane group per tile (ILP across independent blocks)

__device__ __forceinline__ uint32_t rotl8(uint32_t v) { return __byte_perm(v, 0u, 0x2103u); }
__device__ __forceinline__ uint32_t rotl16(uint32_t v) { return __byte_perm(v, 0u, 0x1032u); }
__device__ __forceinline__ uint32_t rotl24(uint32_t v) { return __byte_perm(v, 0u, 0x0321u); }

// te points at this lane's column of the replicated table: te[x * 32]
__device__ __forceinline__ uint32_t te_b0(const uint32_t* te, uint32_t t) { return te[(t & 0xFFu) << 5]; }
__device__ __forceinline__ uint32_t te_b1(const uint32_t* te, uint32_t t) { return te[((t >> 8) & 0xFFu) << 5]; }
__device__ __forceinline__ uint32_t te_b2(const uint32_t* te, uint32_t t) { return te[((t >> 16) & 0xFFu) << 5]; }
__device__ __forceinline__ uint32_t te_b3(const uint32_t* te, uint32_t t) { return te[(t >> 24) << 5]; }
// S-box byte = byte 1 of TE0[x]
__device__ __forceinline__ uint32_t sb(const uint32_t* te, uint32_t x) { return (te[x << 5] >> 8) & 0xFFu; }

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
    constexpr int TROWS = XMR_WARPS * UPW * AES_J;
    using Ring = TileRing<TROWS, 16>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t* te_all = reinterpret_cast<uint32_t*>(smem_raw + Ring::SMEM_BYTES + ((1024 - (Ring::SMEM_BYTES & 1023)) & 1023));
    Ring ring;
    ring.init(smem_raw, tmap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256 * 32; i += XMR_CTA_THREADS) te_all[i] = XMR_AES_TE0[i >> 5];
    __syncthreads();
    const uint32_t* te = te_all + lane;
    const int r = Lanes<NC>::replica(lane);
    const int u = Lanes<NC>::unit(lane);

    // every replica lane expands ITS OWN copy of the key (cloneGlobals: key[] is per-replica memory)
    uint32_t rk[44];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        rk[i] = (uint32_t)a.key[4 * i] | ((uint32_t)a.key[4 * i + 1] << 8) | ((uint32_t)a.key[4 * i + 2] << 16) | ((uint32_t)a.key[4 * i + 3] << 24);
#pragma unroll
    for (int rd = 0; rd < 10; ++rd) {                          // :214-221
        uint32_t w = rk[4 * rd + 3];
        uint32_t rw = __funnelshift_r(w, w, 8);                 // bytes (k13,k14,k15,k12)
        uint32_t sw = sb(te, rw & 0xFFu) | (sb(te, (rw >> 8) & 0xFFu) << 8) | (sb(te, (rw >> 16) & 0xFFu) << 16) | (sb(te, rw >> 24) << 24);
        rk[4 * rd + 4] = rk[4 * rd] ^ sw ^ (uint32_t)XMR_AES_RCON[rd];
        rk[4 * rd + 5] = rk[4 * rd + 1] ^ rk[4 * rd + 4];
        rk[4 * rd + 6] = rk[4 * rd + 2] ^ rk[4 * rd + 5];
        rk[4 * rd + 7] = rk[4 * rd + 3] ^ rk[4 * rd + 6];
    }

    const uint32_t n_tiles = a.n_tiles;
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) ring.issue(0, tile);
    Tally tally;
    uint32_t it = 0;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t next = tile + gridDim.x;
        if (next < n_tiles) ring.issue((it + 1u) & 1u, next);
        const uint8_t* base = ring.wait(it);
        uint32_t s[AES_J][4];
#pragma unroll
        for (int j = 0; j < AES_J; ++j) {
            uint4 q = *reinterpret_cast<const uint4*>(base + ((warp * AES_J + j) * UPW + u) * 16);
            s[j][0] = q.x; s[j][1] = q.y; s[j][2] = q.z; s[j][3] = q.w;
        }
        __syncthreads();

        unsigned long long local[AES_J];
        bool valid[AES_J];
        uint32_t fsite[AES_J], fmask[AES_J];
#pragma unroll
        for (int j = 0; j < AES_J; ++j) {
            local[j] = (unsigned long long)tile * TROWS + (unsigned)((warp * AES_J + j) * UPW + u);
            valid[j] = local[j] < a.n_units;
            fsite[j] = 0xFFFFFFFFu; fmask[j] = 0u;
            if (INJECT) {
                Fault f = fault_for_unit(a, NC, valid[j] ? local[j] : 0ull, [](uint32_t) { return 8u; });
                if (f.active && valid[j]) {
                    if (Lanes<NC>::voter(lane)) tally.injected++;
                    if ((int)f.replica == r) { fsite[j] = f.site; fmask[j] = 1u << f.bit; }
                }
                if (fsite[j] < 16u) s[j][fsite[j] >> 2] ^= fmask[j] << (8u * (fsite[j] & 3u));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) s[j][c] ^= rk[c];       // first half of :143-146 (state ^ key)
        }
#pragma unroll
        for (int rd = 0; rd < 10; ++rd) {
#pragma unroll
            for (int j = 0; j < AES_J; ++j) {
                uint32_t t0 = s[j][0], t1 = s[j][1], t2 = s[j][2], t3 = s[j][3], n[4];
                if (rd < 9) {                                   // SubBytes+ShiftRows+MixColumns via TE0
                    n[0] = te_b0(te, t0) ^ rotl8(te_b1(te, t1)) ^ rotl16(te_b2(te, t2)) ^ rotl24(te_b3(te, t3));
                    n[1] = te_b0(te, t1) ^ rotl8(te_b1(te, t2)) ^ rotl16(te_b2(te, t3)) ^ rotl24(te_b3(te, t0));
                    n[2] = te_b0(te, t2) ^ rotl8(te_b1(te, t3)) ^ rotl16(te_b2(te, t0)) ^ rotl24(te_b3(te, t1));
                    n[3] = te_b0(te, t3) ^ rotl8(te_b1(te, t0)) ^ rotl16(te_b2(te, t1)) ^ rotl24(te_b3(te, t2));
                } else {                                        // round 9: no MixColumns (:168)
                    n[0] = sb(te, t0 & 0xFFu) | (sb(te, (t1 >> 8) & 0xFFu) << 8) | (sb(te, (t2 >> 16) & 0xFFu) << 16) | (sb(te, t3 >> 24) << 24);
                    n[1] = sb(te, t1 & 0xFFu) | (sb(te, (t2 >> 8) & 0xFFu) << 8) | (sb(te, (t3 >> 16) & 0xFFu) << 16) | (sb(te, t0 >> 24) << 24);
                    n[2] = sb(te, t2 & 0xFFu) | (sb(te, (t3 >> 8) & 0xFFu) << 8) | (sb(te, (t0 >> 16) & 0xFFu) << 16) | (sb(te, t1 >> 24) << 24);
                    n[3] = sb(te, t3 & 0xFFu) | (sb(te, (t0 >> 8) & 0xFFu) << 8) | (sb(te, (t1 >> 16) & 0xFFu) << 16) | (sb(te, t2 >> 24) << 24);
                }
                if (INJECT && fsite[j] >= 16u && (fsite[j] - 16u) >> 4 == (uint32_t)rd) {
                    uint32_t i = (fsite[j] - 16u) & 15u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if ((i >> 2) == (uint32_t)c) n[c] ^= fmask[j] << (8u * (i & 3u));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) s[j][c] = n[c] ^ rk[4 * (rd + 1) + c];   // next round's / last AddRoundKey
            }
        }
#pragma unroll
        for (int j = 0; j < AES_J; ++j)
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
    Tally tally;
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
    }
    tally.flush(a.counters);
}

}  // namespace xmr

#define XMR_AES_ENC_KERNEL(NC, INJ)                                                                      \
    extern "C" __global__ void __launch_bounds__(XMR_CTA_THREADS)                                        \
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
