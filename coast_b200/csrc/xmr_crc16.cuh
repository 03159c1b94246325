// xmr_crc16.cuh -- protected CCITT CRC-16 (tests/crc16/crc16.c:21-31 of byuccl/coast)
//
// Unit = one message of unit_bytes (1..255, `unsigned char length` :21) -> u16.  SoR exit = the
// `ret i16` (:30): ONE u16 vote per unit.  Fault sites: s < L: `crc` after byte s (16 bits);
// L <= s < 2L: data byte s-L as loaded by the replica (8 bits).
#pragma once
#include "xmr_common.cuh"

namespace xmr {

// one byte of crc16.c:26-28 on 32-bit registers with the reference's u8/u16 truncations
__device__ __forceinline__ uint32_t crc16_step(uint32_t crc, uint32_t b) {
    uint32_t x = ((crc >> 8) ^ b) & 0xFFu;        // x = crc >> 8 ^ *data_p++   (u8)
    x ^= x >> 4;                                  // x ^= x >> 4
    return ((crc << 8) ^ (x << 12) ^ (x << 5) ^ x) & 0xFFFFu;   // (u16)
}

template <int NC>
__device__ __forceinline__ void crc_vote_store(uint32_t crc, uint16_t* out, unsigned long long local, unsigned long long gunit,
                                               bool valid, int lane, uint32_t flags, Tally& tally) {
    Voted v = vote_u32<NC, 4>(crc, flags & COAST_F_MAJORITY_D);   // one element (the u16 lives alone in the register)
    if (valid && Lanes<NC>::voter(lane)) {
        out[local] = (uint16_t)v.vote;
        tally.unit_exit<NC>(v.bad, 1u, flags, gunit);
    }
}

struct CrcWidth { uint32_t len; __device__ uint32_t operator()(uint32_t site) const { return site < len ? 16u : 8u; } };

// Fast path: unit_bytes == 64, tiles of 8*(32/NC) messages through the TMA ring (64B swizzle).
template <int NC, bool INJECT>
__device__ __forceinline__ void crc16_b64_body(const xmr_args& a, const CUtensorMap* tmap) {
    constexpr int UPW = Lanes<NC>::kUnitsPerWarp;
    constexpr int TU = XMR_WARPS * UPW;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TileRing<TU, 64> ring;
    ring.init(smem_raw, tmap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = Lanes<NC>::replica(lane);
    const int ul = warp * UPW + Lanes<NC>::unit(lane);
    const uint32_t n_tiles = a.n_tiles;
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) ring.issue(0, tile);
    Tally tally(a);
    uint32_t it = 0;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t next = tile + gridDim.x;
        if (next < n_tiles) ring.issue((it + 1u) & 1u, next);
        const uint8_t* base = ring.wait(it);
        uint32_t w[16];
        const uint8_t* row = base + ul * 64;
        const int sw = (ul >> 1) & 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 q = *reinterpret_cast<const uint4*>(row + ((c ^ sw) << 4));
            w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
        }
        __syncthreads();
        const unsigned long long local = (unsigned long long)tile * TU + ul;
        const bool valid = local < a.n_units;
        uint32_t fsite = 0xFFFFFFFFu, fmask = 0u;
        if (INJECT) {
            Fault f = fault_for_unit(a, NC, valid ? local : 0ull, CrcWidth{64u});
            if (f.active && valid) {
                if (Lanes<NC>::voter(lane)) tally.injected++;
                if ((int)f.replica == r) { fsite = f.site; fmask = 1u << f.bit; }
            }
        }
        uint32_t crc = 0xFFFFu;                                // :23
#pragma unroll
        for (int i = 0; i < 64; ++i) {                         // :25
            uint32_t b = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            if (INJECT && fsite == 64u + (uint32_t)i) b ^= fmask;
            crc = crc16_step(crc, b);
            if (INJECT && fsite == (uint32_t)i) crc ^= fmask;
        }
        crc_vote_store<NC>(crc, static_cast<uint16_t*>(a.out), local, a.unit_base + local, valid, lane, a.flags, tally);
    }
    tally.flush(a.counters);
}

// General path: any length 1..255, bytes read straight from global memory.
template <int NC, bool INJECT>
__device__ __forceinline__ void crc16_gen_body(const xmr_args& a) {
    constexpr int UPW = Lanes<NC>::kUnitsPerWarp;
    const int lane = threadIdx.x & 31;
    const int r = Lanes<NC>::replica(lane);
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long n_wtiles = (a.n_units + UPW - 1) / UPW;
    const uint32_t len = a.unit_bytes;
    Tally tally(a);
    for (unsigned long long wt = gwarp; wt < n_wtiles; wt += nwarps) {
        const unsigned long long local = wt * UPW + Lanes<NC>::unit(lane);
        const bool valid = local < a.n_units;
        const uint8_t* msg = static_cast<const uint8_t*>(a.in) + (valid ? local : 0ull) * len;
        uint32_t fsite = 0xFFFFFFFFu, fmask = 0u;
        if (INJECT) {
            Fault f = fault_for_unit(a, NC, valid ? local : 0ull, CrcWidth{len});
            if (f.active && valid) {
                if (Lanes<NC>::voter(lane)) tally.injected++;
                if ((int)f.replica == r) { fsite = f.site; fmask = 1u << f.bit; }
            }
        }
        uint32_t crc = 0xFFFFu;
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t b = __ldg(msg + i);
            if (INJECT && fsite == len + i) b ^= fmask;
            crc = crc16_step(crc, b);
            if (INJECT && fsite == i) crc ^= fmask;
        }
        crc_vote_store<NC>(crc, static_cast<uint16_t*>(a.out), local, a.unit_base + local, valid, lane, a.flags, tally);
    }
    tally.flush(a.counters);
}

}  // namespace xmr

#define XMR_CRC_B64_KERNEL(NC, INJ)                                                                      \
    extern "C" __global__ void __launch_bounds__(XMR_CTA_THREADS)                                        \
    xmr_crc16_b64_nc##NC##_inj##INJ(const __grid_constant__ xmr_args a, const __grid_constant__ CUtensorMap tmap) { \
        xmr::crc16_b64_body<NC, INJ != 0>(a, &tmap);                                                     \
    }
#define XMR_CRC_GEN_KERNEL(NC, INJ)                                                                      \
    extern "C" __global__ void __launch_bounds__(XMR_CTA_THREADS)                                        \
    xmr_crc16_gen_nc##NC##_inj##INJ(const __grid_constant__ xmr_args a) {                                \
        xmr::crc16_gen_body<NC, INJ != 0>(a);                                                            \
    }
XMR_CRC_B64_KERNEL(1, 0) XMR_CRC_B64_KERNEL(2, 0) XMR_CRC_B64_KERNEL(3, 0)
XMR_CRC_B64_KERNEL(1, 1) XMR_CRC_B64_KERNEL(2, 1) XMR_CRC_B64_KERNEL(3, 1)
XMR_CRC_GEN_KERNEL(1, 0) XMR_CRC_GEN_KERNEL(2, 0) XMR_CRC_GEN_KERNEL(3, 0)
XMR_CRC_GEN_KERNEL(1, 1) XMR_CRC_GEN_KERNEL(2, 1) XMR_CRC_GEN_KERNEL(3, 1)
