// xmr_mm.cuh -- protected integer matrix multiply (tests/mm_common/mm_common_tmr.c:3-20 and
// tests/matrixMultiply/matrixMultiply.c:95-112 of byuccl/coast), exact modulo 2^32.
//
// Unit = one element r[i][j] = sum_k f[i][k]*s[k][j].  The reference accumulates in an
// `unsigned long` and truncates on the store (:16 / :108); only the low 32 bits are observable,
// so each replica keeps `sum` mod 2^32.  SoR exit = that element store: ONE mm_t vote per unit.
// Fault sites: s in [0,K): `sum` after k-step s (32 bits).
//
// Layout: NC replica lanes per element, (32/NC) consecutive j per warp -> B rows are read
// coalesced, the A element is a warp-broadcast; replicas of an element share every load
// (same address -> one L1 transaction).
#pragma once
#include "xmr_common.cuh"

namespace xmr {

template <int NC, bool INJECT>
__device__ __forceinline__ void mm_u32_body(const xmr_args& a) {
    constexpr int UPW = Lanes<NC>::kUnitsPerWarp;
    const int lane = threadIdx.x & 31;
    const int r = Lanes<NC>::replica(lane);
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long n_wtiles = (a.n_units + UPW - 1) / UPW;
    const uint32_t* __restrict__ A = static_cast<const uint32_t*>(a.in);
    const uint32_t* __restrict__ B = static_cast<const uint32_t*>(a.aux);
    uint32_t* C = static_cast<uint32_t*>(a.out);
    const uint32_t K = a.K, N = a.N;
    Tally tally(a);
    for (unsigned long long wt = gwarp; wt < n_wtiles; wt += nwarps) {
        const unsigned long long local = wt * UPW + Lanes<NC>::unit(lane);
        const bool valid = local < a.n_units;
        const unsigned long long e = valid ? local : 0ull;
        const uint32_t i = (uint32_t)(e / N), j = (uint32_t)(e % N);
        uint32_t fsite = 0xFFFFFFFFu, fmask = 0u;
        if (INJECT) {
            Fault f = fault_for_unit(a, NC, e, [](uint32_t) { return 32u; });
            if (f.active && valid) {
                if (Lanes<NC>::voter(lane)) tally.injected++;
                if ((int)f.replica == r) { fsite = f.site; fmask = 1u << f.bit; }
            }
        }
        const uint32_t* ap = A + (size_t)i * K;
        const uint32_t* bp = B + j;
        uint32_t sum = 0;
        if (!(a.flags & XMR_F_STORE_VOTES)) {
            for (uint32_t k = 0; k < K; ++k) {                  // :12-14
                sum += __ldg(ap + k) * __ldg(bp + (size_t)k * N);
                if (INJECT && fsite == k) sum ^= fmask;
            }
            Voted v = vote_u32<NC, 4>(sum, a.flags & COAST_F_MAJORITY_D);
            if (valid && Lanes<NC>::voter(lane)) {
                C[local] = v.vote;                              // :16
                tally.unit_exit<NC>(v.bad, 1u, a.flags, a.unit_base + local);
            }
        } else {
            // -storeDataSync / -noMemReplication: `sum += ...` (:13) is voted at every k, the replicas continue with the voted
            // value; K votes + the SoR-exit store (:16)
            const bool majority = a.flags & COAST_F_MAJORITY_D;
            uint32_t bad = 0;
            for (uint32_t k = 0; k < K; ++k) {
                sum += __ldg(ap + k) * __ldg(bp + (size_t)k * N);
                bad += store_vote<NC>(sum, lane, majority);
                if (INJECT && fsite == k) sum ^= fmask;
            }
            bad += store_vote<NC>(sum, lane, majority);
            if (valid && Lanes<NC>::voter(lane)) {
                C[local] = sum;
                tally.unit_exit<NC>(bad, K + 1u, a.flags, a.unit_base + local);
            }
        }
    }
    tally.flush(a.counters);
}

}  // namespace xmr

#define XMR_MM_KERNEL(NC, INJ)                                                                           \
    extern "C" __global__ void __launch_bounds__(XMR_CTA_THREADS)                                        \
    xmr_mm_u32_nc##NC##_inj##INJ(const __grid_constant__ xmr_args a) { xmr::mm_u32_body<NC, INJ != 0>(a); }
XMR_MM_KERNEL(1, 0) XMR_MM_KERNEL(2, 0) XMR_MM_KERNEL(3, 0)
XMR_MM_KERNEL(1, 1) XMR_MM_KERNEL(2, 1) XMR_MM_KERNEL(3, 1)
