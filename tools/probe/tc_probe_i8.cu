// tc_probe_i8.cu -- bring-up probe for tcgen05.mma kind::i8 with UNSIGNED 8-bit operands and s32 accumulate:
// (1) K-major SW128 descriptors with 128-byte K rows, 4 MMAs of K=32 per row; (2) N=32 tiles; (3) does the s32
// accumulator WRAP (needed for exact arithmetic modulo 2^32) or saturate when the sum exceeds 2^31?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe_i8 tc_probe_i8.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r_, &s); printf("%s failed: %s\n", #x, s); exit(1); } } while (0)
#define RK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

constexpr int BN = 32;
// A: [128][K] u8 K-major, B^T: [BN][K] u8 K-major; C[128][BN] s32 = sum_k A*B, accumulated over K/128 k-blocks
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, uint32_t* C, int kblocks, int fmt, int a_in_tmem) {
    extern __shared__ uint8_t dyn[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem; uint8_t* sB = smem + 16384;
    uint64_t* full = (uint64_t*)(smem + 16384 + 4096); uint64_t* empty = full + 1; uint64_t* done = full + 2; uint32_t* slot = (uint32_t*)(full + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(empty)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = *slot;
    if (warp == 0 && lane == 0) {
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(empty, (kb & 1) ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full)), "r"(16384u + BN * 128u) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sA)), "l"(&ma), "r"(s32(full)), "r"(kb * 128), "r"(0) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sB)), "l"(&mb), "r"(s32(full)), "r"(kb * 128), "r"(0) : "memory");
        }
    } else if (warp == 1 && lane == 0) {
        // idesc: c_format S32 (2) [4,6); a/b format 0 = unsigned 8 bit, 1 = signed [7,10)/[10,13); K-major both; N>>3 [17,23); M>>4 [24,29)
        const uint32_t idesc = (2u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(BN >> 3) << 17) | (8u << 24);
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(full, kb & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (a_in_tmem) {
                // stage the whole 128 x 128-byte A tile into TMEM columns [32, 64): 4 copies of 128 rows x 256 bits (one per K=32 step)
                for (int k = 0; k < 4; ++k) {
                    uint64_t da = sdesc(s32(sA) + k * 32, 16, 1024, 2);
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + 32 + k * 8), "l"(da) : "memory");
                }
            }
            for (int k = 0; k < 4; ++k) {
                uint64_t da = sdesc(s32(sA) + k * 32, 16, 1024, 2), db = sdesc(s32(sB) + k * 32, 16, 1024, 2);
                uint32_t acc = (kb | k) ? 1u : 0u;
                if (a_in_tmem)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tb), "r"(tb + 32 + k * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(empty)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(done)) : "memory");
    } else if (warp >= 4) {
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        for (int c0 = 0; c0 < BN; c0 += 4) {
            uint32_t r0, r1, r2, r3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(tb + ((uint32_t)(q * 32) << 16) + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t* dst = C + (q * 32 + lane) * BN + c0;
            dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb) : "memory");
}

int main() {
    RK(cudaSetDevice(0)); RK(cudaFree(0));
    const int M = 128;
    for (int test = 0; test < 6; ++test) {
        const int a_in_tmem = test >= 3;
        const int K = test % 3 == 0 ? 128 : (test % 3 == 1 ? 512 : 128 * 300);      // test 2: 255*255*38400 = 2.5e9 > 2^31 -> wrap or saturate?
        std::vector<uint8_t> A((size_t)M * K), Bt((size_t)BN * K);
        std::vector<uint32_t> C(M * BN), ref(M * BN);
        for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) A[(size_t)i * K + k] = test % 3 == 2 ? 255 : (uint8_t)((i * 7 + k * 13 + 200) & 0xFF);
        for (int j = 0; j < BN; ++j) for (int k = 0; k < K; ++k) Bt[(size_t)j * K + k] = test % 3 == 2 ? (uint8_t)(255 - (j & 1)) : (uint8_t)((j * 11 + k * 3 + 130) & 0xFF);
        for (int i = 0; i < M; ++i) for (int j = 0; j < BN; ++j) { uint32_t s = 0; for (int k = 0; k < K; ++k) s += (uint32_t)A[(size_t)i * K + k] * Bt[(size_t)j * K + k]; ref[i * BN + j] = s; }
        uint8_t *dA, *dB; uint32_t* dC;
        RK(cudaMalloc(&dA, A.size())); RK(cudaMalloc(&dB, Bt.size())); RK(cudaMalloc(&dC, C.size() * 4));
        RK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); RK(cudaMemcpy(dB, Bt.data(), Bt.size(), cudaMemcpyHostToDevice));
        CUtensorMap ma, mb;
        { cuuint64_t gd[2] = {(cuuint64_t)K, M}; cuuint64_t gs[1] = {(cuuint64_t)K}; cuuint32_t bx[2] = {128, 128}; cuuint32_t es[2] = {1, 1};
          CK(cuTensorMapEncodeTiled(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
        { cuuint64_t gd[2] = {(cuuint64_t)K, BN}; cuuint64_t gs[1] = {(cuuint64_t)K}; cuuint32_t bx[2] = {128, BN}; cuuint32_t es[2] = {1, 1};
          CK(cuTensorMapEncodeTiled(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
        RK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        RK(cudaMemset(dC, 0xff, C.size() * 4));
        probe<<<1, 256, 32768>>>(ma, mb, dC, K / 128, 0, a_in_tmem);
        cudaError_t e = cudaDeviceSynchronize();
        printf("test %d (K=%d, A %s): sync = %s\n", test, K, a_in_tmem ? "in TMEM via tcgen05.cp" : "from smem", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        RK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0; for (int i = 0; i < M * BN; ++i) bad += C[i] != ref[i];
        printf("  mismatches vs exact mod 2^32: %d   C[0]=%u ref=%u  C[1]=%u ref=%u  C[last]=%u ref=%u  (0x7fffffff=%u)\n", bad, C[0], ref[0], C[1], ref[1], C[M * BN - 1], ref[M * BN - 1], 0x7fffffffu);
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
    }
    return 0;
}
