// tc_probe.cu -- standalone bring-up probe for tcgen05.mma kind::tf32 (single CTA, one 128x128x32 tile).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r_, &s); printf("%s failed: %s\n", #x, s); exit(1); } } while (0)
#define RK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// mode 0: B row-major [K][N] via 3-D map (MN-major descriptor)   mode 1: B^T [N][K] via 2-D map (K-major descriptor)
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, float* C, uint32_t* info, int mode) {
    extern __shared__ uint8_t dyn[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem; uint8_t* sB = smem + 16384;
    uint64_t* full = (uint64_t*)(smem + 32768); uint64_t* done = full + 1; uint32_t* slot = (uint32_t*)(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = *slot;
    if (threadIdx.x == 0) info[0] = tb;

    // --- TMEM st/ld self test by warps 4..7: write lane*1000+col into columns 256..287, read back
    if (warp >= 4) {
        const int q = warp & 3;
        uint32_t taddr = tb + ((uint32_t)(q * 32) << 16) + 256;
        for (int c = 0; c < 4; ++c) {
            uint32_t v = (q * 32 + lane) * 1000 + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 3 && q == 1) { info[1] = r0; info[2] = r1; info[3] = r2; info[4] = r3; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == 0 && lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full)), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sA)), "l"(&ma), "r"(s32(full)), "r"(0), "r"(0) : "memory");
        if (mode == 2 || mode == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(s32(sB)), "l"(&mb), "r"(s32(full)), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        else if (mode == 0 || mode == 3 || mode == 5 || mode == 6)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(sB)), "l"(&mb), "r"(s32(full)), "r"(0), "r"(0), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sB)), "l"(&mb), "r"(s32(full)), "r"(0), "r"(0) : "memory");
    } else if (warp == 1) {
        mbar_wait(full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((mode != 1 ? 1u : 0u) << 16) | (16u << 17) | (8u << 24);
            for (int k = 0; k < 4; ++k) {
                uint64_t da = sdesc(s32(sA) + k * 32, 16, 1024, 2);
                uint64_t db;
                if (mode == 0) db = sdesc(s32(sB) + k * 1024, 4096, 1024, 2);        // [chunk][k][128B], LBO=chunk stride, SBO=kgroup stride
                else if (mode == 3) db = sdesc(s32(sB) + k * 1024, 1024, 4096, 2);   // same smem, LBO/SBO swapped
                else if (mode == 2) db = sdesc(s32(sB) + k * 4096, 1024, 4096, 2);   // canonical [kgroup][chunk][8][128B]
                else if (mode == 4) db = sdesc(s32(sB) + k * 4096, 4096, 1024, 2);   // canonical smem, swapped
                else if (mode == 5) db = sdesc(s32(sB) + k * 1024, 4096, 512, 1);    // SW128_BASE32B: atom = 4 k-rows x 128 B
                else if (mode == 6) db = sdesc(s32(sB) + k * 1024, 512, 4096, 1);
                else db = sdesc(s32(sB) + k * 32, 16, 1024, 2);
                uint32_t acc = k ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(done)) : "memory");
        }
        __syncwarp();
    } else if (warp >= 4) {
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        for (int c0 = 0; c0 < 128; c0 += 4) {
            uint32_t r0, r1, r2, r3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(tb + ((uint32_t)(q * 32) << 16) + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* dst = C + (q * 32 + lane) * 128 + c0;
            dst[0] = __uint_as_float(r0); dst[1] = __uint_as_float(r1); dst[2] = __uint_as_float(r2); dst[3] = __uint_as_float(r3);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
    RK(cudaSetDevice(0));
    RK(cudaFree(0));
    const int M = 128, N = 128, K = 32;
    std::vector<float> A(M * K), B(K * N), Bt(N * K), C(M * N), ref(M * N);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) A[i * K + k] = (float)((i * 3 + k) % 7) - 3.0f;
    for (int k = 0; k < K; ++k) for (int j = 0; j < N; ++j) { B[k * N + j] = (float)((k * 5 + j) % 5) - 2.0f; Bt[j * K + k] = B[k * N + j]; }
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[k * N + j]; ref[i * N + j] = (float)s; }
    float *dA, *dB, *dBt, *dC; uint32_t* dinfo;
    RK(cudaMalloc(&dA, A.size() * 4)); RK(cudaMalloc(&dB, B.size() * 4)); RK(cudaMalloc(&dBt, Bt.size() * 4)); RK(cudaMalloc(&dC, C.size() * 4)); RK(cudaMalloc(&dinfo, 64));
    RK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); RK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice)); RK(cudaMemcpy(dBt, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap ma, mb3, mbt, mb4, mb5;
    { cuuint64_t gd[3] = {32, K, N / 32}; cuuint64_t gs[2] = {N * 4, 128}; cuuint32_t bx[3] = {32, 32, 4}; cuuint32_t es[3] = {1, 1, 1};
      CK(cuTensorMapEncodeTiled(&mb5, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
    { cuuint64_t gd[4] = {32, 8, N / 32, K / 8}; cuuint64_t gs[3] = {N * 4, 128, 8 * N * 4}; cuuint32_t bx[4] = {32, 8, 4, 4}; cuuint32_t es[4] = {1, 1, 1, 1};
      CK(cuTensorMapEncodeTiled(&mb4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
    { cuuint64_t gd[2] = {K, M}; cuuint64_t gs[1] = {K * 4}; cuuint32_t bx[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
      CK(cuTensorMapEncodeTiled(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
    { cuuint64_t gd[3] = {32, K, N / 32}; cuuint64_t gs[2] = {N * 4, 128}; cuuint32_t bx[3] = {32, 32, 4}; cuuint32_t es[3] = {1, 1, 1};
      CK(cuTensorMapEncodeTiled(&mb3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
    { cuuint64_t gd[2] = {K, N}; cuuint64_t gs[1] = {K * 4}; cuuint32_t bx[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
      CK(cuTensorMapEncodeTiled(&mbt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dBt, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
    RK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
    for (int mode = 0; mode < 7; ++mode) {
        RK(cudaMemset(dC, 0xff, C.size() * 4)); RK(cudaMemset(dinfo, 0, 64));
        probe<<<1, 256, 40000>>>(ma, mode == 1 ? mbt : ((mode == 2 || mode == 4) ? mb4 : (mode >= 5 ? mb5 : mb3)), dC, dinfo, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: sync = %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        uint32_t info[16]; RK(cudaMemcpy(info, dinfo, 64, cudaMemcpyDeviceToHost)); RK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        printf("  tmem_base=0x%x st/ld selftest (expect 35000..35003): %u %u %u %u\n", info[0], info[1], info[2], info[3], info[4]);
        double maxerr = 0; int nz = 0;
        for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)C[i] - ref[i])); nz += C[i] != 0.0f; }
        printf("  nonzero=%d maxerr=%g  C[0][0..3]=%g %g %g %g  ref=%g %g %g %g\n", nz, maxerr, C[0], C[1], C[2], C[3], ref[0], ref[1], ref[2], ref[3]);
        printf("  C[1][0..3]=%g %g %g %g  ref=%g %g %g %g\n", C[128], C[129], C[130], C[131], ref[128], ref[129], ref[130], ref[131]);
    }
    return 0;
}
