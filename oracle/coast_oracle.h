/*
 * coast_oracle.h -- CPU oracle for the COAST protected-region hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under coast_b200/ or include/ may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and there only as the checker or the
 * timed CPU baseline.
 *
 * It is a plain-C restatement of
 *   - the four protected workloads (tests/crc16, tests/sha256_common, tests/aes,
 *     tests/mm_common + tests/matrixMultiply of byuccl/coast), and
 *   - what the code emitted by `opt -TMR / -DWC [-countErrors] [-countSyncs]`
 *     does at run time (projects/dataflowProtection/synchronization.cpp), and
 *   - the campaign's fault model (simulation/platform/resources/injector.py:202-207).
 *
 * Parity pinning: the zero-fault arithmetic is pinned by every golden vector the
 * reference's own tests hold (crc16 0x5ba3, SHA 10-B and 4000-B KATs, 568 NIST
 * AES vectors, 9x9 mm xor_golden) and by oracle/_ref (the reference's C sources
 * compiled where they lie).  The behaviour UNDER FAULTS is pinned by nothing in
 * the reference (no test asserts a TMR_ERROR_CNT value; the real pass needs
 * LLVM 7 which is absent): for that part this oracle is "parity unpinned" --
 * its fidelity is argued from the cited source lines.
 */
#ifndef COAST_ORACLE_H_
#define COAST_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_K_CRC16 = 0, ORC_K_SHA256 = 1, ORC_K_AES128 = 2, ORC_K_MM_U32 = 3, ORC_K_GEMM_TF32 = 4, ORC_K_QSORT = 5,
       ORC_K_CHSTONE_SHA = 6, ORC_K_CHSTONE_AES = 7, ORC_K_COUNT_ = 8 };
enum { ORC_F_COUNT_ERRORS = 1, ORC_F_COUNT_SYNCS = 2, ORC_F_NO_MEM_REPLICATION = 4, ORC_F_MAJORITY = 0x100,
       ORC_F_STORE_DATA_SYNC = 0x200, ORC_F_NO_STORE_DATA_SYNC = 0x400, ORC_F_NO_LOAD_SYNC = 0x800, ORC_F_NO_STORE_ADDR_SYNC = 0x1000 };
/* In-loop store votes (rule C4): -storeDataSync forces them, -noMemReplication needs them (one memory copy: stores are voted,
 * synchronization.cpp:205-215), -noStoreDataSync removes them (:333-335).  Modelled for CRC16, MM_U32 and SHA256 (orc_store_votes_supported). */
int orc_store_votes(uint32_t flags);
int orc_store_votes_supported(uint32_t kernel);
enum { ORC_PLAN_NONE = 0, ORC_PLAN_BERNOULLI = 1, ORC_PLAN_TABLE = 2 };
enum { ORC_AES_DECRYPT = 1, ORC_AES_KEY_PER_UNIT = 2 };

typedef struct orc_plan {
    uint32_t mode, seed_lo, seed_hi, threshold;
    const uint32_t* table; /* host pointer, n_units entries, indexed by LOCAL unit */
} orc_plan;

typedef struct orc_fault { int active; uint32_t replica, site, bit; } orc_fault;

typedef struct orc_stats {
    uint64_t errors_corrected, dwc_detected, syncs, injected, first_fault_unit;
} orc_stats;

typedef struct orc_desc {
    uint32_t kernel, num_clones, flags, mode;
    uint64_t n_units, unit_base;
    uint32_t unit_bytes, M, N, K;
    const void* in; void* out; const void* aux;
    uint8_t key[16];
    const orc_plan* plan;
} orc_desc;

/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 kat_vectors pin it). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* dst[i] = philox(ctr={(word_base+i)/4,0,0,0}, key={seed,0})[(word_base+i)%4] */
void orc_fill_philox(uint32_t* dst, uint64_t n_words, uint64_t word_base, uint32_t seed);

uint32_t orc_fault_sites(uint32_t kernel, uint32_t unit_bytes, uint32_t K);
uint32_t orc_fault_site_bits(uint32_t kernel, uint32_t unit_bytes, uint32_t K, uint32_t site);
uint32_t orc_out_bytes_per_unit(uint32_t kernel);
uint32_t orc_out_bytes(uint32_t kernel, uint32_t unit_bytes);   /* QSORT: unit_bytes (the sorted array) */
uint32_t orc_votes_per_unit(uint32_t kernel);
/* Decide the fault (if any) of GLOBAL unit `unit` (local index unit - unit_base for TABLE). */
void orc_fault_for_unit(const orc_plan* plan, uint32_t kernel, uint32_t num_clones, uint32_t unit_bytes,
                        uint32_t K, uint64_t unit, uint64_t local, orc_fault* f);

/* Single-replica restatements (fault may be NULL). */
uint16_t orc_crc16(const uint8_t* data, uint32_t len, const orc_fault* f);
void     orc_sha256(const uint8_t* data, uint32_t len, uint8_t digest[32], const orc_fault* f);
void     orc_aes128(uint8_t state[16], uint8_t key[16], int dir, const orc_fault* f);
void     orc_chstone_sha(const uint8_t* data, uint32_t len, uint32_t digest[5], const orc_fault* f);   /* len % 64 == 0 */
void     orc_chstone_aes(int32_t statemt[16], const int32_t key[16], int dir, const orc_fault* f);     /* type 128128 */
uint32_t orc_mm_u32_elem(const uint32_t* A, const uint32_t* B, uint32_t K, uint32_t N, uint32_t i, uint32_t j,
                         const orc_fault* f);
float    orc_gemm_tf32_elem(const float* A, const float* B, uint32_t K, uint32_t N, uint32_t i, uint32_t j,
                            const orc_fault* f);

/* The protected region: num_clones replicas per unit + SoR-exit vote/compare/count.
 * Returns 0, or -1 on a bad descriptor. `stats` accumulates (caller zeroes it). */
int orc_run(const orc_desc* d, orc_stats* stats);
/* Same over [u0, u1) only (for multi-process / multi-thread CPU baselines). */
int orc_run_range(const orc_desc* d, uint64_t u0, uint64_t u1, orc_stats* stats);
/* pthread fan-out over n_threads disjoint ranges; stats are summed (min for first_fault_unit). */
int orc_run_mt(const orc_desc* d, int n_threads, orc_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
