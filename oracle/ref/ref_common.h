/*
 * ref_common.h -- glue for oracle/_ref: the reference's OWN C sources compiled
 * where they lie under $(COAST_REF)/tests (never copied into this repo).
 * TEST INFRASTRUCTURE ONLY (see ../coast_oracle.h).
 *
 * The reference marks things with clang-only `__attribute__((annotate(..)))`
 * placed where gcc rejects attributes (e.g. `int checkGolden() __NO_xMR {`,
 * tests/matrixMultiply/matrixMultiply.c:115).  Pre-defining the include guard of
 * tests/COAST.h (:1-2) and giving every macro of :11-67 an empty / gcc-legal
 * body leaves the reference files untouched.
 */
#ifndef REF_COMMON_H_
#define REF_COMMON_H_
#if defined(REF_WANT_FANOUT) && !defined(_GNU_SOURCE)
#define _GNU_SOURCE            /* CPU_SET / pthread_setaffinity_np; must precede every system header */
#endif
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define __COAST_MACROS__
#define __NO_xMR
#define __xMR
#define __xMR_FN_CALL
#define __SKIP_FN_CALL
#define __DEFAULT_xMR int __xMR_DEFAULT_BEHAVIOR__;
#define __DEFAULT_NO_xMR int __xMR_DEFAULT_BEHAVIOR__;
#define __COAST_VOLATILE __attribute__((used))
#define __ISR_FUNC
#define __xMR_RET_VAL
#define __xMR_PROT_LIB
#define __xMR_ALL_AFTER_CALL
#define __COAST_IGNORE_GLOBAL(name)
#define __NO_xMR_ARG(num)
#define __COAST_NO_INLINE __attribute__((noinline))

#define REF_API __attribute__((visibility("default")))

/* xMR wrapper shared by the four harnesses: the SoR-exit vote of
 * projects/dataflowProtection/synchronization.cpp:512-529 (select voter),
 * :1391-1431 (error count), :1117-1192 (DWC compare).  `es` = element size of the
 * C type the reference stores; `nv` = elements per unit. */
typedef struct ref_stats { uint64_t errors_corrected, dwc_detected, syncs, injected, first_fault_unit; } ref_stats;

static inline void ref_vote_n(const uint8_t* const rep[3], uint32_t nc, uint32_t es, uint32_t nv, int count_errors,
                              int count_syncs, uint64_t unit, uint8_t* out, ref_stats* st) {
    int disagree = 0;
    if (nc == 1) { memcpy(out, rep[0], (size_t)es * nv); return; }
    if (nc == 2) {
        if (memcmp(rep[0], rep[1], (size_t)es * nv)) { disagree = 1; st->dwc_detected++; }
        memcpy(out, rep[0], (size_t)es * nv);
    } else {
        for (uint32_t e = 0; e < nv; ++e) {
            const uint8_t *r0 = rep[0] + e * es, *r1 = rep[1] + e * es, *r2 = rep[2] + e * es;
            int c01 = !memcmp(r0, r1, es), c02 = !memcmp(r0, r2, es);
            memcpy(out + e * es, c01 ? r0 : r2, es);
            if (!(c01 && c02)) { disagree = 1; if (count_errors) st->errors_corrected++; }
        }
        if (count_errors && count_syncs) st->syncs += nv;
    }
    if (disagree && unit < st->first_fault_unit) st->first_fault_unit = unit;
}
static inline void ref_vote(uint8_t rep[3][32], uint32_t nc, uint32_t es, uint32_t nv, int count_errors,
                            int count_syncs, uint64_t unit, uint8_t* out, ref_stats* st) {
    const uint8_t* const p[3] = { rep[0], rep[1], rep[2] };
    ref_vote_n(p, nc, es, nv, count_errors, count_syncs, unit, out, st);
}

/* A fault in ONE replica's private copy of its input (memory replication, rule D1,
 * docs/source/passes.rst:329): flip `bit` of byte `byte` of replica `replica`.
 * byte < 0 = none.  Mid-computation sites cannot be reached without editing the
 * reference sources, so _ref only covers input sites. */
typedef struct ref_fault { int replica; int byte; int bit; } ref_fault;


/* pthread fan-out shared by the "reference"-kind CPU baselines: contiguous shards of n units, thread t pinned to the t-th
 * CPU of the process's affinity mask (an unpinned 128-thread pass measured 245 .. 1000 MB/s on two boxes, r01), per-shard
 * stats merged exactly -- first_fault_unit is the minimum over shards of (shard-local index + shard start). */
#ifdef REF_WANT_FANOUT
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <pthread.h>
#include <sched.h>
typedef void (*ref_shard_fn)(void* ctx, uint64_t u0, uint64_t n, ref_stats* st);
typedef struct { ref_shard_fn fn; void* ctx; uint64_t u0, n; ref_stats st; int cpu; } ref_shard;
static void* ref_shard_main(void* p) {
    ref_shard* a = (ref_shard*)p;
    if (a->cpu >= 0) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(a->cpu, &one); pthread_setaffinity_np(pthread_self(), sizeof one, &one); }
    a->fn(a->ctx, a->u0, a->n, &a->st);
    return NULL;
}
static void ref_fanout(ref_shard_fn fn, void* ctx, uint64_t n, int n_threads, ref_stats* st) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    static pthread_t th[256]; static ref_shard a[256];
    int cpus[1024], n_cpus = 0;
    cpu_set_t cur;
    if (!sched_getaffinity(0, sizeof cur, &cur))
        for (int c = 0; c < CPU_SETSIZE && n_cpus < 1024; ++c) if (CPU_ISSET(c, &cur)) cpus[n_cpus++] = c;
    uint64_t per = (n + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t u0 = per * (uint64_t)t; if (u0 > n) u0 = n;
        uint64_t u1 = u0 + per; if (u1 > n) u1 = n;
        a[t].fn = fn; a[t].ctx = ctx; a[t].u0 = u0; a[t].n = u1 - u0; a[t].cpu = n_cpus ? cpus[t % n_cpus] : -1;
        memset(&a[t].st, 0, sizeof(ref_stats)); a[t].st.first_fault_unit = ~(uint64_t)0;
        pthread_create(&th[t], NULL, ref_shard_main, &a[t]);
    }
    if (!st->errors_corrected && !st->dwc_detected && !st->syncs && !st->injected && !st->first_fault_unit)
        st->first_fault_unit = ~(uint64_t)0;               /* a zero-initialised caller struct means "no fault yet" */
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        st->errors_corrected += a[t].st.errors_corrected; st->dwc_detected += a[t].st.dwc_detected;
        st->syncs += a[t].st.syncs; st->injected += a[t].st.injected;
        if (a[t].st.first_fault_unit != ~(uint64_t)0 && a[t].st.first_fault_unit + a[t].u0 < st->first_fault_unit)
            st->first_fault_unit = a[t].st.first_fault_unit + a[t].u0;
    }
}
#endif

#endif
