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

static inline void ref_vote(uint8_t rep[3][32], uint32_t nc, uint32_t es, uint32_t nv, int count_errors,
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

/* A fault in ONE replica's private copy of its input (memory replication, rule D1,
 * docs/source/passes.rst:329): flip `bit` of byte `byte` of replica `replica`.
 * byte < 0 = none.  Mid-computation sites cannot be reached without editing the
 * reference sources, so _ref only covers input sites. */
typedef struct ref_fault { int replica; int byte; int bit; } ref_fault;

#endif
