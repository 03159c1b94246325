/* oracle/_ref/libref_chaes.so : tests/chstone/aes/{aes,aes_enc,aes_dec,aes_func,aes_key}.c compiled from the reference tree. */
#include "ref_common.h"
#include <stdio.h>
#include <stdarg.h>
/* encrypt()/decrypt() print the message (aes_enc.c:127-133, aes_dec.c:127-133): keep the harness quiet unless asked */
static int chaes_quiet = 1;
static int chaes_printf(const char* fmt, ...) {
    if (chaes_quiet) return 0;
    va_list ap; va_start(ap, fmt); int n = vprintf(fmt, ap); va_end(ap); return n;
}
#define printf chaes_printf
#define main ref_chaes_main
#include "chstone/aes/aes.c"
#include "chstone/aes/aes_key.c"
#include "chstone/aes/aes_func.c"
#include "chstone/aes/aes_enc.c"
#include "chstone/aes/aes_dec.c"
#undef main
#undef printf

/* the benchmark as shipped: prints both messages and RESULT: PASS; returns main()'s value */
REF_API int ref_chaes_run_main(void) { chaes_quiet = 0; int rc = ref_chaes_main(); chaes_quiet = 1; return rc; }
/* one block through the reference's own encrypt()/decrypt() (type 128128); returns what the call added to main_result, i.e.
 * the number of bytes that differ from the benchmark's built-in expected vector */
REF_API int ref_chaes(int st[32], int k[32], int dir) {
    int before = main_result;
    if (dir) decrypt(st, k, 128128); else encrypt(st, k, 128128);
    return main_result - before;
}
/* n blocks (16 ints each) under nc replicas; keys: 16 ints, shared or per unit; faults[u] optional (NULL = none): a flip in the
 * replica's private copy of the input block */
REF_API void ref_chaes_xmr(const int32_t* in, int32_t* out, uint64_t n, const int32_t* keys, int key_per_unit, int dir, uint32_t nc,
                           int count_errors, int count_syncs, const ref_fault* faults, ref_stats* st) {
    for (uint64_t u = 0; u < n; ++u) {
        int32_t rep32[3][32];
        uint8_t rep[3][64];
        for (uint32_t r = 0; r < nc; ++r) {
            int k32[32];
            memset(rep32[r], 0, sizeof rep32[r]); memset(k32, 0, sizeof k32);
            memcpy(rep32[r], in + u * 16, 64);
            memcpy(k32, keys + (key_per_unit ? u * 16 : 0), 64);
            if (faults && faults[u].byte >= 0 && faults[u].replica == (int)r) {
                rep32[r][faults[u].byte] ^= (1 << faults[u].bit);
                st->injected++;
            }
            ref_chaes(rep32[r], k32, dir);
            memcpy(rep[r], rep32[r], 64);
        }
        const uint8_t* const p[3] = { rep[0], rep[1], rep[2] };
        ref_vote_n(p, nc, 4, 16, count_errors, count_syncs, u, (uint8_t*)(out + u * 16), st);   /* 16 `int` elements (aes_enc.c:130-131) */
    }
}
