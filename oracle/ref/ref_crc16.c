/* oracle/_ref/libref_crc16.so : tests/crc16/crc16.c compiled from the reference tree. */
#include "ref_common.h"
#define main ref_crc16_main
#include "crc16/crc16.c"
#undef main

REF_API int ref_crc16_run_main(void) { return ref_crc16_main(); }   /* prints "result: 5ba3" */
REF_API unsigned short ref_crc16(const unsigned char* p, unsigned char len) { return crc16(p, len); }

/* n units of `len` bytes each -> n u16, under nc replicas; faults[u] optional (NULL = none). */
REF_API void ref_crc16_xmr(const uint8_t* in, uint16_t* out, uint64_t n, uint32_t len, uint32_t nc,
                           int count_errors, int count_syncs, const ref_fault* faults, ref_stats* st) {
    uint8_t priv[3][256];
    uint8_t rep[3][32];
    for (uint64_t u = 0; u < n; ++u) {
        for (uint32_t r = 0; r < nc; ++r) {
            memcpy(priv[r], in + u * len, len);
            if (faults && faults[u].byte >= 0 && faults[u].replica == (int)r) {
                priv[r][faults[u].byte] ^= (uint8_t)(1u << faults[u].bit);
                if (r == (uint32_t)faults[u].replica) st->injected++;
            }
            unsigned short c = crc16(priv[r], (unsigned char)len);
            memcpy(rep[r], &c, 2);
        }
        ref_vote(rep, nc, 2, 1, count_errors, count_syncs, u, (uint8_t*)(out + u), st);
    }
}

/* pthread fan-out for the CPU baseline ("reference" kind): crc16() is a pure function of its arguments */
#include <pthread.h>
typedef struct { const uint8_t* in; uint16_t* out; uint64_t n; uint32_t len, nc; int ce, cs; ref_stats st; } crc_mt;
static void* crc_mt_main(void* p) {
    crc_mt* a = (crc_mt*)p;
    ref_crc16_xmr(a->in, a->out, a->n, a->len, a->nc, a->ce, a->cs, NULL, &a->st);
    return NULL;
}
REF_API void ref_crc16_xmr_mt(const uint8_t* in, uint16_t* out, uint64_t n, uint32_t len, uint32_t nc,
                              int count_errors, int count_syncs, int n_threads, ref_stats* st) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; crc_mt a[256];
    uint64_t per = (n + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t u0 = per * (uint64_t)t; if (u0 > n) u0 = n;
        uint64_t u1 = u0 + per; if (u1 > n) u1 = n;
        a[t].in = in + u0 * len; a[t].out = out + u0; a[t].n = u1 - u0; a[t].len = len; a[t].nc = nc;
        a[t].ce = count_errors; a[t].cs = count_syncs; memset(&a[t].st, 0, sizeof(ref_stats));
        a[t].st.first_fault_unit = ~(uint64_t)0;
        pthread_create(&th[t], NULL, crc_mt_main, &a[t]);
    }
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        st->errors_corrected += a[t].st.errors_corrected; st->dwc_detected += a[t].st.dwc_detected;
        st->syncs += a[t].st.syncs; st->injected += a[t].st.injected;
    }
}
