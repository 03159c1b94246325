/* oracle/_ref/libref_sha256.so : tests/sha256_common/sha256_common_tmr.c (+ sha_data.inc,
 * and the 4000-byte KAT of tests/hifive1/sha256.tmr/sha_data.inc) from the reference tree. */
#include "ref_common.h"
typedef uint32_t mm_t;
unsigned error;
#include "sha256_common/sha_data.inc"
#include "sha256_common/sha256_common_tmr.c"

/* the shipped 10-byte KAT, exactly as sha256_tmr.c:17-34 drives it */
REF_API unsigned ref_sha256_kat10(void) { sha_run_test(); return checkGolden(); }
REF_API const uint8_t* ref_sha256_kat10_msg(void) { return hash_data; }
REF_API const uint8_t* ref_sha256_kat10_golden(void) { return golden; }

REF_API void ref_sha256(const uint8_t* msg, uint32_t len, uint8_t digest[32]) {
    unsigned char cd[64]; uint32_t bl[2], st[8];
    sha256_hash(cd, bl, st, (unsigned char*)msg, len, digest);
}

REF_API void ref_sha256_xmr(const uint8_t* in, uint8_t* out, uint64_t n, uint32_t len, uint32_t nc,
                            int count_errors, int count_syncs, const ref_fault* faults, ref_stats* stt) {
    /* every replica owns its ctx_data / ctx_bitlen / ctx_state / data copies (cloneGlobals,
     * projects/dataflowProtection/cloning.cpp:2417-2462) */
    unsigned char cd[3][64]; uint32_t bl[3][2], st[3][8];
    uint8_t* priv = (uint8_t*)malloc((size_t)len * 3 + 1);
    uint8_t rep[3][32];
    for (uint64_t u = 0; u < n; ++u) {
        for (uint32_t r = 0; r < nc; ++r) {
            uint8_t* p = priv + (size_t)r * len;
            memcpy(p, in + u * len, len);
            if (faults && faults[u].byte >= 0 && faults[u].replica == (int)r) {
                p[faults[u].byte] ^= (uint8_t)(1u << faults[u].bit);
                stt->injected++;
            }
            sha256_hash(cd[r], bl[r], st[r], p, len, rep[r]);
        }
        ref_vote(rep, nc, 1, 32, count_errors, count_syncs, u, out + u * 32, stt);
    }
    free(priv);
}

/* pthread fan-out for the CPU baseline ("reference" kind) */
#include <pthread.h>
typedef struct { const uint8_t* in; uint8_t* out; uint64_t n; uint32_t len, nc; int ce, cs; ref_stats st; } sha_mt;
static void* sha_mt_main(void* p) {
    sha_mt* a = (sha_mt*)p;
    ref_sha256_xmr(a->in, a->out, a->n, a->len, a->nc, a->ce, a->cs, NULL, &a->st);
    return NULL;
}
REF_API void ref_sha256_xmr_mt(const uint8_t* in, uint8_t* out, uint64_t n, uint32_t len, uint32_t nc,
                               int count_errors, int count_syncs, int n_threads, ref_stats* st) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; sha_mt a[256];
    uint64_t per = (n + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t u0 = per * (uint64_t)t; if (u0 > n) u0 = n;
        uint64_t u1 = u0 + per; if (u1 > n) u1 = n;
        a[t].in = in + u0 * len; a[t].out = out + u0 * 32; a[t].n = u1 - u0; a[t].len = len; a[t].nc = nc;
        a[t].ce = count_errors; a[t].cs = count_syncs; memset(&a[t].st, 0, sizeof(ref_stats));
        a[t].st.first_fault_unit = ~(uint64_t)0;
        pthread_create(&th[t], NULL, sha_mt_main, &a[t]);
    }
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        st->errors_corrected += a[t].st.errors_corrected; st->dwc_detected += a[t].st.dwc_detected;
        st->syncs += a[t].st.syncs; st->injected += a[t].st.injected;
    }
}
