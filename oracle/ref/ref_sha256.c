/* oracle/_ref/libref_sha256.so : tests/sha256_common/sha256_common_tmr.c (+ sha_data.inc,
 * and the 4000-byte KAT of tests/hifive1/sha256.tmr/sha_data.inc) from the reference tree. */
#define REF_WANT_FANOUT
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

/* the caller-visible scratch sha256_hash leaves behind (ctx_data / ctx_bitlen / ctx_state), for the entry-point test */
REF_API void ref_sha256_ctx(const uint8_t* msg, uint32_t len, uint8_t digest[32], uint8_t cd[64], uint32_t bl[2], uint32_t st[8]) {
    sha256_hash(cd, bl, st, (unsigned char*)msg, len, digest);
}

/* pthread fan-out for the CPU baseline ("reference" kind): sha256_hash() is a pure function of its arguments */
typedef struct { const uint8_t* in; uint8_t* out; uint32_t len, nc; int ce, cs; } sha_mt;
static void sha_shard(void* p, uint64_t u0, uint64_t n, ref_stats* st) {
    sha_mt* a = (sha_mt*)p;
    ref_sha256_xmr(a->in + u0 * a->len, a->out + u0 * 32, n, a->len, a->nc, a->ce, a->cs, NULL, st);
}
REF_API void ref_sha256_xmr_mt(const uint8_t* in, uint8_t* out, uint64_t n, uint32_t len, uint32_t nc,
                               int count_errors, int count_syncs, int n_threads, ref_stats* st) {
    sha_mt a = { in, out, len, nc, count_errors, count_syncs };
    ref_fanout(sha_shard, &a, n, n_threads, st);
}
