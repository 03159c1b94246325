/* oracle/_ref/libref_crc16.so : tests/crc16/crc16.c compiled from the reference tree. */
#define REF_WANT_FANOUT
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
typedef struct { const uint8_t* in; uint16_t* out; uint32_t len, nc; int ce, cs; } crc_mt;
static void crc_shard(void* p, uint64_t u0, uint64_t n, ref_stats* st) {
    crc_mt* a = (crc_mt*)p;
    ref_crc16_xmr(a->in + u0 * a->len, a->out + u0, n, a->len, a->nc, a->ce, a->cs, NULL, st);
}
REF_API void ref_crc16_xmr_mt(const uint8_t* in, uint16_t* out, uint64_t n, uint32_t len, uint32_t nc,
                              int count_errors, int count_syncs, int n_threads, ref_stats* st) {
    crc_mt a = { in, out, len, nc, count_errors, count_syncs };
    ref_fanout(crc_shard, &a, n, n_threads, st);
}
