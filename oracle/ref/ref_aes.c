/* oracle/_ref/libref_aes.so : tests/aes/TI_aes_128.c + tests/aes/aes.c (568 NIST KATs). */
#define REF_WANT_FANOUT
#include "ref_common.h"
#include "aes/TI_aes_128.c"
#define main ref_aes_main
#include "aes/aes.c"
#undef main

REF_API int ref_aes_kat_errors(void) { local_errors = 0; aes_test(); return local_errors; }  /* aes.c:29-103 */
REF_API void ref_aes_enc_dec(unsigned char* state, unsigned char* key, unsigned char dir) { aes_enc_dec(state, key, dir); }

/* the four KAT tables as (ptr,count) so tests can replay them through the GPU path; 80-byte
 * records key|key2|cipher|plain|input (aes.c:62-66) */
REF_API const unsigned char* ref_aes_kat_table(int k, unsigned* count) {
    switch (k) {
    case 0: *count = ECBGFSbox128_count; return ECBGFSbox128;
    case 1: *count = ECBKeySbox128_count; return ECBKeySbox128;
    case 2: *count = ECBVarKey128_count; return ECBVarKey128;
    default: *count = ECBVarTxt128_count; return ECBVarTxt128;
    }
}

REF_API void ref_aes_xmr(const uint8_t* in, uint8_t* out, uint64_t n, const uint8_t* keys, int key_per_unit,
                         int dir, uint32_t nc, int count_errors, int count_syncs, const ref_fault* faults,
                         ref_stats* st) {
    uint8_t rep[3][32], key[3][16];
    for (uint64_t u = 0; u < n; ++u) {
        for (uint32_t r = 0; r < nc; ++r) {
            memcpy(rep[r], in + u * 16, 16);
            memcpy(key[r], keys + (key_per_unit ? u * 16 : 0), 16);
            if (faults && faults[u].byte >= 0 && faults[u].replica == (int)r) {
                rep[r][faults[u].byte] ^= (uint8_t)(1u << faults[u].bit);
                st->injected++;
            }
            aes_enc_dec(rep[r], key[r], (unsigned char)dir);
        }
        ref_vote(rep, nc, 1, 16, count_errors, count_syncs, u, out + u * 16, st);
    }
}

/* pthread fan-out for the CPU baseline ("reference" kind): aes_enc_dec() only reads the global tables */
typedef struct { const uint8_t* in; uint8_t* out; const uint8_t* keys; int kpu, dir; uint32_t nc; int ce, cs; const ref_fault* faults; } aes_mt;
static void aes_shard(void* p, uint64_t u0, uint64_t n, ref_stats* st) {
    aes_mt* a = (aes_mt*)p;
    ref_aes_xmr(a->in + u0 * 16, a->out + u0 * 16, n, a->keys + (a->kpu ? u0 * 16 : 0), a->kpu, a->dir, a->nc, a->ce, a->cs,
                a->faults ? a->faults + u0 : NULL, st);
}
REF_API void ref_aes_xmr_mt(const uint8_t* in, uint8_t* out, uint64_t n, const uint8_t* keys, int key_per_unit, int dir,
                            uint32_t nc, int count_errors, int count_syncs, const ref_fault* faults, int n_threads,
                            ref_stats* st) {
    aes_mt a = { in, out, keys, key_per_unit, dir, nc, count_errors, count_syncs, faults };
    ref_fanout(aes_shard, &a, n, n_threads, st);
}
