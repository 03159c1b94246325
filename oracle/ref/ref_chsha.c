/* oracle/_ref/libref_chsha.so : tests/chstone/sha/{sha.c,sha_data.c,sha_driver.c} compiled from the reference tree. */
#include "ref_common.h"
/* sha.c defines its OWN memset (4 arguments, :55) and memcpy (word-wise, :71); keep them out of libc's names */
#define memset chs_memset
#define memcpy chs_memcpy
#define main ref_chsha_main
#include "chstone/sha/sha.c"
#include "chstone/sha/sha_data.c"
#include "chstone/sha/sha_driver.c"
#undef main
#undef memset
#undef memcpy

REF_API int ref_chsha_run_main(void) { return ref_chsha_main(); }        /* prints "Result: 5" / "RESULT: PASS" */
REF_API const uint8_t* ref_chsha_indata(void) { return &indata[0][0]; }
REF_API uint32_t ref_chsha_len(void) { return (uint32_t)VSIZE * BLOCK_SIZE; }
REF_API const uint32_t* ref_chsha_golden(void) { return outData; }

/* one stream of `len` bytes through the reference's sha_init / sha_update / sha_final */
REF_API void ref_chsha(const uint8_t* data, int len, uint32_t out[5]) {
    sha_init();
    sha_update(data, len);
    sha_final();
    for (int i = 0; i < 5; ++i) out[i] = sha_info_digest[i];
}

/* n streams of `len` bytes each -> n x 5 u32, under nc replicas; faults[u] optional (NULL = none). */
REF_API void ref_chsha_xmr(const uint8_t* in, uint32_t* out, uint64_t n, uint32_t len, uint32_t nc,
                           int count_errors, int count_syncs, const ref_fault* faults, ref_stats* st) {
    uint8_t* priv = (uint8_t*)malloc(len ? len : 1);
    uint8_t rep[3][32];
    for (uint64_t u = 0; u < n; ++u) {
        for (uint32_t r = 0; r < nc; ++r) {
            memcpy(priv, in + u * len, len);
            if (faults && faults[u].byte >= 0 && faults[u].replica == (int)r) {
                priv[faults[u].byte] ^= (uint8_t)(1u << faults[u].bit);
                st->injected++;
            }
            uint32_t dg[5];
            ref_chsha(priv, (int)len, dg);
            memcpy(rep[r], dg, 20);
        }
        ref_vote(rep, nc, 4, 5, count_errors, count_syncs, u, (uint8_t*)(out + 5 * u), st);
    }
    free(priv);
}
