/* oracle/_ref/libref_sha4000.so : the 4000-byte KAT of tests/hifive1/sha256.tmr/sha_data.inc */
#include "ref_common.h"
typedef uint32_t mm_t;
unsigned error;
#include "hifive1/sha256.tmr/sha_data.inc"
#include "sha256_common/sha256_common_tmr.c"
REF_API unsigned ref_sha4000_kat(void) { sha_run_test(); return checkGolden(); }
REF_API const uint8_t* ref_sha4000_msg(void) { return hash_data; }
REF_API uint32_t ref_sha4000_len(void) { return LEN; }
REF_API const uint8_t* ref_sha4000_golden(void) { return golden; }
