/* oracle/_ref/libref_mm.so : tests/mm_common/mm_tmr.c (9x9 uint32 + xor_golden). */
#include "ref_common.h"
#define main ref_mm_main
#include "mm_common/mm_tmr.c"
#undef main
REF_API int ref_mm_error(void) { mm_run_test(); return checkGolden(); }            /* mm_tmr.c:33-39 */
REF_API const uint32_t* ref_mm_first(void) { return &first_matrix[0][0]; }
REF_API const uint32_t* ref_mm_second(void) { return &second_matrix[0][0]; }
REF_API const uint32_t* ref_mm_results(void) { return &results_matrix[0][0]; }
REF_API uint32_t ref_mm_xor_golden(void) { return xor_golden; }
REF_API int ref_mm_side(void) { return side; }
REF_API void ref_mm_multiply(const uint32_t* f, const uint32_t* s, uint32_t* r) {
    matrix_multiply((mm_t(*)[side])f, (mm_t(*)[side])s, (mm_t(*)[side])r);
}
