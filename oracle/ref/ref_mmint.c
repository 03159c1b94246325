/* oracle/_ref/libref_mmint.so : tests/matrixMultiply/matrixMultiply.c (9x9 int, self-golden). */
#include "ref_common.h"
#define main ref_mmint_main
#include "matrixMultiply/matrixMultiply.c"
#undef main
REF_API int ref_mmint_run_main(void) { return ref_mmint_main(); }   /* prints "Number of errors: 0" */
REF_API const int* ref_mmint_first(void) { return &first_matrix[0][0]; }
REF_API const int* ref_mmint_second(void) { return &second_matrix[0][0]; }
REF_API const unsigned* ref_mmint_results(void) { return &results_matrix[0][0]; }
REF_API int ref_mmint_side(void) { return side; }
