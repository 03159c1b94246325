/* oracle/_ref/libref_qsort.so : tests/quicksort/quicksort.c compiled from the reference tree.
 * Its main() sorts forever (a radiation-test loop, :198-252), so only quick_sort() (:121-136), init_array() (:93-115)
 * and checker() (:142-197) are driven from here. */
#include "ref_common.h"
#define main ref_qsort_main
#include "quicksort/quicksort.c"
#undef main

REF_API int ref_qsort_elements(void) { return array_elements; }
REF_API void ref_quick_sort(int* A, int len) { quick_sort(A, len); }
/* the benchmark's own input for a given seed (init_array with seed_value = seed): returns `array` */
REF_API const int* ref_qsort_init(int seed) { seed_value = seed; init_array(); return array; }
/* what qsort_test() does once: sort `array`, compare with the golden it computed the same way -> number of errors */
REF_API int ref_qsort_selfcheck(int seed) {
    seed_value = seed; init_array();
    quick_sort(golden_array, array_elements);
    quick_sort(array, array_elements);
    in_block = 1;                                  /* keep checker() quiet */
    return checker(golden_array, array, 0);
}
