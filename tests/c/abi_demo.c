/* abi_demo.c -- a plain-C caller of the C ABI (include/coast_rt.h): no CUDA headers, no Python.
 * What a reference-side integration looks like: batch launch with device buffers,
 * an injected fault plan, counters folded into the reference's globals, a user DWC handler, and the board-style
 * report line simulation/platform/resources/decoder.py:66 parses.
 *   gcc -I include tests/c/abi_demo.c -L coast_b200 -lcoast_rt -Wl,-rpath,$PWD/coast_b200 -o abi_demo && ./abi_demo */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "COAST.h"
#include "coast_rt.h"

__DEFAULT_NO_xMR

static int dwc_seen = 0;
void FAULT_DETECTED_DWC(void) { dwc_seen++; }          /* overrides the library's weak abort() default */

#define CHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, coast_last_error()); return 2; } } while (0)

int main(void) {
    CHECK(coast_init(0));
    const uint64_t n = 100000;
    uint32_t nc, flags;
    CHECK(coast_parse_opt_passes("-TMR -countErrors -countSyncs", &nc, &flags));

    /* CRC16 of the shipped message through the one-call host path */
    CHECK(coast_set_opt_passes("-TMR -countErrors"));
    unsigned short crc = coast_xmr_crc16((const unsigned char*)"Automated TMR", 13);
    printf("result: %hx\n", crc);                      /* tests/crc16/crc16.c:42 */

    /* SHA-256 TMR over device buffers with a Bernoulli(1/16) single-bit-flip plan */
    void *d_in, *d_out;
    CHECK(coast_malloc(&d_in, n * 64));
    CHECK(coast_malloc(&d_out, n * 32));
    CHECK(coast_fill_philox(d_in, n * 16, 0, 2, NULL));
    coast_fault_plan plan = { COAST_PLAN_BERNOULLI, 22, 0, 1u << 28, NULL };
    coast_launch_desc d;
    memset(&d, 0, sizeof d);
    d.kernel = COAST_K_SHA256; d.num_clones = nc; d.flags = flags; d.n_units = n; d.unit_bytes = 64;
    d.d_in = d_in; d.d_out = d_out; d.plan = &plan;
    const uint32_t before = TMR_ERROR_CNT;
    CHECK(coast_launch(&d, NULL));
    coast_stats st;
    CHECK(coast_sync(NULL, &st));
    unsigned char* faulty = malloc(n * 32), *clean = malloc(n * 32);
    CHECK(coast_memcpy_d2h(faulty, d_out, n * 32, NULL));
    CHECK(coast_stream_sync(NULL));
    d.plan = NULL;
    CHECK(coast_launch(&d, NULL));
    CHECK(coast_sync(NULL, NULL));
    CHECK(coast_memcpy_d2h(clean, d_out, n * 32, NULL));
    CHECK(coast_stream_sync(NULL));
    unsigned errors = memcmp(faulty, clean, n * 32) != 0;
    printf("injected=%llu corrected_votes=%llu syncs=%llu TMR_ERROR_CNT+=%u\n", (unsigned long long)st.injected,
           (unsigned long long)st.errors_corrected, (unsigned long long)st.syncs, TMR_ERROR_CNT - before);
    printf("C:0 E:%u F:%u T:%uus\n", errors, (unsigned)(TMR_ERROR_CNT - before), 0u);
    if (errors || st.injected == 0 || TMR_ERROR_CNT - before != (uint32_t)st.errors_corrected || st.syncs != 32 * n) return 3;

    /* the same plan under DWC: the handler runs once, after the kernel, because mismatches were detected */
    d.num_clones = COAST_DWC; d.flags = 0; d.plan = &plan;
    CHECK(coast_launch(&d, NULL));
    CHECK(coast_sync(NULL, &st));
    printf("dwc_detected=%llu handler_calls=%d\n", (unsigned long long)st.dwc_detected, dwc_seen);
    if (dwc_seen != 1 || st.dwc_detected == 0) return 4;
    CHECK(coast_free(d_in)); CHECK(coast_free(d_out));
    CHECK(coast_shutdown());
    puts("abi_demo ok");
    return 0;
}
