/*
 * One per-algorithm plugin: compiled once per registry name with -DSB_ALGO=<name>
 * (bin/algos/bf, bin/algos/kmp, ...).  It exports the reference's entry point
 *     int search(unsigned char *x, int m, unsigned char *y, int n)      (main.h:39, bf.c:25, kmp.c:44, ...)
 * with the reference's semantics: number of overlapping occurrences, -1 when the algorithm does not
 * apply to this m, *pre_time / *run_time filled in milliseconds.  x and y are borrowed and never written.
 */
#include "sb_shim.h"

#include "sb_client.h"
#include "smart_b200.h"

#define SB_STR2(s) #s
#define SB_STR(s) SB_STR2(s)

int search(unsigned char *x, int m, unsigned char *y, int n)
{
    double pre = 0, run = 0;
    const int algo = sb_algo_id(SB_STR(SB_ALGO));
    if (algo < 0) {
        fprintf(stderr, "smart_b200: unknown algorithm %s\n", SB_STR(SB_ALGO));
        exit(1);
    }
    if (m < 1 || n < 0) return -1;
    if (sb_algo_applicable(algo, m) != 0) return -1; /* e.g. hash3.c:31, raita.c:38, qf23.c:44 */
    const int count = sb_client_search(algo, x, m, y, n, sb_shim_pkey, sb_shim_tkey, &pre, &run);
    if (count < 0) {
        /* a CUDA failure must not look like a count: non-zero exit => the driver prints [--] */
        fprintf(stderr, "smart_b200: search failed: %s\n", sb_last_error());
        exit(2);
    }
    *pre_time = pre;
    *run_time = run > 0 ? run : 1e-6; /* the reports treat 0 as "no result" (output.h:140) */
    return count;
}
