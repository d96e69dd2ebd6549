/* A plain-C client of the C-ABI (include/smart_b200.h): what a SMART-side maintainer would write.
 * Build: make -C smart_b200/host (-> bin/abi_demo).  Run on a B200: prints one line per check, exit 0 = all ok. */
#include "smart_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static long long naive(const unsigned char *x, int m, const unsigned char *y, size_t n)
{
    long long c = 0;
    for (size_t j = 0; j + (size_t)m <= n; j++) c += (memcmp(x, y + j, (size_t)m) == 0);
    return c;
}

int main(void)
{
    int bad = 0;
    if (sb_device_count() < 1) {
        fprintf(stderr, "abi_demo: no CUDA device: %s\n", sb_last_error());
        return 2;
    }
    /* 1. resident text generated on the device, host mirror of the generator for the check */
    const size_t n = 3u << 20;
    unsigned char *y = malloc(n);
    sb_textgen_host(y, 0, n, 4, 99);
    sb_text *t = sb_text_generate(n, 4, 99, NULL, 1);
    if (!t) {
        fprintf(stderr, "abi_demo: %s\n", sb_last_error());
        return 2;
    }
    const char *algos[] = {"bf", "kmp", "hor", "so", "bm", "hash3", "qf23", "smith", "ac"};
    const int ms[] = {1, 3, 8, 12, 40, 300, 4200};
    for (unsigned a = 0; a < sizeof algos / sizeof *algos; a++)
        for (unsigned k = 0; k < sizeof ms / sizeof *ms; k++) {
            const int m = ms[k], id = sb_algo_id(algos[a]);
            const unsigned char *x = y + 777777;
            double pre, run;
            long long got = sb_search64(id, t, x, m, 0, NULL, 0, &pre, &run);
            long long want = sb_algo_applicable(id, m) == 0 ? naive(x, m, y, n) : -1;
            if (got != want || (want >= 0 && run <= 0)) {
                printf("FAIL %s m=%d got %lld want %lld\n", algos[a], m, got, want);
                bad++;
            }
        }
    printf("sb_search64: %s\n", bad ? "FAILED" : "ok");
    /* 2. sorted positions */
    {
        unsigned long long pos[64];
        long long c = sb_search64(sb_algo_id("bf"), t, y + 1000, 9, SB_FLAG_POSITIONS, pos, 64, NULL, NULL);
        int ok = c >= 1;
        for (long long i = 0; i < c && i < 64; i++) {
            ok = ok && memcmp(y + pos[i], y + 1000, 9) == 0;
            if (i) ok = ok && pos[i] > pos[i - 1];
        }
        printf("positions (%lld): %s\n", c, ok ? "ok" : "FAILED");
        bad += !ok;
    }
    /* 3. the whole pattern set of one length in one pass */
    {
        enum { NP = 100, M = 24 };
        static unsigned char pats[NP * M];
        unsigned long long counts[NP];
        for (int p = 0; p < NP; p++) memcpy(pats + p * M, y + 5000 + 4099 * p, M);
        long long total = sb_search_multi64(sb_algo_id("kmp"), t, pats, NP, M, 0, counts, NULL, NULL);
        int ok = total >= NP;
        for (int p = 0; p < NP; p += 17) ok = ok && (long long)counts[p] == naive(pats + p * M, M, y, n);
        printf("multi-pattern (%lld): %s\n", total, ok ? "ok" : "FAILED");
        bad += !ok;
    }
    /* 4. the legacy plugin body on host buffers (int n, int count), and its -1 domain */
    {
        double pre = 0, run = 0;
        int c = sb_plugin_search(sb_algo_id("qs"), y + 31, 16, y, (int)n, &pre, &run);
        int ok = c == (int)naive(y + 31, 16, y, n) && run > 0;
        ok = ok && sb_plugin_search(sb_algo_id("raita"), y, 1, y, (int)n, &pre, &run) == -1;
        printf("plugin search (%d): %s\n", c, ok ? "ok" : "FAILED");
        bad += !ok;
    }
    sb_text_free(t);
    free(y);
    return bad ? 1 : 0;
}
