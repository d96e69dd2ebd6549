/*
 * smart — experiment driver, rewritten from scratch to the conventions of the reference's src/smart.c so
 * that it can drive both our per-algorithm executables and the reference's:
 *
 *   - text corpus: data/<buffer>/index.txt lists "#file#" entries, files are concatenated in index order
 *     and truncated to the text size (smart.c:95-138); "-text A-B-C" runs one experiment per buffer;
 *   - handoff: 5 SysV shared-memory segments — text (TSIZE+10), pattern (XSIZE+1), count (4), search time
 *     (8) and preprocessing time (8), keys below 1000 — and one process per search started as
 *         <bindir>/<algo> shared <pkey> <m> <tkey> <n> <rkey> <ekey> <prekey>        (smart.c:140-146)
 *   - patterns of length m are cut from the text at random offsets, so every search must report >= 1
 *     occurrence: 0 marks the cell [ERROR], -1 (not applicable / failed) [--], a mean above the time
 *     bound [OUT] (smart.c:330-343,376-378);
 *   - reported time per cell = mean over the pattern set of search time (+ preprocessing unless -pre),
 *     milliseconds, with best / worst / standard deviation on request.
 *
 * What is new: the text stays DEVICE-RESIDENT across the whole algorithm x length x pattern loop through
 * the resident worker (smartd, started and stopped here unless -noworker); -seed makes the pattern sets
 * reproducible (the reference seeds with time(NULL), smart.c:435); -algo selects algorithms on the command
 * line (the reference reads source/algorithms.h, function.h:62-77, which -alist still accepts).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <libgen.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ipc.h>
#include <sys/shm.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "sb_client.h"
#include "sb_proto.h"

#define XSIZE 4200 /* define.h:25 */
#define SIGMA 256
#define MAX_ALGOS 500 /* NumAlgo, smart.c:21 */
#define MAX_LENS 17   /* NumPatt */
#define MB (1024 * 1024)

/* pattern-length sets, sets.h:23-25 */
static const int LEN_VERY_SHORT[MAX_LENS] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 0};
static const int LEN_SHORT[MAX_LENS] = {2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 0};
static const int LEN_LARGE[MAX_LENS] = {2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 0, 0, 0, 0, 0};

typedef struct {
    int key, id;
    void *ptr;
    size_t size;
} segment;

typedef struct {
    const int *lens;
    int minlen, maxlen;
    int npat;       /* -pset */
    long tsize;     /* bytes */
    int occ, pre, dif, std, txt, xml, tex, html;
    int limit_ms;   /* -tb */
    unsigned seed;
    int use_worker;
    int batch;      /* -batch: one worker request per (algorithm, length) for the whole pattern set */
    int pipeline;   /* -pipeline: the pattern set as independent searches, submitted without waiting for each other */
    char bindir[512];
    char datadir[512];
    char resdir[512];
    char code[64];
    int nalgo;
    char algos[MAX_ALGOS][32];
    const char *simple_p, *simple_t;
} config;

static segment seg_text, seg_pat, seg_count, seg_time, seg_pre;
static segment seg_mpat, seg_mcounts; /* -batch: pattern set (npat x m) and per-pattern counts (npat x u64) */

/* Pattern sampling has its own generator (splitmix64) so that a given -seed always yields the same pattern
 * sets: glibc's rand() and random() share one state, and the number of rand() calls spent on picking free
 * shm keys varies from run to run. */
static unsigned long long pat_rng_state;
static unsigned long long pat_rng(void)
{
    unsigned long long z = (pat_rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void usage(void)
{
    printf("\n\tSMART driver (smart_b200)\n\n"
           "\t-text A[-B...]   corpus buffers under <datadir> (index.txt convention)\n"
           "\t-pset N          patterns per length (default 500)\n"
           "\t-plen L U        only pattern lengths in [L,U]\n"
           "\t-short | -vshort lengths 2,4,..,32 | 1,2,..,16 (default 2,4,8,..,4096)\n"
           "\t-tsize S         text size in MB (default 1)\n"
           "\t-tb MS           time bound per cell in ms (default 300)\n"
           "\t-occ -pre -dif -std -txt   as in SMART\n"
           "\t-xml             also write results/<code>/<text>.xml (schema of SMART's outputXML)\n"
           "\t-tex             also write results/<code>/<text>.tex (tabular, as SMART's outputLatex)\n"
           "\t-html            also write results/<code>/<text>.html and results/<code>/index.html\n"
           "\t-simple P T      one search of pattern P in text T\n"
           "\t-algo a,b,c      algorithms to run (default bf,kmp,hor,so)\n"
           "\t-alist FILE      read the selection from a SMART registry file (#1 #name lines)\n"
           "\t-bin DIR         directory of the algorithm executables (default: bin/algos next to this binary)\n"
           "\t-data DIR        corpus directory (default data)\n"
           "\t-seed S          seed for pattern sampling (default: time)\n"
           "\t-noworker        do not start the resident GPU worker\n"
           "\t-batch           scan the text ONCE per (algorithm, length) for the whole pattern set\n"
           "\t                 (needs the worker; reported time = time of the pass / number of patterns)\n"
           "\t-pipeline        one scan per pattern as usual, but the whole set of a length is handed to the\n"
           "\t                 worker at once and the scans run back to back on the GPU (asynchronous\n"
           "\t                 submit/collect); reported time = CUDA-event time of each pattern's own scan\n\n");
}

/* ---- shared memory: keys are small random integers like the reference's rand()%1000 ---- */
static int seg_create(segment *s, size_t size, const segment *others[], int nothers)
{
    for (int attempt = 0; attempt < 200; attempt++) {
        int key = 1 + rand() % 999, clash = 0;
        for (int i = 0; i < nothers; i++) clash |= (others[i]->key == key);
        if (clash) continue;
        int id = shmget((key_t)key, size, IPC_CREAT | IPC_EXCL | 0666);
        if (id < 0) continue;
        void *p = shmat(id, NULL, 0);
        if (p == (void *)-1) {
            shmctl(id, IPC_RMID, 0);
            continue;
        }
        s->key = key;
        s->id = id;
        s->ptr = p;
        s->size = size;
        return 0;
    }
    perror("shmget");
    return -1;
}

static void seg_destroy(segment *s)
{
    if (s->ptr) shmdt(s->ptr);
    if (s->id > 0) shmctl(s->id, IPC_RMID, 0);
    memset(s, 0, sizeof *s);
}

static void destroy_all(void)
{
    seg_destroy(&seg_text);
    seg_destroy(&seg_pat);
    seg_destroy(&seg_count);
    seg_destroy(&seg_time);
    seg_destroy(&seg_pre);
    seg_destroy(&seg_mpat);
    seg_destroy(&seg_mcounts);
}

/* results/<code>/<label>.<ext>, directories created on the way (output.h:119-128) */
static FILE *open_report(const config *cf, const char *label, const char *ext)
{
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", cf->resdir, cf->code);
    mkdir(cf->resdir, 0775);
    mkdir(path, 0775);
    snprintf(path, sizeof path, "%s/%s/%s.%s", cf->resdir, cf->code, label, ext);
    FILE *f = fopen(path, "w");
    if (f) printf("\tSaving data on %s/%s.%s\n", cf->code, label, ext);
    else printf("\tError in writing file %s/%s.%s\n", cf->code, label, ext);
    return f;
}

static void upper_name(char *dst, size_t cap, const char *name)
{
    snprintf(dst, cap, "%s", name);
    for (char *q = dst; *q; q++) *q = (char)toupper((unsigned char)*q);
}

/* ---- corpus loading: index.txt with #file# markers (smart.c:95-138) ---- */
static long load_text(unsigned char *T, const char *dir, long tsize, int *nalpha, int *maxcode)
{
    char path[2048];
    long n = 0;
    snprintf(path, sizeof path, "%s/index.txt", dir);
    FILE *idx = fopen(path, "r");
    if (!idx) {
        printf("\tError in loading text buffer. No index file exists.\n");
        return 0;
    }
    int c;
    while (n < tsize && (c = getc(idx)) != EOF) {
        if (c != '#') continue;
        char name[512];
        size_t k = 0;
        while ((c = getc(idx)) != EOF && c != '#')
            if (k + 1 < sizeof name) name[k++] = (char)c;
        name[k] = 0;
        snprintf(path, sizeof path, "%s/%s", dir, name);
        printf("\tLoading the file %s\n", path);
        FILE *in = fopen(path, "rb");
        if (!in) {
            printf("\tError in loading text file %s\n", path);
            continue;
        }
        n += (long)fread(T + n, 1, (size_t)(tsize - n), in);
        fclose(in);
    }
    fclose(idx);
    T[n] = 0; /* smart.c:125 */
    long freq[SIGMA] = {0};
    for (long i = 0; i < n; i++) freq[T[i]]++;
    *nalpha = *maxcode = 0;
    for (int ch = 0; ch < SIGMA; ch++)
        if (freq[ch]) {
            (*nalpha)++;
            *maxcode = ch;
        }
    printf("\tAlphabet of %d characters.\n\tGreater chararacter has code %d.\n", *nalpha, *maxcode);
    return n;
}

/* ---- algorithm selection ---- */
static void lower(char *s)
{
    for (; *s; s++) *s = (char)tolower((unsigned char)*s);
}

static void add_algos_csv(config *cf, const char *csv)
{
    char buf[4096];
    snprintf(buf, sizeof buf, "%s", csv);
    for (char *tok = strtok(buf, ","); tok && cf->nalgo < MAX_ALGOS; tok = strtok(NULL, ",")) {
        snprintf(cf->algos[cf->nalgo], sizeof cf->algos[0], "%s", tok);
        lower(cf->algos[cf->nalgo++]);
    }
}

/* registry file: entries "#<0|1> #<name> " (function.h:62-77); keeps the enabled ones */
static int add_algos_registry(config *cf, const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int c;
    while ((c = getc(f)) != EOF) {
        if (c != '#') continue;
        int flag = getc(f);
        if (flag != '0' && flag != '1') continue;
        while ((c = getc(f)) != EOF && c != '#') {}
        char name[32];
        size_t k = 0;
        while ((c = getc(f)) != EOF && c != ' ' && c != '\n' && c != '\r')
            if (k + 1 < sizeof name) name[k++] = (char)c;
        name[k] = 0;
        if (flag == '1' && k && cf->nalgo < MAX_ALGOS) {
            snprintf(cf->algos[cf->nalgo], sizeof cf->algos[0], "%s", name);
            lower(cf->algos[cf->nalgo++]);
        }
    }
    fclose(f);
    return 0;
}

/* ---- one search = one process (smart.c:140-146) ---- */
static int run_search(const config *cf, const char *algo, int m, long n)
{
    char cmd[1024];
    snprintf(cmd, sizeof cmd, "%s/%s shared %d %d %d %ld %d %d %d", cf->bindir, algo, seg_pat.key, m, seg_text.key, n,
             seg_count.key, seg_time.key, seg_pre.key);
    int rc = system(cmd);
    if (rc != 0) return -1;
    return *(int *)seg_count.ptr;
}

/* ---- resident worker ---- */
static int worker_started;

static int worker_shared; /* a worker somebody else started is listening: use it, never stop it */

static void worker_start(const config *cf, const char *selfdir)
{
    char cmd[1200];
    if (!cf->use_worker) return;
    snprintf(cmd, sizeof cmd, "%s/smartd > /dev/null 2>&1", selfdir);
    const int rc = system(cmd);
    if (rc == 0) {
        worker_started = 1;
        printf("\tResident GPU worker started (text stays in HBM across searches)\n");
    } else if (rc != -1 && WIFEXITED(rc) && WEXITSTATUS(rc) == 7) {
        worker_shared = 1;
        printf("\tResident GPU worker already running (shared; it is left running at the end)\n");
    } else {
        printf("\tNo resident GPU worker (each search runs in its own process)\n");
    }
}

static void worker_stop(const char *selfdir)
{
    char cmd[1200];
    if (!worker_started) return;
    snprintf(cmd, sizeof cmd, "%s/smartd -stop > /dev/null 2>&1", selfdir);
    if (system(cmd) != 0) { /* already gone */ }
    worker_started = 0;
}

/* -batch: the whole pattern set of length m in one request; returns 0 and fills sum/min of the counts and
 * the pass times, or -1 (worker unreachable / failed) */
static int run_batch(const config *cf, const char *algo, int m, long n, unsigned char (*set)[XSIZE + 1],
                     long long *sum, long long *min_occ, double *pre_ms, double *run_ms)
{
    unsigned char *dst = seg_mpat.ptr;
    unsigned long long *counts = seg_mcounts.ptr;
    for (int i = 0; i < cf->npat; i++) memcpy(dst + (size_t)i * m, set[i], (size_t)m);
    memset(counts, 0, (size_t)cf->npat * sizeof *counts);
    sb_request rq;
    sb_reply rp;
    memset(&rq, 0, sizeof rq);
    rq.magic = SB_PROTO_MAGIC;
    rq.kind = SB_REQ_SEARCH_MULTI;
    rq.algo = -1;
    snprintf(rq.algo_name, sizeof rq.algo_name, "%s", algo);
    rq.pkey = seg_mpat.key;
    rq.m = m;
    rq.tkey = seg_text.key;
    rq.n = (int)n;
    rq.npat = cf->npat;
    rq.ckey = seg_mcounts.key;
    {
        const char *gen = getenv("SMART_B200_TEXT_GEN");
        rq.text_gen = gen ? strtoull(gen, NULL, 10) : 0;
    }
    int fd = sb_worker_connect();
    if (fd < 0) return -1;
    int rc = sb_worker_call(fd, &rq, &rp);
    close(fd);
    if (rc != 0 || rp.status != 0) return -1;
    if (rp.count == -1) { /* not applicable */
        *sum = -1;
        *min_occ = -1;
        return 0;
    }
    *sum = 0;
    *min_occ = (long long)counts[0];
    for (int i = 0; i < cf->npat; i++) {
        *sum += (long long)counts[i];
        if ((long long)counts[i] < *min_occ) *min_occ = (long long)counts[i];
    }
    *pre_ms = rp.pre_ms + rp.upload_ms;
    *run_ms = rp.run_ms;
    return 0;
}

/* -pipeline: the pattern set of length m as npat independent scans in one worker request (SB_REQ_SEARCH_SET);
 * counts[k], run[k], pre[k] per pattern.  Returns 0, or -1 (worker unreachable / failed). */
static int run_pipeline(const config *cf, const char *algo, int m, long n, unsigned char (*set)[XSIZE + 1],
                        const unsigned long long **counts, const double **run, const double **pre)
{
    unsigned char *dst = seg_mpat.ptr;
    unsigned char *res = seg_mcounts.ptr;
    for (int i = 0; i < cf->npat; i++) memcpy(dst + (size_t)i * m, set[i], (size_t)m);
    memset(res, 0, (size_t)cf->npat * 24);
    sb_request rq;
    sb_reply rp;
    memset(&rq, 0, sizeof rq);
    rq.magic = SB_PROTO_MAGIC;
    rq.kind = SB_REQ_SEARCH_SET;
    rq.algo = -1;
    snprintf(rq.algo_name, sizeof rq.algo_name, "%s", algo);
    rq.pkey = seg_mpat.key;
    rq.m = m;
    rq.tkey = seg_text.key;
    rq.n = (int)n;
    rq.npat = cf->npat;
    rq.ckey = seg_mcounts.key;
    {
        const char *gen = getenv("SMART_B200_TEXT_GEN");
        rq.text_gen = gen ? strtoull(gen, NULL, 10) : 0;
    }
    int fd = sb_worker_connect();
    if (fd < 0) return -1;
    int rc = sb_worker_call(fd, &rq, &rp);
    close(fd);
    if (rc != 0 || rp.status != 0) return -1;
    *counts = (const unsigned long long *)res;
    *run = (const double *)(res + (size_t)cf->npat * 8);
    *pre = (const double *)(res + (size_t)cf->npat * 16);
    return 0;
}

static void rule(void)
{
    printf("\t");
    for (int i = 0; i < 60; i++) putchar('_');
    printf("\n");
}

/* ---- the experiment on one text (smart.c:178-400) ---- */
static void run_experiment(const config *cf, const char *label, unsigned char *T, long n)
{
    unsigned char *P = seg_pat.ptr;
    double *e_time = seg_time.ptr, *pre_time = seg_pre.ptr;
    static double mean[MAX_ALGOS][MAX_LENS];
    const int simple = cf->simple_p != NULL;
    unsigned char (*set)[XSIZE + 1] = malloc((size_t)cf->npat * sizeof *set);
    double *samples = malloc((size_t)cf->npat * sizeof *samples);
    memset(mean, 0, sizeof mean);

    for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++) {
        const int m = cf->lens[il];
        if (m < cf->minlen || m > cf->maxlen) continue;
        if (!simple && m >= n) continue;
        for (int i = 0; i < cf->npat; i++) { /* smart.c:148-158 */
            if (simple) {
                memcpy(set[i], cf->simple_p, (size_t)m);
            } else {
                long k = (long)(pat_rng() % (unsigned long long)(n - m));
                memcpy(set[i], T + k, (size_t)m);
            }
            set[i][m] = 0;
        }
        printf("\n");
        rule();
        if (simple) printf("\tExperimental results on %s\n", T);
        else printf("\tExperimental results on %s: %s\n", label, cf->code);
        printf("\tSearching for a set of %d patterns with length %d\n", cf->npat, m);
        printf("\tTesting %d algorithms\n\n", cf->nalgo);

        for (int a = 0; a < cf->nalgo; a++) {
            char head[64];
            snprintf(head, sizeof head, "\t - [%d/%d] %s ", a + 1, cf->nalgo, cf->algos[a]);
            for (char *q = head; *q; q++) *q = (char)toupper((unsigned char)*q);
            printf("%s", head);
            for (int i = (int)strlen(head); i < 35; i++) putchar('.');
            fflush(stdout);
            double sum = 0, sum_pre = 0, best = 1e30, worst = 0;
            long long total_occ = 0;
            int done = 0, verdict = 1; /* 1 ok, 0 [ERROR], -1 [--], -2 [OUT] */
            if (cf->batch && !simple) {
                long long min_occ = 0;
                double pre_b = 0, run_b = 0;
                if (run_batch(cf, cf->algos[a], m, n, set, &total_occ, &min_occ, &pre_b, &run_b) != 0 || min_occ < 0)
                    verdict = -1;
                else if (min_occ == 0)
                    verdict = 0;
                else {
                    const double per = (run_b + (cf->pre ? 0 : pre_b)) / cf->npat;
                    sum = per * cf->npat;
                    sum_pre = pre_b;
                    best = worst = per;
                    for (int k = 0; k < cf->npat; k++) samples[k] = per;
                    done = cf->npat;
                    if (per > cf->limit_ms) verdict = -2;
                }
            } else if (cf->pipeline && !simple) {
                const unsigned long long *counts;
                const double *run_k, *pre_k;
                if (run_pipeline(cf, cf->algos[a], m, n, set, &counts, &run_k, &pre_k) != 0) verdict = -1;
                else
                    for (int k = 0; k < cf->npat; k++) { /* same verdict rules as the per-process loop below */
                        const long long occ = (long long)counts[k];
                        double t = run_k[k];
                        if (!cf->pre) t += pre_k[k];
                        samples[k] = t;
                        sum += t;
                        sum_pre += pre_k[k];
                        if (t < best) best = t;
                        if (t > worst) worst = t;
                        total_occ += occ;
                        done++;
                        if (occ <= 0) {
                            verdict = occ == 0 ? 0 : -1;
                            break;
                        }
                        if (t > cf->limit_ms) {
                            verdict = -2;
                            break;
                        }
                    }
            } else
            for (int k = 0; k < cf->npat; k++) {
                memcpy(P, set[k], (size_t)m);
                P[m] = 0; /* smart.c:313 */
                *e_time = *pre_time = 0.0;
                *(int *)seg_count.ptr = 0;
                const int occ = run_search(cf, cf->algos[a], m, n);
                double t = *e_time;
                if (!cf->pre) t += *pre_time; /* smart.c:323 */
                samples[k] = t;
                sum += t;
                sum_pre += *pre_time;
                if (t < best) best = t;
                if (t > worst) worst = t;
                total_occ += occ;
                done++;
                if (occ <= 0 && !simple) {
                    verdict = occ == 0 ? 0 : -1;
                    break;
                }
                if (occ < 0) {
                    verdict = -1;
                    break;
                }
                if (t > cf->limit_ms) {
                    verdict = -2;
                    break;
                }
            }
            if (verdict == 1) {
                const double mu = sum / cf->npat, mu_pre = sum_pre / cf->npat;
                double var = 0;
                for (int k = 0; k < done; k++) var += (samples[k] - mu) * (samples[k] - mu);
                mean[a][il] = mu;
                printf(".[OK]  ");
                if (cf->pre) printf("\t%.2f + %.2f ms", mu_pre, mu);
                else printf("\t%.2f ms", mu);
                if (cf->dif) printf("  [%.2f, %.2f]", best, worst);
                if (cf->std) printf("  std %.2f", sqrt(var / cf->npat));
                if (cf->occ) printf("\tocc %lld", total_occ / cf->npat);
                printf("\n");
            } else if (verdict == 0) printf(".[ERROR] \n");
            else if (verdict == -1) printf(".[--]  \n");
            else printf(".[OUT]  \n");
        }
    }
    printf("\n");
    rule();
    if (cf->txt && !simple) { /* output.h:116-152: one row per algorithm, one column per length, "-" = no result */
        char path[1024];
        snprintf(path, sizeof path, "%s/%s", cf->resdir, cf->code);
        mkdir(cf->resdir, 0775);
        mkdir(path, 0775);
        snprintf(path, sizeof path, "%s/%s/%s.txt", cf->resdir, cf->code, label);
        FILE *f = fopen(path, "w");
        if (f) {
            printf("\tSaving data on %s/%s.txt\n", cf->code, label);
            for (int a = 0; a < cf->nalgo; a++) {
                char up[32];
                snprintf(up, sizeof up, "%s", cf->algos[a]);
                for (char *q = up; *q; q++) *q = (char)toupper((unsigned char)*q);
                fprintf(f, "%-20s", up);
                int first = 1;
                for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++) {
                    if (cf->lens[il] < cf->minlen || cf->lens[il] > cf->maxlen) continue;
                    if (!first) fputc('\t', f);
                    first = 0;
                    if (mean[a][il] == 0) fputc('-', f);
                    else fprintf(f, "%.2f", mean[a][il]);
                }
                fputc('\n', f);
            }
            fclose(f);
        }
    }
    if (cf->xml && !simple) { /* output.h:196-245: RESULTS/CODE/TEXT, one ALGO block per algorithm, BEST per length */
        char path[1024];
        snprintf(path, sizeof path, "%s/%s", cf->resdir, cf->code);
        mkdir(cf->resdir, 0775);
        mkdir(path, 0775);
        snprintf(path, sizeof path, "%s/%s/%s.xml", cf->resdir, cf->code, label);
        FILE *f = fopen(path, "w");
        if (f) {
            printf("\tSaving data on %s/%s.xml\n", cf->code, label);
            fprintf(f, "<RESULTS>\n\t<CODE>%s</CODE>\n\t<TEXT>%s</TEXT>\n", cf->code, label);
            for (int a = 0; a < cf->nalgo; a++) {
                char up[32];
                snprintf(up, sizeof up, "%s", cf->algos[a]);
                for (char *q = up; *q; q++) *q = (char)toupper((unsigned char)*q);
                fprintf(f, "\t<ALGO>\n\t\t<NAME>%s</NAME>\n", up);
                for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++) {
                    if (cf->lens[il] < cf->minlen || cf->lens[il] > cf->maxlen) continue;
                    /* output.h:225-228 prints the "-" marker AND THEN the ordinary block (there is no else), so a
                     * length without a result takes two DATA nodes; kept, the files are diffed against the
                     * reference driver's.  Four decimals instead of two: a GPU scan of 1 MiB takes ~0.02 ms. */
                    if (mean[a][il] == 0) fprintf(f, "\t\t<DATA>-</DATA>\n");
                    fprintf(f, "\t\t<DATA>\n\t\t\t<SEARCH>%.4f</SEARCH>\n\t\t</DATA>\n", mean[a][il]);
                }
                fprintf(f, "\t</ALGO>\n");
            }
            fprintf(f, "\t<BEST>\n");
            for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++) {
                if (cf->lens[il] < cf->minlen || cf->lens[il] > cf->maxlen) continue;
                double best = 0;
                for (int a = 0; a < cf->nalgo; a++)
                    if (mean[a][il] > 0 && (best == 0 || mean[a][il] < best)) best = mean[a][il];
                fprintf(f, "\t\t<DATA>%.4f</DATA>\n", best);
            }
            fprintf(f, "\t</BEST>\n</RESULTS>");
            fclose(f);
        }
    }
    if (cf->tex && !simple) { /* output.h:153-194: a tabular with one row per algorithm, "-" = no result */
        FILE *f = open_report(cf, label, "tex");
        if (f) {
            fprintf(f, "\\begin{tabular}{|l|");
            for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++)
                if (cf->lens[il] >= cf->minlen && cf->lens[il] <= cf->maxlen) fputc('l', f);
            fprintf(f, "|}\n\\hline\n$m$");
            for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++)
                if (cf->lens[il] >= cf->minlen && cf->lens[il] <= cf->maxlen) fprintf(f, " & $%d$", cf->lens[il]);
            fprintf(f, "\\\\\n");
            for (int a = 0; a < cf->nalgo; a++) {
                char up[32];
                upper_name(up, sizeof up, cf->algos[a]);
                fprintf(f, "\\textsc{%s}", up);
                for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++) {
                    if (cf->lens[il] < cf->minlen || cf->lens[il] > cf->maxlen) continue;
                    if (mean[a][il] == 0) fprintf(f, " & -");
                    else fprintf(f, " & %.4f", mean[a][il]);
                }
                fprintf(f, "\\\\\n");
            }
            fprintf(f, "\\hline\n\\end{tabular}");
            fclose(f);
        }
    }
    if (cf->html && !simple) {
        /* output.h:635-704 renders the XML with a script; here the same table (lengths across, one row per
         * algorithm, best time of a column underlined) is written out statically, next to an index.html in
         * the layout of output.h:706-745 so that results/<code>/ can be browsed like the reference's */
        FILE *f = open_report(cf, label, "html");
        if (f) {
            fprintf(f, "<html>\n<head>\n\t<title>SMART: experimental results %s on %s</title>\n"
                       "\t<meta http-equiv=\"Content-Type\" content=\"text/html; charset=utf-8\" />\n"
                       "\t<link href=\"../style.css\" rel=\"stylesheet\" type=\"text/css\" />\n</head>\n<body>\n"
                       "\t<h2>SMART: experimental results %s on %s</h2><br>Mean search times in milliseconds "
                       "(smart_b200, GPU)<br/>\n\t<table align=\"center\">\n\t\t<tr><td></td>",
                    cf->code, label, cf->code, label);
            for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++)
                if (cf->lens[il] >= cf->minlen && cf->lens[il] <= cf->maxlen) fprintf(f, "<td><b>%d</b></td>", cf->lens[il]);
            fprintf(f, "</tr>\n");
            for (int a = 0; a < cf->nalgo; a++) {
                char up[32];
                upper_name(up, sizeof up, cf->algos[a]);
                fprintf(f, "\t\t<tr><td><b>%s</b></td>", up);
                for (int il = 0; il < MAX_LENS && cf->lens[il] > 0; il++) {
                    if (cf->lens[il] < cf->minlen || cf->lens[il] > cf->maxlen) continue;
                    double best = 0;
                    for (int b = 0; b < cf->nalgo; b++)
                        if (mean[b][il] > 0 && (best == 0 || mean[b][il] < best)) best = mean[b][il];
                    if (mean[a][il] == 0) fprintf(f, "<td>--</td>");
                    else if (mean[a][il] == best) fprintf(f, "<td><b><u>%.4f</u></b></td>", mean[a][il]);
                    else fprintf(f, "<td>%.4f</td>", mean[a][il]);
                }
                fprintf(f, "</tr>\n");
            }
            fprintf(f, "\t</table>\n</body>\n</html>");
            fclose(f);
        }
        char path[1024];
        snprintf(path, sizeof path, "%s/%s/index.html", cf->resdir, cf->code);
        const int fresh = access(path, F_OK) != 0;
        f = fopen(path, fresh ? "w" : "r+");
        if (f) { /* one row per text buffer; appended before the closing tags on later buffers */
            static const char tail[] = "\t</table>\n</body>\n</html>";
            if (fresh)
                fprintf(f, "<html>\n<head>\n\t<title>SMART: experimental results %s</title>\n"
                           "\t<meta http-equiv=\"Content-Type\" content=\"text/html; charset=utf-8\" />\n"
                           "\t<link href=\"../style.css\" rel=\"stylesheet\" type=\"text/css\" />\n</head>\n<body>\n"
                           "\t<h2>SMART: experimental results %s</h2><br>\n\t<table align=\"center\">\n",
                        cf->code, cf->code);
            else
                fseek(f, -(long)(sizeof tail - 1), SEEK_END);
            fprintf(f, "\t\t<tr><td><a href=\"%s.html\">Experimental results on %s</a></td></tr>\n%s", label, label, tail);
            fclose(f);
            printf("\tWriting %s/index.html\n", cf->code);
        }
    }
    free(set);
    free(samples);
}

int main(int argc, char **argv)
{
    config cf;
    memset(&cf, 0, sizeof cf);
    cf.lens = LEN_LARGE;
    cf.minlen = 1;
    cf.maxlen = XSIZE;
    cf.npat = 500;
    cf.tsize = MB;
    cf.limit_ms = 300;
    cf.seed = (unsigned)time(NULL);
    cf.use_worker = 1;
    char selfbuf[512], texts[512] = "";
    snprintf(selfbuf, sizeof selfbuf, "%s", argv[0]);
    const char *selfdir = dirname(selfbuf);
    snprintf(cf.bindir, sizeof cf.bindir, "%s/algos", selfdir);
    snprintf(cf.datadir, sizeof cf.datadir, "data");
    snprintf(cf.resdir, sizeof cf.resdir, "results");

    if (argc == 1) {
        printf("No parameter given. Use -h for help.\n\n");
        return 0;
    }
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
#define NEED(k)                                                            \
    if (i + (k) >= argc) {                                                 \
        printf("Error in input parameters. Use -h for help.\n\n");         \
        return 0;                                                          \
    }
        if (!strcmp(a, "-h")) { usage(); return 0; }
        else if (!strcmp(a, "-pset")) { NEED(1); cf.npat = atoi(argv[++i]); }
        else if (!strcmp(a, "-tsize")) { NEED(1); cf.tsize = atol(argv[++i]) * MB; }
        else if (!strcmp(a, "-tb")) { NEED(1); cf.limit_ms = atoi(argv[++i]); }
        else if (!strcmp(a, "-text")) { NEED(1); snprintf(texts, sizeof texts, "%s", argv[++i]); }
        else if (!strcmp(a, "-plen")) { NEED(2); cf.minlen = atoi(argv[++i]); cf.maxlen = atoi(argv[++i]); }
        else if (!strcmp(a, "-simple")) { NEED(2); cf.simple_p = argv[++i]; cf.simple_t = argv[++i]; }
        else if (!strcmp(a, "-occ")) cf.occ = 1;
        else if (!strcmp(a, "-pre")) cf.pre = 1;
        else if (!strcmp(a, "-dif")) cf.dif = 1;
        else if (!strcmp(a, "-std")) cf.std = 1;
        else if (!strcmp(a, "-txt")) cf.txt = 1;
        else if (!strcmp(a, "-xml")) cf.xml = 1;
        else if (!strcmp(a, "-tex")) cf.tex = 1;
        else if (!strcmp(a, "-html")) cf.html = 1;
        else if (!strcmp(a, "-short")) cf.lens = LEN_SHORT;
        else if (!strcmp(a, "-vshort")) cf.lens = LEN_VERY_SHORT;
        else if (!strcmp(a, "-algo")) { NEED(1); add_algos_csv(&cf, argv[++i]); }
        else if (!strcmp(a, "-alist")) { NEED(1); if (add_algos_registry(&cf, argv[++i]) != 0) { printf("cannot read %s\n", argv[i]); return 1; } }
        else if (!strcmp(a, "-bin")) { NEED(1); snprintf(cf.bindir, sizeof cf.bindir, "%s", argv[++i]); }
        else if (!strcmp(a, "-data")) { NEED(1); snprintf(cf.datadir, sizeof cf.datadir, "%s", argv[++i]); }
        else if (!strcmp(a, "-results")) { NEED(1); snprintf(cf.resdir, sizeof cf.resdir, "%s", argv[++i]); }
        else if (!strcmp(a, "-seed")) { NEED(1); cf.seed = (unsigned)strtoul(argv[++i], NULL, 10); }
        else if (!strcmp(a, "-noworker")) cf.use_worker = 0;
        else if (!strcmp(a, "-batch")) cf.batch = 1;
        else if (!strcmp(a, "-pipeline")) cf.pipeline = 1;
        else { printf("Error in input parameters. Use -h for help.\n\n"); return 0; }
    }
    if (cf.minlen < 1 || cf.minlen > XSIZE || cf.maxlen < cf.minlen) {
        printf("Error in input parameters. The pattern length bounds are not valid.\n\n");
        return 0;
    }
    if (texts[0] && cf.simple_p) { printf("Error in input parameters. Both parameters -simple and -text defined.\n\n"); return 0; }
    if (!texts[0] && !cf.simple_p) { printf("Error in input parameters. No filename given.\n\n"); return 0; }
    if (cf.npat < 1 || cf.tsize < 1) { printf("Error in input parameters. Use -h for help.\n\n"); return 0; }
    if (cf.nalgo == 0 && add_algos_registry(&cf, "source/algorithms.h") != 0) add_algos_csv(&cf, "bf,kmp,hor,so");
    if (cf.nalgo == 0) { printf("No algorithm selected.\n"); return 0; }

    srand(cf.seed ^ (unsigned)getpid()); /* shm keys only */
    pat_rng_state = cf.seed;
    snprintf(cf.code, sizeof cf.code, "EXP%ld", (long)time(NULL)); /* smart.c:90-93 */

    const segment *none[1] = {NULL};
    if (seg_create(&seg_text, (size_t)cf.tsize + 10 + 2 * XSIZE, none, 0)) return 1;
    const segment *o1[] = {&seg_text};
    if (seg_create(&seg_time, 8, o1, 1)) { destroy_all(); return 1; }
    const segment *o2[] = {&seg_text, &seg_time};
    if (seg_create(&seg_pre, 8, o2, 2)) { destroy_all(); return 1; }
    const segment *o3[] = {&seg_text, &seg_time, &seg_pre};
    if (seg_create(&seg_pat, XSIZE + 1, o3, 3)) { destroy_all(); return 1; }
    const segment *o4[] = {&seg_text, &seg_time, &seg_pre, &seg_pat};
    if (seg_create(&seg_count, 4, o4, 4)) { destroy_all(); return 1; }
    unsigned char *T = seg_text.ptr;

    if (cf.batch || cf.pipeline) {
        const segment *o5[] = {&seg_text, &seg_time, &seg_pre, &seg_pat, &seg_count};
        if (seg_create(&seg_mpat, (size_t)cf.npat * XSIZE, o5, 5)) { destroy_all(); return 1; }
        const segment *o6[] = {&seg_text, &seg_time, &seg_pre, &seg_pat, &seg_count, &seg_mpat};
        if (seg_create(&seg_mcounts, (size_t)cf.npat * 24, o6, 6)) { destroy_all(); return 1; }
    }
    worker_start(&cf, selfdir);
    if ((cf.batch || cf.pipeline) && !worker_started && !worker_shared) {
        printf("\t-batch / -pipeline need the resident GPU worker\n");
        destroy_all();
        return 1;
    }
    if (cf.simple_p) {
        static int one[MAX_LENS];
        const long n = (long)strlen(cf.simple_t);
        if (n > cf.tsize) { printf("text too long\n"); destroy_all(); return 1; }
        memcpy(T, cf.simple_t, (size_t)n + 1);
        one[0] = (int)strlen(cf.simple_p);
        cf.lens = one;
        cf.minlen = 1;
        cf.maxlen = XSIZE;
        cf.npat = 1;
        run_experiment(&cf, "simple", T, n);
    } else {
        char *save = NULL;
        for (char *buf = strtok_r(texts, "-", &save); buf; buf = strtok_r(NULL, "-", &save)) {
            char dir[1100];
            int nalpha, maxcode;
            snprintf(dir, sizeof dir, "%s/%s", cf.datadir, buf);
            printf("\n\tTry to process archive %s\n", buf);
            const long n = load_text(T, dir, cf.tsize, &nalpha, &maxcode);
            if (n <= 0) {
                printf("\tText buffer %s is empty or missing\n", buf);
                continue;
            }
            printf("\tText buffer of dimension %ld byte\n", n);
            {   /* tell the worker (through the algorithm executables' environment) that the text changed */
                static unsigned long long text_gen;
                char gen[40];
                snprintf(gen, sizeof gen, "%llu", ((unsigned long long)getpid() << 24) | ++text_gen);
                setenv("SMART_B200_TEXT_GEN", gen, 1);
            }
            run_experiment(&cf, buf, T, n);
        }
    }
    worker_stop(selfdir);
    destroy_all(); /* all five: the reference leaks its two 8-byte segments */
    return 0;
}
