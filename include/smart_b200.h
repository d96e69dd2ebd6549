/*
 * smart_b200 — C-ABI of the B200-native exact-string-matching scan engine.
 *
 * This header is the drop-in boundary for SMART's search hot path.  Every entry
 * point is `extern "C"`, takes plain pointers and sizes, and names the reference
 * interface it replaces (paths relative to the reference tree):
 *
 *   reference                                         replacement here
 *   ------------------------------------------------  ---------------------------------
 *   int search(uchar*,int,uchar*,int)                 sb_plugin_search(), and the per-algorithm
 *     src/algos/include/main.h:39, one per            `search` symbols in smart_b200/host/algo_stub.c
 *     src/algos/<name>.c (bf.c:25, kmp.c:44, ...)
 *   double *run_time,*pre_time (main.h:34-37)         pre_ms / run_ms out-parameters (milliseconds)
 *   OUTPUT(j) count++ (define.h:33)                   64-bit count return + optional sorted positions
 *   XSIZE/SIGMA/WSIZE/WORD (define.h:25-32)           SB_XSIZE/SB_SIGMA/SB_WSIZE/SB_WORD
 *   getText() text buffer in shm (smart.c:95-138)     sb_text (device-resident, possibly sharded)
 *   textgen (textgen.c:34-54)                         sb_text_generate / sb_textgen_host
 *   preKmp/preMp/Pre_Horspool/preBmBc/preBmGs/        sb_prep_tables: stand-alone builders of bit-identical
 *     preQsBc/preSo/preSA + inline HASHq/KR/QF tables   tables (an ABI and parity artefact, see below)
 *
 * Counting semantics are those of src/algos/bf.c:33-36: the number of (overlapping)
 * start positions j in [0, n-m] with x[0..m) == y[j..j+m).  A return of -1 means
 * "algorithm not applicable for this m" exactly where the reference algorithm
 * returns -1 (e.g. hash3.c:31, raita.c:38, qf23.c:44).
 *
 * The library never reads or writes text bytes at or beyond y[n]
 * (86 reference algorithms do).
 */
#ifndef SMART_B200_H
#define SMART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* limits, same values as src/algos/include/define.h:25-32 */
#define SB_XSIZE 4200 /* maximal pattern length */
#define SB_WSIZE 256  /* HASHq shift-table size */
#define SB_SIGMA 256  /* alphabet size */
#define SB_WORD 32    /* Shift-Or/And word size in bits */

/* error codes returned where a count is expected (all < -1; -1 is "not applicable") */
#define SB_NOT_APPLICABLE (-1LL)
#define SB_ERR_ARG (-2LL)
#define SB_ERR_CUDA (-3LL)
#define SB_ERR_NOMEM (-4LL)

/* flags for sb_search64 */
#define SB_FLAG_POSITIONS 0x1u    /* fill `positions` (sorted ascending, global offsets) */
#define SB_FLAG_NO_SKIP 0x2u      /* force the read-every-byte kernels even for skip-family algorithms (the
                                     sub-linear scan_skip kernel is their default for m >= 71) */
#define SB_FLAG_NO_PERIOD 0x8u    /* tuning/test override: verify every candidate in full (no period shortcut) */
#define SB_FLAG_DENSE 0x4u        /* tuning override: m <= 4, count every position even if matches look rare */
#define SB_FLAG_FILTER_SHIFT 0x10u  /* tuning override: force the per-position 8-byte window kernel (m >= 5) */
#define SB_FLAG_FILTER_B4 0x20u     /* tuning override: aligned-block fingerprint width 4 */
#define SB_FLAG_FILTER_B8 0x40u
#define SB_FLAG_FILTER_B16 0x80u
#define SB_FLAG_FILTER_B32 0x100u
#define SB_FLAG_FILTER_B4L8 0x200u   /* overlapping windows: 4-byte alignment, 8-byte fingerprint (m >= 11) */
#define SB_FLAG_FILTER_B4L16 0x400u  /* 4-byte alignment, 16-byte fingerprint (m >= 19) */
#define SB_FLAG_FILTER_B8L16 0x800u  /* 8-byte alignment, 16-byte fingerprint (m >= 23) */
#define SB_FLAG_FILTER_B2 0x1000u    /* windows of min(8, m-1) bytes at every even position (m >= 4) */
#define SB_FLAG_NO_PACK 0x2000u      /* do not use the packed recoding of a text over <= 16 byte values */
#define SB_FLAG_TABLE_SCAN 0x4000u   /* run the algorithm's OWN table-driven scan on the device: its preprocessing table
                                        (kmpNext / mpNext, hbc, Shift-Or/And masks, HASHq shift table - bit-identical to
                                        the reference's; also bmBc+bmGs, qsBc, the Karp-Rabin hash, the QF bit table) is
                                        uploaded, staged in shared memory and drives the scan loop of kmp.c / mp.c / hor.c /
                                        bm.c / qs.c / raita.c / tunedbm.c / so.c / sa.c / kr.c / hash3,5,8.c / qf*.c / aut.c /
                                        bom.c / ebom.c / sbndm2.c (m <= 32) / bndm.c / br.c / tvsbs.c / smith.c / skip.c /
                                        fjs.c / ssabs.c / fs.c / graspm.c / tbm.c / ac.c and, by the end of round 2, 111 of the
                                        143 registered algorithms (the list is sb_timing.filter_kind below), one thread per
                                        text segment.
                                        Same counts as every other mode; slower than the default fingerprint filters
                                        (DESIGN.md section 3.4).  Count only; other algorithms ignore the flag. */

/* ---- algorithm registry (names are the lower-cased SMART registry names, smart.c:142) ---- */
int sb_algo_count(void);
const char *sb_algo_name(int algo);      /* NULL if out of range */
int sb_algo_id(const char *name);        /* -1 if unknown */
/* 0 if algorithm `algo` accepts pattern length m, else the reference's -1 */
int sb_algo_applicable(int algo, int m);

/* ---- device / error plumbing ---- */
int sb_device_count(void);               /* number of visible CUDA devices, <0 on error */
/* CUDA device of the host-buffer entry points (sb_search_host, sb_plugin_search); default: the environment
 * variable SMART_B200_DEVICE, else 0.  One process per GPU sets its local rank.  0, or SB_ERR_ARG. */
int sb_set_device(int device);
int sb_get_device(void);
const char *sb_last_error(void);         /* thread-local message for the last failing call */
const char *sb_version(void);

/* ---- device-resident text (replaces the shm text segment of smart.c:553-568) ---- */
typedef struct sb_text sb_text;

/* Upload n host bytes, sharded over `ndev` devices (devices==NULL: devices 0..ndev-1; ndev<=0: 1 device).
 * Each shard holds its slice plus an (SB_XSIZE-1)-byte right halo, so any m <= SB_XSIZE can be
 * searched without re-sharding. Host bytes are only read in [0,n). */
sb_text *sb_text_upload(const unsigned char *host, size_t n, const int *devices, int ndev);
/* Generate n bytes ON the device(s) with the counter-based generator (see sb_textgen_host):
 * byte i depends only on (seed, sigma, i), so shards are generated in place. sigma in [1,256]. */
sb_text *sb_text_generate(size_t n, int sigma, uint64_t seed, const int *devices, int ndev);
/* textgen PARITY mode: the first n bytes (n <= 5 000 000) of SMART's own data/rand<sigma>/rand<sigma>.txt, generated
 * ON the device(s).  src/textgen.c:34-54 writes those files as byte = rand() % sigma from one never-seeded glibc
 * rand() stream (sigma = 2, 4, 8, 16, 32, 64, 128, 250 in that order, 5 000 000 bytes each); glibc's generator is the
 * linear recurrence r[i] = r[i-3] + r[i-31], so the host jumps to every 4096-output block and the blocks are
 * generated in parallel.  Bit-identical to the files (md5 in SURVEY.md Appendix B). */
sb_text *sb_text_generate_rand(int sigma, size_t n, const int *devices, int ndev);
/* host mirror: bytes [off, off+len) of rand<sigma>.txt by the same jump + recurrence (0, or -1 on bad arguments) */
int sb_textgen_rand_host(unsigned char *dst, int sigma, size_t off, size_t len);
/* One process per GPU (bench.py under torchrun): this process holds only part `part` of `nparts` of a
 * global text of n_total bytes, on `device`. Searches on such a handle count only the start positions the
 * part owns; the caller sums the per-part counts (no collective on the data path). */
sb_text *sb_text_generate_part(size_t n_total, int sigma, uint64_t seed, int nparts, int part, int device);
/* host bytes of the whole text are visible to the caller; only the part's slice + halo is uploaded */
sb_text *sb_text_upload_part(const unsigned char *host, size_t n_total, int nparts, int part, int device);
/* the partition rule itself: part g of nparts owns global start positions [*base, *base + *own) and needs
 * bytes [*base, *base + *len). Returns 0, -1 on bad arguments. */
int sb_shard_range(size_t n_total, int nparts, int part, size_t *base, size_t *own, size_t *len);
size_t sb_text_length(const sb_text *t);
/* selectivity estimate per text byte (mean of the collision entropy and the min-entropy of a sampled byte
 * histogram taken when the text was created; log2(sigma) for uniform texts); drives the choice of candidate
 * filter (a B-byte fingerprint passes ~2^-(B*bits) of positions) */
double sb_text_entropy_bits(const sb_text *t);
/* bits per character of the packed copy kept next to a text over <= 16 distinct byte values (4; 2 for <= 4
 * values; 1 for two), else 0 */
int sb_text_packed_bits(const sb_text *t);
int sb_text_shards(const sb_text *t);
/* copy text bytes [off, off+len) back to the host (tests, pattern extraction) */
int sb_text_download(const sb_text *t, size_t off, size_t len, unsigned char *host);
void sb_text_free(sb_text *t);

/* host mirror of the device generator: writes bytes [off, off+len) of stream (seed, sigma) */
void sb_textgen_host(unsigned char *dst, size_t off, size_t len, int sigma, uint64_t seed);

/* pinned host memory for the end-to-end path (cudaHostAlloc / cudaFreeHost) */
void *sb_host_alloc(size_t bytes);
void sb_host_free(void *p);

/* ---- the search hot path ---- */
/* Extended entry (64-bit n and count; SURVEY.md section 8b "Extended ABI").
 * Returns the occurrence count (>=0), SB_NOT_APPLICABLE, or an SB_ERR_* code.
 * positions/cap are used only with SB_FLAG_POSITIONS: the min(count, cap) SMALLEST occurrence positions
 * are written, ascending; the return value is still the full count.  The text is scanned once: the kernels mark
 * occurrences in a bitmap over the start positions, which is compacted in order on the device (no sort).
 * *pre_ms  = table build + upload (wall, ms)   -> replaces *pre_time (main.h:30)
 * *run_ms  = scan incl. count read-back (ms)   -> replaces *run_time (main.h:31); never left 0 */
long long sb_search64(int algo, const sb_text *t, const unsigned char *x, int m, unsigned flags,
                      unsigned long long *positions, size_t cap, double *pre_ms, double *run_ms);

/* Asynchronous form of the count-only search: the shape of the driver's loop (smart.c:312-345) is MANY searches
 * over one resident text, so the hot path must not wait for each 32-byte result before launching the next scan.
 * sb_search64_submit enqueues the search (pattern upload, one scan kernel per shard, count read-back — each on
 * its own stream so the scan stream carries kernels back to back) and returns a ticket >= 0, or an SB_ERR_* code;
 * up to 256 tickets may be outstanding.  sb_search_collect waits for ONE ticket (in any order), frees it and
 * returns what sb_search64 would have returned (count, SB_NOT_APPLICABLE, SB_ERR_*); *pre_ms = host preprocessing
 * + enqueue time, *run_ms = CUDA-event time of the ticket's scan kernels (max over devices; wall time has no
 * meaning for a pipelined search).  SB_FLAG_POSITIONS is not accepted here (position output is a blocking call).
 * The pattern bytes are copied at submit time: x may be reused as soon as submit returns. */
long long sb_search64_submit(int algo, const sb_text *t, const unsigned char *x, int m, unsigned flags);
long long sb_search_collect(long long ticket, double *pre_ms, double *run_ms);

/* Multi-pattern entry: count `npat` patterns of the SAME length m (stored back
 * to back, npat*m bytes) in ONE pass over the resident text instead of npat passes — the shape of the driver's
 * inner loop (smart.c:312: 500 patterns per length).  counts[p] receives the occurrence count of pattern p;
 * returns their sum, SB_NOT_APPLICABLE, or an SB_ERR_* code.  Falls back to one scan per pattern when no
 * aligned-block width is selective enough for the whole set (m < 7, tiny alphabets). */
long long sb_search_multi64(int algo, const sb_text *t, const unsigned char *patterns, int npat, int m, unsigned flags,
                            unsigned long long *counts, double *pre_ms, double *run_ms);

typedef struct sb_timing {
    double prep_host_ms;   /* host table build */
    double upload_ms;      /* pattern/table H2D */
    double kernel_ms;      /* max over shards of the CUDA-event time of the scan kernel(s) */
    double wall_ms;        /* host wall clock, first launch -> last count on host */
    int kernels_launched;  /* scan kernels launched by the last search (all shards) */
    int filter_kind;       /* 0 = scan_swar (m <= 4), 1/2/3 = scan_window 8/12/16 B, 4/8/16/32 = scan_block<B,B>,
                              408/416/816 = scan_block<B,L> (overlapping windows),
                              23..28 = scan_block<2,L> (L = 3..8 bytes at every even position),
                              301/302/304/308 = scan_packed<K> on the 2-bit recoding (300: pattern has a byte
                              the text does not contain, nothing scanned), 311/312 = scan_packed_half<KH>,
                              501..532 = scan_pswar<1,KF> (SWAR chain of KF characters on the 1-bit recoding; exact when
                              KF == m), 541..556 = scan_pswar<2,KF> on the 2-bit recoding,
                              108/116/132 = scan_skip<G>, 204/208/216 = scan_multi<B>, -1 = per-pattern fallback,
                              601 kmp / 602 mp / 603 hor / 604 so / 605 sa / 606 hashq / 607 bm / 608 qs / 609 raita /
                              610 tunedbm / 611 kr / 612 qf / 613 aut / 614 bom / 615 ebom / 616 sbndm2 (m <= 32) / 617 bndm / 618 br / 619 tvsbs / 620 smith / 621 skip / 622 fjs / 623 ssabs / 624 fs / 625 graspm / 626 tbm / 627 ac / 628 zt / 629 skip3,5..8 / 630 wfr, wfr2..8, wfrq2..8, twfr, twfr2..8, twfrq2..8 / 631 bndmq2,4,6 / 632 fndm / 633 sbndm / 634 ts / 635 kmpskip / 636 bom2 / 637 iom / 638 qlqs / 639 tsw / 640 tvsbs-w2 / 641 bxs, bxs2..8 / 642 sbndm-bmh / 643 lbndm / 644 bsdm, bsdm4,5,7,8 / 645 aoso2,4,6 / 646 svm0 / 647 wc / 648 fs-w1 / 649 fbom, sfbom / 650 wom / 651 pbmh = table-driven scans; sebom runs 615 (SB_FLAG_TABLE_SCAN) */
    unsigned long long verified; /* candidates sent to exact verification (all shards) */
} sb_timing;
void sb_last_timing(sb_timing *out);

/* CUDA-event stopwatch on the library's launching stream of `device` (bench.py brackets its K timed steps
 * with it: torch.cuda.Event only sees torch's own stream). begin records an event; end records a second
 * one, waits for it and returns the elapsed milliseconds (<0 on error). */
int sb_stream_timer_begin(int device);
double sb_stream_timer_end_ms(int device);

/* Host-buffer entry: text in host memory (pinned or pageable), streamed to device 0 in chunks with an
 * m-1 overlap, copy of chunk k+1 overlapped with the scan of chunk k. This is what the legacy
 * `search()` plugins call, and what bench.py times as `e2e`. */
long long sb_search_host(int algo, const unsigned char *x, int m, const unsigned char *y, size_t n,
                         unsigned flags, double *pre_ms, double *run_ms);

/* Batch form of the host-buffer entry: MANY searches over ONE host text - the shape of the driver's loop, which
 * runs every selected algorithm and every pattern of a set over the same text segment (smart.c:290-380).  The text
 * crosses PCIe once per call (per 32 jobs), not once per search: every 64 MiB chunk is copied and then scanned by
 * all jobs.  count receives what sb_search_host would return for that job (count, SB_NOT_APPLICABLE, SB_ERR_ARG);
 * pre_ms / run_ms receive the job's share of the call.  Returns the sum of the non-negative counts, or SB_ERR_*. */
typedef struct sb_host_job {
    int algo;
    const unsigned char *x;
    int m;
    unsigned flags;
    unsigned long long starts; /* 0: every start position; else only start positions j < starts are counted (the host
                                  buffer is one part of a larger text followed by its m-1 halo: the part's own starts) */
    long long count;
    double pre_ms, run_ms;
} sb_host_job;
long long sb_search_host_batch(sb_host_job *jobs, int njobs, const unsigned char *y, size_t n);

/* Legacy plugin body: exact reference signature semantics (main.h:39), int n, int count.
 * Returns -1 where the reference algorithm does; counts above INT_MAX saturate to INT_MAX. */
int sb_plugin_search(int algo, unsigned char *x, int m, unsigned char *y, int n, double *pre_ms,
                     double *run_ms);

/* ---- pattern preprocessing stage (tables bit-identical to the reference's) ----
 * sb_prep_tables is a stand-alone entry point: the search calls do NOT run it.  What a search preprocesses and
 * places on the device is its scan plan: the pattern words, the candidate fingerprints of the scan direction of
 * the algorithm (kernel parameter block = constant bank), the pattern's period from the border table (mp.c:26-35)
 * and, for the sampling kernels, a Bloom bitmap + exact q-gram table (shared memory / L2) — cached per pattern.
 */
typedef struct sb_tables {
    int m;
    int mpNext[SB_XSIZE + 1];   /* mp.c:26-35 */
    int kmpNext[SB_XSIZE + 1];  /* kmp.c:27-41 */
    int hbc[SB_SIGMA];          /* hor.c:26-30 == bm.c:27-33 (bmBc) == raita.c/tunedbm.c preBmBc;
                                   fs/graspm: the variant over all m characters (fs.c:50-51) */
    int qsBc[SB_SIGMA];         /* qs.c:26-30 */
    int suff[SB_XSIZE];         /* bm.c:36-52 */
    int bmGs[SB_XSIZE];         /* bm.c:54-66 */
    unsigned int soS[SB_SIGMA]; /* so.c:27-38 (first min(m,32) chars, as search_large does so.c:69-80) */
    unsigned int soLim;
    unsigned int saS[SB_SIGMA]; /* sa.c:27-34 */
    unsigned int saF;           /* sa.c:45 */
    int hashShift[SB_WSIZE];    /* hash3.c:37-55 / hash5.c / hash8.c for the q of the algorithm */
    int hashSh1;
    int hashQ;                  /* 0 if the algorithm is not a HASHq */
    int krD, krHx;              /* kr.c:34-40 */
    int acEll;                  /* ac.c:52-53 */
    int bucketHead[SB_SIGMA];   /* skip.c:31-39 / kmpskip.c:71-77 z[]; graspm.c:41-45 lists (see prep.c) */
    int bucketNext[SB_XSIZE + 1]; /* ... list[] / next pointers, -1 terminated */
    int fsGs[SB_XSIZE + 1];     /* fs.c:26-41 Pre_GS, entries [0..m] */
    unsigned int bndmB[SB_SIGMA]; /* bndm.c:33-38 (first min(m,32) chars, reversed) */
    int brBc[SB_SIGMA * SB_SIGMA]; /* br.c:26-37 == tvsbs.c:26-37, row-major [a][b] */
    int qfQ, qfS;               /* qf<Q><S>.c:28-29; 0 if not a QF */
    int qfB[1 << 18];           /* qf23.c:50-56; entries [0, 1<<(Q*S)) */
    int wfrQ;                   /* wfrq<Q>.c / twfrq<Q>.c: q-gram length; 0 for wfr, wfr2..8, twfr*, lwfr* */
    unsigned char wfrF[65536];  /* wfr.c:41-58, wfrq4.c:50-65: weak-factor table, 1 = some factor hashes here */
    int ztBc[SB_SIGMA * SB_SIGMA]; /* zt.c:64-74, row-major [a][b]; filled for zt only */
} sb_tables;
/* Build every table the algorithm family of `algo` uses.  The other small tables and all scalars (qfQ, qfS, wfrQ,
 * hashQ, ...) are zeroed; the four big arrays (brBc, qfB, wfrF, ztBc) are written only for the families that use
 * them and are otherwise left as the caller passed them.
 * Returns 0, or -1 when the algorithm is not applicable to m (same rule as sb_search64). */
int sb_prep_tables(int algo, const unsigned char *x, int m, sb_tables *out);
/* The string-matching automaton of aut.c:26-45 / preSMA (include/AUTOMATON.h:185-198): delta[q*256 + c] = next state
 * for q in [0, m] (entries equal the reference's int table; 16 bits suffice).  delta holds (m+1)*256 entries.
 * Used by SB_FLAG_TABLE_SCAN for `aut`.  Returns 0, or -2 on bad arguments. */
int sb_prep_automaton(const unsigned char *x, int m, unsigned short *delta);
/* The factor oracle of bom.c:49-70 as a dense table: trans[p*256 + c] = target state or -1 for p in [0, m] (what
 * getTransition, bom.c:28-39, answers on the reference's lists), term[p] = 1 for the states on the supply path of
 * state 0 (bom.c:65-69).  trans holds (m+1)*256 entries, term m+1.  Used by SB_FLAG_TABLE_SCAN for `bom`. */
int sb_prep_bom(const unsigned char *x, int m, short *trans, unsigned char *term);

#ifdef __cplusplus
}
#endif
#endif /* SMART_B200_H */
