/* Internal declarations shared by prep.c (host C) and the CUDA translation units. */
#ifndef SB_INTERNAL_H
#define SB_INTERNAL_H

#include "smart_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum sb_family {
    SB_FAM_BF = 0,
    SB_FAM_MP,
    SB_FAM_KMP,
    SB_FAM_BM,
    SB_FAM_HOR,
    SB_FAM_QS,
    SB_FAM_RAITA,
    SB_FAM_TUNEDBM,
    SB_FAM_SO,
    SB_FAM_SA,
    SB_FAM_KR,
    SB_FAM_HASHQ,
    SB_FAM_QF,
    SB_FAM_SMITH,
    SB_FAM_TBM,
    SB_FAM_AG,
    SB_FAM_AC,
    SB_FAM_BNDM,
    SB_FAM_BR,
    SB_FAM_FJS,
    SB_FAM_SKIP,
    SB_FAM_KMPSKIP,
    SB_FAM_FS,
    SB_FAM_GRASPM,
    /* SURVEY section 8(f)-1, second batch: registered with their -1 domain and scan direction; the result is
     * the common exact count (bf.c:33-36), which is what each of them returns wherever it is correct */
    SB_FAM_BITPAR,  /* forward bit-parallel: aoso*, faoso*, svm0, blim, ufndmq* */
    SB_FAM_AUT,     /* deterministic automaton: aut */
    SB_FAM_TWOWAY,  /* constant-space linear: smoa, tw, rcolussi */
    SB_FAM_BNDMX,   /* backward bit-parallel windows: bndmq*, lbndm, sbndm*, fsbndm-w2, fndm, tndm*, bxs*, bww, dbww* */
    SB_FAM_FACTOR,  /* factor oracles / suffix automata: bom*, ebom, fbom, sebom, sfbom, ildm*, fdm, bsdm*, ww */
    SB_FAM_WFR,     /* hashed weak-factor recognition: wfr*, twfr*, lwfr* */
    SB_FAM_SKIPQ,   /* q-gram bucket skip search: skip3..8 */
    SB_FAM_BADCHAR, /* further bad-character / good-suffix shift variants: bfs, ffs, iom, wom, pbmh, qlqs, tsw, wc, zt, ... */
    SB_FAM_SSEF     /* SSE filtering: ssef (m >= 32) */
};

/* which end of the pattern the candidate fingerprint is taken from: the left-to-right families
 * (BF, MP/KMP, Shift-Or/And, KR) anchor on the prefix, the right-to-left / skip families
 * (BM, Horspool, QS, Raita, Tuned-BM, HASHq, QF) on the suffix, like their CPU scan order. */
enum sb_anchor { SB_ANCHOR_PREFIX = 0, SB_ANCHOR_SUFFIX = 1 };

typedef struct sb_algo_desc {
    const char *name;
    int family;
    int anchor;
    int q;     /* HASHq rank */
    int Q, S;  /* QF parameters; WFR: Q = q-gram length of the wfrq/twfrq variants */
    int min_m; /* smaller m -> -1, as the reference */
    int min_ratio; /* n / min(m, 16) below this -> -1 (dbww.c:33-35); 0 = no such rule */
} sb_algo_desc;

const sb_algo_desc *sb_algo_desc_get(int algo);

void sb_prep_mp(const unsigned char *x, int m, int *mpNext);
void sb_prep_kmp(const unsigned char *x, int m, const int *mpNext, int *kmpNext);
void sb_prep_hbc(const unsigned char *x, int m, int *hbc);
void sb_prep_qsbc(const unsigned char *x, int m, int *qsBc);
void sb_prep_suffixes(const unsigned char *x, int m, int *suff);
void sb_prep_bmgs(const unsigned char *x, int m, int *suff, int *bmGs);
unsigned int sb_prep_so(const unsigned char *x, int m, unsigned int *S);
unsigned int sb_prep_sa(const unsigned char *x, int m, unsigned int *S);
void sb_prep_buckets(const unsigned char *x, int m, int *head, int *next);
void sb_prep_lastocc(const unsigned char *x, int m, int *bc);
void sb_prep_fsgs(const unsigned char *x, int m, int *gs);
void sb_prep_graspm(const unsigned char *x, int m, int *head, int *next);
void sb_prep_bndm(const unsigned char *x, int m, unsigned int *B);
void sb_prep_brbc(const unsigned char *x, int m, int *brBc);
void sb_prep_hashq(const unsigned char *x, int m, int q, int *shift, int *sh1);
void sb_prep_kr(const unsigned char *x, int m, int *d, int *hx);
void sb_prep_qf(const unsigned char *x, int m, int Q, int S, int *B);
void sb_prep_wfr(const unsigned char *x, int m, int Q, unsigned char *F);
void sb_prep_ztbc(const unsigned char *x, int m, int *ztBc);
void sb_prep_skipq(const unsigned char *x, int m, int Q, int *head, int *next);
void sb_prep_bsdm(const unsigned char *x, int m, int Q, int *pos, int *st, int *len);
int sb_prep_sbndm2(const unsigned char *x, int m, unsigned int *B, int *shift);

#ifdef __cplusplus
}
#endif
#endif
