/*
 * Pattern-preprocessing stage (host side, C): the algorithm registry (names, -1 domains, scan
 * direction) and the per-algorithm tables of the SMART algorithm families.  The search entry points use
 * the registry and sb_prep_mp (the pattern's period); the full table set is built only when a caller asks
 * through sb_prep_tables (parity tests, the table-driven device filters) — never on the timed search path.
 *
 * Each builder must produce a table bit-identical to the one the reference builds
 * (tests compare against oracle/ and against the reference's own exported pre*
 * functions compiled into oracle/_ref).  They are written from the tables' definitions,
 * not transcribed from the reference loops; the reference function each one mirrors is
 * cited next to it (paths relative to the reference tree).
 */
#include "sb_internal.h"

#include <string.h>

/* ------------------------------------------------------------------ registry */

static const sb_algo_desc g_algos[] = {
    /* name      family        anchor      q  Q  S  min_m  min n/m */
    {"bf", SB_FAM_BF, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* src/algos/bf.c:25 */
    {"mp", SB_FAM_MP, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* mp.c:37 */
    {"kmp", SB_FAM_KMP, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},         /* kmp.c:44 */
    {"bm", SB_FAM_BM, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},           /* bm.c:69 */
    {"hor", SB_FAM_HOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},         /* hor.c:33 */
    {"qs", SB_FAM_QS, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},           /* qs.c:32 */
    {"raita", SB_FAM_RAITA, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0},     /* raita.c:35, m<2 -> -1 (:38) */
    {"tunedbm", SB_FAM_TUNEDBM, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* tunedbm.c:34 */
    {"so", SB_FAM_SO, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* so.c:40, search_large :69 */
    {"sa", SB_FAM_SA, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* sa.c:36, search_large :65 */
    {"kr", SB_FAM_KR, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* kr.c:28 */
    {"hash3", SB_FAM_HASHQ, SB_ANCHOR_SUFFIX, 3, 0, 0, 3, 0},     /* hash3.c:28, m<3 -> -1 (:31) */
    {"hash5", SB_FAM_HASHQ, SB_ANCHOR_SUFFIX, 5, 0, 0, 5, 0},     /* hash5.c:28 */
    {"hash8", SB_FAM_HASHQ, SB_ANCHOR_SUFFIX, 8, 0, 0, 8, 0},     /* hash8.c:28 */
    {"qf23", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 2, 3, 3, 0},         /* qf23.c:39, m<=Q -> -1 (:44) */
    {"qf24", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 2, 4, 3, 0},
    {"qf26", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 2, 6, 3, 0},
    {"qf28", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 2, 8, 3, 0},
    {"qf33", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 3, 3, 4, 0},
    {"qf34", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 3, 4, 4, 0},
    {"qf36", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 3, 6, 4, 0},
    {"qf42", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 4, 2, 5, 0},
    {"qf43", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 4, 3, 5, 0},
    {"qf44", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 4, 4, 5, 0},
    {"qf62", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 6, 2, 7, 0},
    {"qf63", SB_FAM_QF, SB_ANCHOR_SUFFIX, 0, 6, 3, 7, 0},
    /* further clean families whose tables are the ones above */
    {"smith", SB_FAM_SMITH, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},     /* smith.c:41: max(bmBc, qsBc) shift */
    {"tbm", SB_FAM_TBM, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},         /* tbm.c:66: Turbo-BM (bmGs, bmBc) */
    {"ag", SB_FAM_AG, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},           /* ag.c:59: Apostolico-Giancarlo (bmGs, suff, bmBc) */
    {"ac", SB_FAM_AC, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* ac.c:44: Apostolico-Crochemore (kmpNext, ell) */
    {"bndm", SB_FAM_BNDM, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},       /* bndm.c:26, search_large :62 (32-char prefix) */
    {"br", SB_FAM_BR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},           /* br.c:40: Berry-Ravindran brBc[256][256] */
    {"tvsbs", SB_FAM_BR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},        /* tvsbs.c:40: same table, first/last char first */
    {"fjs", SB_FAM_FJS, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},         /* fjs.c:47: Franek-Jennings-Smyth (qsBc + kmpNext) */
    {"ssabs", SB_FAM_QS, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},        /* ssabs.c:35: qsBc, first/last char first */
    {"ts", SB_FAM_BF, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0},           /* ts.c:25: Tailed-Substring, no tables */
    {"skip", SB_FAM_SKIP, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},       /* skip.c:26: Skip-Search, bucket lists per character */
    {"kmpskip", SB_FAM_KMPSKIP, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* kmpskip.c:59: buckets + mpNext + kmpNext */
    {"fs", SB_FAM_FS, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},           /* fs.c:44: Fast-Search (bc over all m chars, Pre_GS) */
    {"graspm", SB_FAM_GRASPM, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0},   /* graspm.c:37: GRASPm (bc + lists keyed by x[i]) */
    /* second batch of SURVEY section 8(f)-1: the other algorithms of the reference that agree with bf.c on every
     * golden input (tests/golden/counts.json holds their recorded outputs).  name, family, scan direction,
     * -1 domain; columns as above plus min n/m.  Not registered, because the reference's own output is wrong
     * on some golden input: the 63 of SURVEY Appendix A.2 and, found here, kbndm, ksa (absent pattern counted
     * once on rand2 for m >= 62) and ssm (m = 1). */
    {"aoso2", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 3, 0}, /* aoso2.c:42, m<3 -> -1 (:47) */
    {"aoso4", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 5, 0}, /* aoso4.c:42, m<5 -> -1 (:47) */
    {"aoso6", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 7, 0}, /* aoso6.c:42, m<7 -> -1 (:47) */
    {"faoso2", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 3, 0}, /* faoso2.c:47, m<3 -> -1 (:56) */
    {"faoso4", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 5, 0}, /* faoso4.c:47, m<5 -> -1 (:56) */
    {"svm0", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0}, /* svm0.c:36 */
    {"blim", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0}, /* blim.c:30 */
    {"ufndmq2", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 2, 0}, /* ufndmq2.c:31, m<2 -> -1 (:36) */
    {"ufndmq4", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 4, 0}, /* ufndmq4.c:29, m<4 -> -1 (:34) */
    {"ufndmq6", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 6, 0}, /* ufndmq6.c:29, m<6 -> -1 (:34) */
    {"ufndmq8", SB_FAM_BITPAR, SB_ANCHOR_PREFIX, 0, 0, 0, 8, 0}, /* ufndmq8.c:29, m<8 -> -1 (:34) */
    {"aut", SB_FAM_AUT, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0}, /* aut.c:26 */
    {"smoa", SB_FAM_TWOWAY, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0}, /* smoa.c:55 */
    {"tw", SB_FAM_TWOWAY, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0}, /* tw.c:92 */
    {"rcolussi", SB_FAM_TWOWAY, SB_ANCHOR_PREFIX, 0, 0, 0, 1, 0}, /* rcolussi.c:96 */
    {"bndmq2", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* bndmq2.c:31, m<2 -> -1 (:35) */
    {"bndmq4", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 4, 0}, /* bndmq4.c:29, m<4 -> -1 (:33) */
    {"bndmq6", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 6, 0}, /* bndmq6.c:29, m<6 -> -1 (:33) */
    {"lbndm", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* lbndm.c:42, m<2 -> -1 (:49) */
    {"sbndm", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* sbndm.c:28, m<2 -> -1 (:36) */
    {"sbndm-bmh", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* sbndm-bmh.c:27, m<2 -> -1 (:34) */
    {"fsbndm-w2", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* fsbndm-w2.c:47 */
    {"fndm", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* fndm.c:28 */
    {"tndm", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* tndm.c:28 */
    {"tndma", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* tndma.c:28 */
    {"bxs", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* bxs.c:30 */
    {"bxs2", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* bxs2.c:30, m<2 -> -1 (:33) */
    {"bxs4", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 4, 0}, /* bxs4.c:30, m<4 -> -1 (:33) */
    {"bxs6", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 6, 0}, /* bxs6.c:30, m<6 -> -1 (:33) */
    {"bxs8", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 8, 0}, /* bxs8.c:30, m<8 -> -1 (:33) */
    {"bww", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* bww.c:28 */
    {"dbww", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 2}, /* dbww.c:28, n/min(m,16)<2 -> -1 (:33-35) */
    {"dbww2", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 2}, /* dbww2.c:31, n/min(m,16)<2 -> -1 (:36-38) */
    {"bom", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* bom.c:77 */
    {"bom2", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* bom2.c:28 */
    {"ebom", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* ebom.c:28, m<2 -> -1 (:35) */
    {"fbom", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* fbom.c:28 */
    {"sebom", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* sebom.c:25, m<2 -> -1 (:32) */
    {"sfbom", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* sfbom.c:25 */
    {"ildm1", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* ildm1.c:29 */
    {"ildm2", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* ildm2.c:40 */
    {"fdm", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* fdm.c:27 */
    {"bsdm", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* bsdm.c:30 */
    {"bsdm4", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 4, 0}, /* bsdm4.c:32, m<4 -> -1 (:36) */
    {"bsdm5", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 5, 0}, /* bsdm5.c:32, m<5 -> -1 (:36) */
    {"bsdm7", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 7, 0}, /* bsdm7.c:32, m<7 -> -1 (:36) */
    {"bsdm8", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 8, 0}, /* bsdm8.c:32, m<8 -> -1 (:36) */
    {"ww", SB_FAM_FACTOR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* ww.c:43 */
    {"wfr", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* wfr.c:61 */
    {"wfr2", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* wfr2.c:61, m<2 -> -1 (:65) */
    {"wfr3", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 3, 0}, /* wfr3.c:60, m<3 -> -1 (:64) */
    {"wfr4", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 4, 0}, /* wfr4.c:60, m<4 -> -1 (:64) */
    {"wfr5", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 5, 0}, /* wfr5.c:59, m<5 -> -1 (:63) */
    {"wfr6", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 6, 0}, /* wfr6.c:60, m<6 -> -1 (:64) */
    {"wfr7", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 7, 0}, /* wfr7.c:61, m<7 -> -1 (:65) */
    {"wfr8", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 8, 0}, /* wfr8.c:60, m<8 -> -1 (:64) */
    {"wfrq2", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 2, 0, 2, 0}, /* wfrq2.c:68, m<2 -> -1 (:72) */
    {"wfrq3", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 3, 0, 3, 0}, /* wfrq3.c:67, m<3 -> -1 (:71) */
    {"wfrq4", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 4, 0, 4, 0}, /* wfrq4.c:67, m<4 -> -1 (:71) */
    {"wfrq5", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 5, 0, 5, 0}, /* wfrq5.c:66, m<5 -> -1 (:70) */
    {"wfrq6", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 6, 0, 6, 0}, /* wfrq6.c:67, m<6 -> -1 (:71) */
    {"wfrq7", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 7, 0, 7, 0}, /* wfrq7.c:68, m<7 -> -1 (:72) */
    {"wfrq8", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 8, 0, 8, 0}, /* wfrq8.c:67, m<8 -> -1 (:71) */
    {"twfr", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* twfr.c:42 */
    {"twfr2", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* twfr2.c:41, m<2 -> -1 (:45) */
    {"twfr3", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 3, 0}, /* twfr3.c:41, m<3 -> -1 (:45) */
    {"twfr4", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 4, 0}, /* twfr4.c:41, m<4 -> -1 (:45) */
    {"twfr5", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 5, 0}, /* twfr5.c:41, m<5 -> -1 (:45) */
    {"twfr6", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 6, 0}, /* twfr6.c:41, m<6 -> -1 (:45) */
    {"twfr7", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 7, 0}, /* twfr7.c:41, m<7 -> -1 (:45) */
    {"twfr8", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 8, 0}, /* twfr8.c:41, m<8 -> -1 (:45) */
    {"twfrq2", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 2, 0, 2, 0}, /* twfrq2.c:49, m<2 -> -1 (:53) */
    {"twfrq3", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 3, 0, 3, 0}, /* twfrq3.c:48, m<3 -> -1 (:52) */
    {"twfrq4", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 4, 0, 4, 0}, /* twfrq4.c:48, m<4 -> -1 (:52) */
    {"twfrq5", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 5, 0, 5, 0}, /* twfrq5.c:48, m<5 -> -1 (:52) */
    {"twfrq6", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 6, 0, 6, 0}, /* twfrq6.c:48, m<6 -> -1 (:52) */
    {"twfrq7", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 7, 0, 7, 0}, /* twfrq7.c:48, m<7 -> -1 (:52) */
    {"twfrq8", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 8, 0, 8, 0}, /* twfrq8.c:48, m<8 -> -1 (:52) */
    {"lwfr2", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* lwfr2.c:60, m<2 -> -1 (:65) */
    {"lwfr3", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 3, 0}, /* lwfr3.c:60, m<3 -> -1 (:65) */
    {"lwfr4", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 4, 0}, /* lwfr4.c:60, m<4 -> -1 (:65) */
    {"lwfr5", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 5, 0}, /* lwfr5.c:60, m<5 -> -1 (:65) */
    {"lwfr6", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 6, 0}, /* lwfr6.c:60, m<6 -> -1 (:65) */
    {"lwfr7", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 7, 0}, /* lwfr7.c:60, m<7 -> -1 (:65) */
    {"lwfr8", SB_FAM_WFR, SB_ANCHOR_SUFFIX, 0, 0, 0, 8, 0}, /* lwfr8.c:60, m<8 -> -1 (:65) */
    {"skip3", SB_FAM_SKIPQ, SB_ANCHOR_SUFFIX, 0, 0, 0, 3, 0}, /* skip3.c:28, m<3 -> -1 (:31) */
    {"skip5", SB_FAM_SKIPQ, SB_ANCHOR_SUFFIX, 0, 0, 0, 5, 0}, /* skip5.c:28, m<5 -> -1 (:31) */
    {"skip6", SB_FAM_SKIPQ, SB_ANCHOR_SUFFIX, 0, 0, 0, 6, 0}, /* skip6.c:28, m<6 -> -1 (:31) */
    {"skip7", SB_FAM_SKIPQ, SB_ANCHOR_SUFFIX, 0, 0, 0, 7, 0}, /* skip7.c:28, m<7 -> -1 (:31) */
    {"skip8", SB_FAM_SKIPQ, SB_ANCHOR_SUFFIX, 0, 0, 0, 8, 0}, /* skip8.c:28, m<8 -> -1 (:31) */
    {"bfs", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* bfs.c:64 */
    {"ffs", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* ffs.c:52 */
    {"iom", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* iom.c:32 */
    {"wom", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* wom.c:57 */
    {"pbmh", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* pbmh.c:29 */
    {"qlqs", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* qlqs.c:38 */
    {"tsw", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* tsw.c:40 */
    {"wc", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* wc.c:68 */
    {"zt", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* zt.c:76 */
    {"tvsbs-w2", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* tvsbs-w2.c:36 */
    {"fs-w1", SB_FAM_BADCHAR, SB_ANCHOR_SUFFIX, 0, 0, 0, 1, 0}, /* fs-w1.c:65 */
    {"ssef", SB_FAM_SSEF, SB_ANCHOR_SUFFIX, 0, 0, 0, 32, 0}, /* ssef.c:40, m<32 -> -1 (:41) */
    /* sbndm2 is the ONE algorithm the reference enables by default (src/algorithms.h:133), so a drop-in must know the
     * name.  Its own output is the exact count for m <= 32; for m > 32 its search_large works on the 32-character
     * prefix and is unreliable: periodic patterns lose occurrences (a^33 in a^65536 gives 2185 instead of
     * 65504) and a pattern that shares only that prefix with the text can be counted (bible64k, m = 33, last byte
     * changed: 1 instead of 0).  We return the exact count there - a documented divergence from the reference's
     * (wrong) answer, never from bf.c. */
    {"sbndm2", SB_FAM_BNDMX, SB_ANCHOR_SUFFIX, 0, 0, 0, 2, 0}, /* sbndm2.c:27, m<2 -> -1 (:30), search_large :77 */
};
#define SB_NALGOS ((int)(sizeof(g_algos) / sizeof(g_algos[0])))

int sb_algo_count(void) { return SB_NALGOS; }

const sb_algo_desc *sb_algo_desc_get(int algo)
{
    return (algo >= 0 && algo < SB_NALGOS) ? &g_algos[algo] : NULL;
}

const char *sb_algo_name(int algo)
{
    const sb_algo_desc *d = sb_algo_desc_get(algo);
    return d ? d->name : NULL;
}

int sb_algo_id(const char *name)
{
    if (!name) return -1;
    for (int i = 0; i < SB_NALGOS; i++) {
        const char *a = g_algos[i].name, *b = name;
        /* case-insensitive like str2lower() in function.h:79 */
        while (*a && *b && (*a == *b || (*b >= 'A' && *b <= 'Z' && *b + 32 == *a))) a++, b++;
        if (!*a && !*b) return i;
    }
    return -1;
}

int sb_algo_applicable(int algo, int m)
{
    const sb_algo_desc *d = sb_algo_desc_get(algo);
    if (!d) return -1;
    return m >= d->min_m ? 0 : -1;
}

/* ------------------------------------------------------------------ border tables */

/* mpNext[i] = length of the longest proper border of x[0..i), mpNext[0] = -1  (mp.c:26-35) */
void sb_prep_mp(const unsigned char *x, int m, int *mpNext)
{
    mpNext[0] = -1;
    int b = -1;
    for (int len = 1; len <= m; len++) {
        unsigned char c = x[len - 1];
        while (b >= 0 && x[b] != c) b = mpNext[b];
        mpNext[len] = ++b;
    }
}

/* kmpNext[i] (i<m) = longest border b of x[0..i) whose continuation differs, x[b] != x[i];
 * -1 if none. kmpNext[m] = plain border of the whole pattern  (kmp.c:27-41).
 * Derived from the plain border table: follow it while the continuation is equal. */
void sb_prep_kmp(const unsigned char *x, int m, const int *mpNext, int *kmpNext)
{
    kmpNext[0] = -1;
    for (int i = 1; i < m; i++) {
        int b = mpNext[i];
        kmpNext[i] = (x[b] == x[i]) ? kmpNext[b] : b;
    }
    if (m >= 1) kmpNext[m] = mpNext[m];
}

/* ------------------------------------------------------------------ bad-character tables */

/* hbc[c] = distance from the last occurrence of c in x[0..m-1) to position m-1; m if absent.
 * (hor.c:26-30, bm.c:27-33, raita.c:26-32, tunedbm.c:26-32) */
void sb_prep_hbc(const unsigned char *x, int m, int *hbc)
{
    for (int c = 0; c < SB_SIGMA; c++) hbc[c] = m;
    for (int i = m - 2, seen = 0; i >= 0 && seen < SB_SIGMA; i--)
        if (hbc[x[i]] == m) {
            hbc[x[i]] = m - 1 - i;
            seen++;
        }
}

/* qsBc[c] = m - (last index of c in x); m+1 if absent  (qs.c:26-30) */
void sb_prep_qsbc(const unsigned char *x, int m, int *qsBc)
{
    for (int c = 0; c < SB_SIGMA; c++) qsBc[c] = m + 1;
    for (int i = m - 1, seen = 0; i >= 0 && seen < SB_SIGMA; i--)
        if (qsBc[x[i]] == m + 1) {
            qsBc[x[i]] = m - i;
            seen++;
        }
}

/* ------------------------------------------------------------------ Boyer-Moore good suffix */

/* suff[i] = length of the longest common suffix of x[0..i] and x  (bm.c:36-52).
 * Computed as the Z-function of the reversed pattern: suff[i] = Z_rev[m-1-i]. */
void sb_prep_suffixes(const unsigned char *x, int m, int *suff)
{
    if (m <= 0) return;
    /* z[k] = longest common prefix of r and r[k..), r = reverse(x); r[k] = x[m-1-k] */
    suff[m - 1] = m;
    int l = 0, r = 0; /* current Z-box [l, r) in reversed coordinates */
    for (int k = 1; k < m; k++) {
        int z = 0;
        if (k < r) {
            int zk = suff[m - 1 - (k - l)];
            z = zk < r - k ? zk : r - k;
        }
        while (k + z < m && x[m - 1 - z] == x[m - 1 - (k + z)]) z++;
        if (k + z > r) {
            l = k;
            r = k + z;
        }
        suff[m - 1 - k] = z;
    }
}

/* bmGs[i] = shift after a mismatch at x[i] with x[i+1..m) matched  (bm.c:54-66).
 * case 1: a prefix of x of length i+1 that is also a suffix (suff[i]==i+1) bounds the shift
 *         for every j < m-1-i still unset, scanning i downward;
 * case 2: a re-occurrence of the suffix of length suff[i] ending at i gives m-1-i for the
 *         mismatch position m-1-suff[i]; later (larger) i overwrite, giving the smallest shift. */
void sb_prep_bmgs(const unsigned char *x, int m, int *suff, int *bmGs)
{
    sb_prep_suffixes(x, m, suff);
    for (int j = 0; j < m; j++) bmGs[j] = m;
    int j = 0;
    for (int i = m - 1; i >= 0; i--) {
        if (suff[i] != i + 1) continue;
        int shift = m - 1 - i;
        while (j < shift) {
            if (bmGs[j] == m) bmGs[j] = shift;
            j++;
        }
    }
    for (int i = 0; i + 1 < m; i++) bmGs[m - 1 - suff[i]] = m - 1 - i;
}

/* ------------------------------------------------------------------ bit-parallel masks */

/* Shift-Or: bit i of S[c] is 0 iff x[i]==c, over the first min(m,32) characters;
 * lim has the bits >= w-1 set, w = min(m,32)  (so.c:27-38; search_large keeps 32 chars, so.c:74-80) */
unsigned int sb_prep_so(const unsigned char *x, int m, unsigned int *S)
{
    int w = m < SB_WORD ? m : SB_WORD;
    for (int c = 0; c < SB_SIGMA; c++) S[c] = ~0u;
    for (int i = 0; i < w; i++) S[x[i]] &= ~(1u << i);
    unsigned int used = (w == 32) ? ~0u : ((1u << w) - 1u);
    return ~(used >> 1);
}

/* Shift-And: bit i of S[c] is 1 iff x[i]==c; F = bit (w-1)  (sa.c:27-34, :45) */
unsigned int sb_prep_sa(const unsigned char *x, int m, unsigned int *S)
{
    int w = m < SB_WORD ? m : SB_WORD;
    memset(S, 0, SB_SIGMA * sizeof(unsigned int));
    for (int i = 0; i < w; i++) S[x[i]] |= 1u << i;
    return 1u << (w - 1);
}

/* Bucket lists of Skip-Search and KMP-Skip-Search: head[c] = right-most index i with x[i] == c (-1 if none),
 * next[i] = the next index to the left holding the same character (-1 at the end): the order in which the
 * reference's linked lists are traversed (skip.c:31-39 pushes front; kmpskip.c:71-77 z[]/list[]). */
void sb_prep_buckets(const unsigned char *x, int m, int *head, int *next)
{
    for (int c = 0; c < SB_SIGMA; c++) head[c] = -1;
    for (int i = 0; i < m; i++) {
        next[i] = head[x[i]];
        head[x[i]] = i;
    }
}

/* bc[c] = m-1 - (right-most index of c in the WHOLE pattern), m if absent: unlike hbc the last character
 * counts, so bc[x[m-1]] == 0 stops the skip loop (fs.c:50-51, graspm.c:47-48, ffs.c:60-61) */
void sb_prep_lastocc(const unsigned char *x, int m, int *bc)
{
    for (int c = 0; c < SB_SIGMA; c++) bc[c] = m;
    for (int j = 0; j < m; j++) bc[x[j]] = m - 1 - j;
}

/* Fast-Search good-suffix table gs[0..m] (fs.c:26-41 Pre_GS): the classical Knuth/BM construction through
 * the border array f of the reversed pattern.  Entries 1..m are used by the scan (gs[m-j+1]). */
void sb_prep_fsgs(const unsigned char *x, int m, int *gs)
{
    static _Thread_local int f[SB_XSIZE + 2];
    for (int i = 0; i <= m; i++) gs[i] = 0;
    int j = m + 1;
    f[m] = j;
    for (int i = m; i > 0; i--) {
        while (j <= m && x[i - 1] != x[j - 1]) {
            if (gs[j] == 0) gs[j] = j - i;
            j = f[j];
        }
        f[i - 1] = --j;
    }
    int p = f[0];
    for (j = 0; j <= m; j++) {
        if (gs[j] == 0) gs[j] = p;
        if (j == p) p = f[p];
    }
}

/* GRASPm candidate lists (graspm.c:41-45): for a text character c seen just before an occurrence of the
 * pattern's last character, the alignments k to try: every k = i+1 with x[i] == c and x[i+1] == x[m-1]
 * (right-most first), followed by k = 0 for EVERY c when x[0] == x[m-1].  Encoded as head[c] / next[k],
 * -1 terminated; the shared tail element 0 has next[0] == -1. */
void sb_prep_graspm(const unsigned char *x, int m, int *head, int *next)
{
    const int tail = (x[0] == x[m - 1]) ? 0 : -1;
    for (int c = 0; c < SB_SIGMA; c++) head[c] = tail;
    next[0] = -1;
    for (int i = 0; i + 1 < m; i++)
        if (x[i + 1] == x[m - 1]) {
            next[i + 1] = head[x[i]];
            head[x[i]] = i + 1;
        }
}

/* BNDM: bit k of B[c] is set iff x[w-1-k] == c, over the first w = min(m,32) characters: the masks of the
 * REVERSED prefix (bndm.c:33-38; search_large keeps 32 characters, bndm.c:70-76) */
void sb_prep_bndm(const unsigned char *x, int m, unsigned int *B)
{
    const int w = m < SB_WORD ? m : SB_WORD;
    memset(B, 0, SB_SIGMA * sizeof(unsigned int));
    for (int k = 0; k < w; k++) B[x[w - 1 - k]] |= 1u << k;
}

/* Berry-Ravindran: brBc[a][b] = shift aligning the two text characters after the window with their
 * right-most occurrence in the pattern (br.c:26-37, tvsbs.c:26-37).  Filled by decreasing priority:
 * default m+2; b == x[0] -> m+1; pair x[i]x[i+1] -> m-i (right-most pair wins); a == x[m-1] -> 1. */
void sb_prep_brbc(const unsigned char *x, int m, int *brBc)
{
    for (int i = 0; i < SB_SIGMA * SB_SIGMA; i++) brBc[i] = m + 2;
    for (int a = 0; a < SB_SIGMA; a++) brBc[a * SB_SIGMA + x[0]] = m + 1;
    for (int i = 0; i + 1 < m; i++) brBc[x[i] * SB_SIGMA + x[i + 1]] = m - i;
    for (int b = 0; b < SB_SIGMA; b++) brBc[x[m - 1] * SB_SIGMA + b] = 1;
}

/* ------------------------------------------------------------------ q-gram hash shift (HASHq) */

/* h(i) = (sum_{k<q} x[i-k] << k) mod 256: the reference computes it either in an unsigned char
 * (hash3.c:29,42-44) or in an unsigned int reduced %WSIZE (hash5.c:46, hash8.c:52); both are the
 * same value because every step is linear mod 256. */
static unsigned sb_hashq_at(const unsigned char *s, int i, int q)
{
    unsigned h = 0;
    for (int k = q - 1; k >= 0; k--) h = (h << 1) + s[i - k];
    return h & (SB_WSIZE - 1);
}

/* shift[h] = m-1-i for the right-most q-gram ending at i < m-1 with hash h, default m-q+1;
 * the q-gram ending at m-1 gets shift 0 and its previous value becomes sh1 (>=1)
 * (hash3.c:37-55, hash5.c:39-65, hash8.c:38-71). */
void sb_prep_hashq(const unsigned char *x, int m, int q, int *shift, int *sh1)
{
    for (int h = 0; h < SB_WSIZE; h++) shift[h] = m - q + 1;
    for (int i = q - 1; i < m - 1; i++) shift[sb_hashq_at(x, i, q)] = m - 1 - i;
    unsigned last;
    if (m > q) {
        last = sb_hashq_at(x, m - 1, q);
    } else {
        /* m == q: the reference's loop index is left at q, not m-1, so its "last" q-gram is
         * x[1..q] and includes the terminator x[m] == 0 the driver stores (smart.c:313);
         * the first q-gram already has shift m-q == 0. Reproduced here without reading x[m]. */
        unsigned char tmp[16];
        shift[sb_hashq_at(x, q - 1, q)] = 0; /* = m - q, hash3.c:45 */
        memcpy(tmp, x + 1, (size_t)(q - 1));
        tmp[q - 1] = 0;
        last = sb_hashq_at(tmp, q - 1, q);
    }
    *sh1 = shift[last] ? shift[last] : 1;
    shift[last] = 0;
}

/* ------------------------------------------------------------------ Karp-Rabin */

/* d = 2^(m-1), hx = sum x[i] * 2^(m-1-i), both in wrapping 32-bit arithmetic  (kr.c:34-40) */
void sb_prep_kr(const unsigned char *x, int m, int *d, int *hx)
{
    unsigned dd = (m - 1 >= 32) ? 0u : (1u << (m - 1));
    unsigned h = 0;
    for (int i = 0; i < m; i++) h = (h << 1) + x[i];
    *d = (int)dd;
    *hx = (int)h;
}

/* ------------------------------------------------------------------ Q-gram filter (QF) */

/* B[g] has bit ((m-i) mod Q) set for every pattern q-gram starting at i (0 <= i <= m-Q) whose
 * fingerprint g = (sum_{k<Q} x[i+k] << (S*k)) & mask  (qf23.c:50-56). */
void sb_prep_qf(const unsigned char *x, int m, int Q, int S, int *B)
{
    int asize = 1 << (Q * S);
    unsigned mask = (unsigned)asize - 1u;
    memset(B, 0, (size_t)asize * sizeof(int));
    for (int i = 0; i + Q <= m; i++) {
        unsigned g = 0;
        for (int k = Q - 1; k >= 0; k--) g = (g << S) + x[i + k];
        B[g & mask] |= 1 << ((m - i) % Q);
    }
}

/* Zhu-Takaoka two-character shift (zt.c:64-74), row-major [a][b]: how far the window may move when its last two
 * characters are a, b: m if the pair does not occur (m-1 if b is the first pattern character), else the distance
 * of the pair's rightmost occurrence that ends before the last pattern character from the window end. */
void sb_prep_ztbc(const unsigned char *x, int m, int *ztBc)
{
    for (int i = 0; i < 65536; i++) ztBc[i] = m;
    for (int a = 0; a < 256; a++) ztBc[a * 256 + x[0]] = m - 1;
    for (int i = 1; i < m - 1; i++) ztBc[x[i - 1] * 256 + x[i]] = m - 1 - i;
}

/* Skip Search on q-grams (skip3.c:33-42, skip5.c:36-45): pattern offset i is filed under the hash of x[i..i+Q);
 * head[h] = the last offset filed under h (-1: none), next[i] = the one filed before i.  skip3 hashes
 * x[i]<<4 + x[i+1]<<2 + x[i+2] (skip3.c:25), skipQ for Q >= 5 sum_k x[i+k] << (Q-1-k) (skip5.c:25): < 65536. */
static unsigned skipq_hash(const unsigned char *p, int Q)
{
    if (Q == 3) return ((unsigned)p[0] << 4) + ((unsigned)p[1] << 2) + p[2];
    unsigned h = 0;
    for (int k = 0; k < Q; k++) h += (unsigned)p[k] << (Q - 1 - k);
    return h;
}

void sb_prep_skipq(const unsigned char *x, int m, int Q, int *head, int *next)
{
    for (int i = 0; i < 65536; i++) head[i] = -1;
    for (int i = 0; i + Q <= m; i++) {
        const unsigned h = skipq_hash(x + i, Q);
        next[i] = head[h];
        head[h] = i;
    }
}

/* Backward SNR DAWG Matching (bsdm.c:33-55, bsdm4.c:38-66): the longest run of pattern q-grams without a repeated
 * q-gram (its first gram index st, its length len in grams) and pos[h] = index inside that run of the gram that
 * hashes to h (-1: not in the run).  Q = 1: single characters (pos[256] in the reference, the first 256 entries
 * here); the hash shifts by 2 bits per character for Q = 4, by 1 bit for Q = 5, 7, 8 (bsdm4.c:29, bsdm5.c:29). */
static unsigned bsdm_hash(const unsigned char *p, int Q)
{
    const int step = Q == 4 ? 2 : 1;
    unsigned h = 0;
    for (int t = 0; t < Q; t++) h += (unsigned)p[t] << (step * (Q - 1 - t));
    return h & 0xffffu;
}

void sb_prep_bsdm(const unsigned char *x, int m, int Q, int *pos, int *st_out, int *len_out)
{
    static _Thread_local unsigned char occ[65536];
    memset(occ, 0, sizeof occ);
    int st = 0, len = 0;
    for (int i = 0, j = 0; i < m - Q + 1; i++) {
        const unsigned c = bsdm_hash(x + i, Q);
        if (occ[c]) {
            unsigned d = bsdm_hash(x + j, Q);
            while (d != c) {
                occ[d] = 0;
                j++;
                d = bsdm_hash(x + j, Q);
            }
            occ[d] = 0;
            j++;
        }
        occ[c] = 1;
        if (len < i - j + 1) {
            len = i - j + 1;
            st = j;
        }
    }
    for (int i = 0; i < 65536; i++) pos[i] = -1;
    for (int i = 0; i < len; i++) pos[bsdm_hash(x + st + i, Q)] = i;
    *st_out = st;
    *len_out = len;
}

/* Weak-factor-recognition table (wfr.c:41-58, wfrq4.c:50-65): F[h] = 1 for the 16-bit hash h of every pattern
 * factor the scan can ask about.  The hash is built right to left, as the scan reads the window.
 * Q <= 1 (wfr, wfr2..8, twfr*, lwfr*): every factor x[j..i] of at most 16 characters, h <- (h << 2) + x[j].
 * Q >= 2 (wfrqQ, twfrqQ): factors made of whole q-grams ending at i, i-Q, i-2Q, ... and not reaching further
 * back than min(m, 8) characters; a q-gram ending at e hashes to g(e) = sum_k x[e-k] << 2(Q-1-k) and
 * h <- (h << 2Q) + g(e).  All arithmetic modulo 2^16 (the reference keeps h in an unsigned short). */
void sb_prep_wfr(const unsigned char *x, int m, int Q, unsigned char *F)
{
    memset(F, 0, 65536);
    if (Q <= 1) {
        const int fact = m < 16 ? m : 16;
        for (int i = 0; i < m; i++) {
            const int stop = i - fact + 1 > 0 ? i - fact + 1 : 0;
            unsigned h = 0;
            for (int j = i; j >= stop; j--) {
                h = ((h << 2) + x[j]) & 0xffffu;
                F[h] = 1;
            }
        }
        return;
    }
    const int fact = m < 8 ? m : 8;
    for (int i = Q - 1; i < m; i++) {
        const int stop = i - fact + 1 >= Q - 1 ? i - fact + 1 : Q - 1;
        unsigned h = 0;
        for (int e = i; e == i || e >= stop; e -= Q) {
            unsigned g = 0;
            for (int k = 0; k < Q; k++) g += (unsigned)x[e - k] << (2 * (Q - 1 - k));
            h = (e == i ? g : (h << (2 * Q)) + g) & 0xffffu;
            F[h] = 1;
        }
    }
}

/* ------------------------------------------------------------------ all tables of a family */

/* String-matching automaton of aut.c:26-45 (table built by preSMA, include/AUTOMATON.h:185-198): state q = length of
 * the longest prefix of x that is a suffix of the text read so far; delta[q*256 + c] = next state, q in [0, m].
 * From the definition: delta(q, c) = q+1 if q < m and x[q] == c, else delta(border(q), c), delta(0, c) = [c == x[0]];
 * border(q) = mpNext[q] (mp.c:26-35).  States fit 16 bits (m <= 4200). */
int sb_prep_automaton(const unsigned char *x, int m, unsigned short *delta)
{
    static _Thread_local int border[SB_XSIZE + 1];
    if (!x || !delta || m < 1 || m > SB_XSIZE) return -2;
    sb_prep_mp(x, m, border);
    for (int c = 0; c < SB_SIGMA; c++) delta[c] = 0;
    delta[x[0]] = 1;
    for (int q = 1; q <= m; q++) {
        const unsigned short *fall = delta + (size_t)border[q] * SB_SIGMA;
        unsigned short *row = delta + (size_t)q * SB_SIGMA;
        for (int c = 0; c < SB_SIGMA; c++) row[c] = fall[c];
        if (q < m) row[x[q]] = (unsigned short)(q + 1);
    }
    return 0;
}

/* Factor oracle of bom.c:49-70 (Backward Oracle Matching) as a dense table.  State i (0 <= i <= m) stands for the
 * suffix x[i..m); the oracle is grown from state m down to 0: state i-1 is entered from i by x[i-1], and every state
 * on the supply path of i that has no transition by x[i-1] yet gets one to i-1 (the online construction of the
 * factor oracle, here on the reversed reading direction the backward scan needs).  trans[p*256 + c] = target state or
 * -1 (bom.c:28-39 getTransition), term[p] = 1 for the states on the supply path of state 0 (bom.c:65-69): reaching
 * one of them means a prefix of x ends here.  supply[] is scratch of m+2 ints. */
int sb_prep_bom(const unsigned char *x, int m, short *trans, unsigned char *term)
{
    static _Thread_local int supply[SB_XSIZE + 2];
    if (!x || !trans || !term || m < 1 || m > SB_XSIZE) return -2;
    for (size_t i = 0; i < (size_t)(m + 1) * SB_SIGMA; i++) trans[i] = -1;
    for (int i = 0; i <= m; i++) term[i] = 0;
    supply[m] = m + 1;  /* "no state" */
    for (int i = m; i > 0; --i) {
        const unsigned char c = x[i - 1];
        trans[(size_t)i * SB_SIGMA + c] = (short)(i - 1);
        int p = supply[i], q = -1;
        while (p <= m && (q = trans[(size_t)p * SB_SIGMA + c]) < 0) {
            trans[(size_t)p * SB_SIGMA + c] = (short)(i - 1);
            p = supply[p];
        }
        supply[i - 1] = p == m + 1 ? m : q;
    }
    for (int p = 0; p <= m; p = supply[p]) term[p] = 1;
    return 0;
}

/* sbndm2.c:33-47 (m <= 32): B[c] has bit b set iff x[m-1-b] == c (the pattern read right to left), and the shift
 * after an occurrence, m - j for the longest j <= m-1 such that the j characters before the last one, x[m-1-j .. m-2],
 * are a prefix of x.  (j = m-1 always qualifies - x[0 .. m-2] is a prefix of x - so the reference's shift is 1.) */
int sb_prep_sbndm2(const unsigned char *x, int m, unsigned int *B, int *shift)
{
    if (!x || !B || !shift || m < 2 || m > SB_WORD) return -2;
    for (int c = 0; c < SB_SIGMA; c++) B[c] = 0;
    for (int b = 0; b < m; b++) B[x[m - 1 - b]] |= 1u << b;
    int best = 0;
    for (int j = 1; j <= m - 1; j++)
        if (memcmp(x + (m - 1 - j), x, (size_t)j) == 0) best = j;
    *shift = best ? m - best : 0;
    return 0;
}

int sb_prep_tables(int algo, const unsigned char *x, int m, sb_tables *out)
{
    const sb_algo_desc *d = sb_algo_desc_get(algo);
    if (!d || !x || !out || m < 1 || m > SB_XSIZE) return -2;
    /* everything up to the first big array is cleared, and so are the scalars that lie between the big arrays;
     * the big arrays themselves (brBc, qfB, wfrF, ztBc: 256 KiB - 1 MiB each) are only DEFINED for the families
     * that fill them - a reused struct may still hold another family's values there */
    memset(out, 0, offsetof(sb_tables, brBc));
    out->qfQ = out->qfS = out->wfrQ = 0;
    out->m = m;
    if (m < d->min_m) return -1;
    switch (d->family) {
    case SB_FAM_BF: break;
    case SB_FAM_MP: sb_prep_mp(x, m, out->mpNext); break;
    case SB_FAM_KMP:
        sb_prep_mp(x, m, out->mpNext);
        sb_prep_kmp(x, m, out->mpNext, out->kmpNext);
        break;
    case SB_FAM_BM:
        sb_prep_bmgs(x, m, out->suff, out->bmGs);
        sb_prep_hbc(x, m, out->hbc);
        break;
    case SB_FAM_HOR:
    case SB_FAM_RAITA:
    case SB_FAM_TUNEDBM: sb_prep_hbc(x, m, out->hbc); break;
    case SB_FAM_QS: sb_prep_qsbc(x, m, out->qsBc); break;
    case SB_FAM_SMITH:
        sb_prep_hbc(x, m, out->hbc);
        sb_prep_qsbc(x, m, out->qsBc);
        break;
    case SB_FAM_TBM:
    case SB_FAM_AG:
        sb_prep_bmgs(x, m, out->suff, out->bmGs);
        sb_prep_hbc(x, m, out->hbc);
        break;
    case SB_FAM_SKIP: sb_prep_buckets(x, m, out->bucketHead, out->bucketNext); break;
    case SB_FAM_KMPSKIP:
        sb_prep_buckets(x, m, out->bucketHead, out->bucketNext);
        sb_prep_mp(x, m, out->mpNext);
        sb_prep_kmp(x, m, out->mpNext, out->kmpNext);
        break;
    case SB_FAM_FS:
        sb_prep_lastocc(x, m, out->hbc);
        sb_prep_fsgs(x, m, out->fsGs);
        break;
    case SB_FAM_GRASPM:
        sb_prep_lastocc(x, m, out->hbc);
        sb_prep_graspm(x, m, out->bucketHead, out->bucketNext);
        break;
    case SB_FAM_BNDM: sb_prep_bndm(x, m, out->bndmB); break;
    case SB_FAM_BR: sb_prep_brbc(x, m, out->brBc); break;
    case SB_FAM_FJS:
        sb_prep_qsbc(x, m, out->qsBc);
        sb_prep_mp(x, m, out->mpNext);
        sb_prep_kmp(x, m, out->mpNext, out->kmpNext);
        break;
    case SB_FAM_AC: {
        /* ell = length of the leading run of x[0] if the pattern is not a single run, else 0 (ac.c:52-53) */
        int ell = 1;
        sb_prep_mp(x, m, out->mpNext);
        sb_prep_kmp(x, m, out->mpNext, out->kmpNext);
        while (ell < m && x[ell] == x[0]) ell++;
        out->acEll = (ell == m) ? 0 : ell;
        break;
    }
    case SB_FAM_SO: out->soLim = sb_prep_so(x, m, out->soS); break;
    case SB_FAM_SA: out->saF = sb_prep_sa(x, m, out->saS); break;
    case SB_FAM_KR: sb_prep_kr(x, m, &out->krD, &out->krHx); break;
    case SB_FAM_HASHQ:
        out->hashQ = d->q;
        sb_prep_hashq(x, m, d->q, out->hashShift, &out->hashSh1);
        break;
    case SB_FAM_QF:
        out->qfQ = d->Q;
        out->qfS = d->S;
        sb_prep_qf(x, m, d->Q, d->S, out->qfB);
        break;
    case SB_FAM_WFR:
        out->wfrQ = d->Q;
        sb_prep_wfr(x, m, d->Q, out->wfrF);
        if (!strncmp(d->name, "lwfr", 4)) { /* lwfr4.c:36 preKmp: the linear variant also carries kmpNext */
            sb_prep_mp(x, m, out->mpNext);
            sb_prep_kmp(x, m, out->mpNext, out->kmpNext);
        }
        break;
    case SB_FAM_BADCHAR: /* members whose tables are the reference's own copies of builders above */
        if (!strcmp(d->name, "tsw")) sb_prep_brbc(x, m, out->brBc);          /* tsw.c preBrBc == br.c:26-37 */
        else if (!strcmp(d->name, "fs-w1")) sb_prep_fsgs(x, m, out->fsGs);   /* fs-w1.c Pre_GS == fs.c Pre_GS */
        else if (!strcmp(d->name, "qlqs")) sb_prep_qsbc(x, m, out->qsBc);    /* qlqs.c:32-36 == qs.c:26-30 (forward table) */
        else if (!strcmp(d->name, "zt")) {                                   /* zt.c suffixes/preBmGs == bm.c:36-66 */
            sb_prep_suffixes(x, m, out->suff);
            sb_prep_bmgs(x, m, out->suff, out->bmGs);
            sb_prep_ztbc(x, m, out->ztBc);                                   /* zt.c:64-74 */
        }
        break;
    default: break; /* the other second-batch families expose no tables yet */
    }
    return 0;
}
