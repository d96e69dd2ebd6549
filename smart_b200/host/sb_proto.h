/* Wire protocol between the thin per-algorithm executables and the resident GPU worker (smartd).
 * One request / one reply per search over a Unix stream socket.  The text and the pattern never travel
 * over the socket: they stay in the SysV segments the SMART driver created (smart.c:553-568, :236-250 in
 * the reference tree) and the worker attaches them by key, exactly as an algorithm executable would
 * (main.h:60-98). */
#ifndef SB_PROTO_H
#define SB_PROTO_H
#include <stdint.h>

#define SB_PROTO_MAGIC 0x53423230u /* "SB20" */
enum { SB_REQ_SEARCH = 1, SB_REQ_PING = 2, SB_REQ_QUIT = 3, SB_REQ_STATS = 4, SB_REQ_SEARCH_MULTI = 5,
       /* the pattern set of one length as npat INDEPENDENT searches, pipelined with sb_search64_submit /
        * sb_search_collect (one scan per pattern, no wait between them); the result segment holds npat counts
        * (uint64), then npat scan times and npat preprocessing times (double, ms) */
       SB_REQ_SEARCH_SET = 6 };

typedef struct sb_request {
    uint32_t magic;
    int32_t kind;
    int32_t algo; /* sb_algo_id */
    int32_t pkey; /* SysV key of the pattern segment */
    int32_t m;
    int32_t tkey; /* SysV key of the text segment */
    int32_t n;
    int32_t npat; /* SEARCH_MULTI / SEARCH_SET: the pattern segment holds npat patterns of length m back to back */
    int32_t ckey; /* SEARCH_MULTI: SysV key of the per-pattern count segment (npat x uint64); SEARCH_SET: npat x 24 bytes */
    int32_t reserved;
    char algo_name[24]; /* used when algo < 0 (the driver does not link the library's registry) */
    uint64_t text_gen;  /* non-zero: the sender vouches that (tkey, n, text_gen) identifies the text content
                           (our driver exports SMART_B200_TEXT_GEN after every text load); zero: the worker
                           hashes the text to detect a change (unmodified SMART driver) */
} sb_request;

typedef struct sb_reply {
    uint32_t magic;
    int32_t status; /* 0 ok, else errno-like */
    int64_t count;  /* occurrences, -1 not applicable */
    double pre_ms, run_ms;
    double upload_ms;     /* text (re)upload paid by this request, 0 when the resident text was reused */
    uint64_t searches;    /* STATS: searches served */
    uint64_t text_uploads; /* STATS: text uploads (cache misses) */
} sb_reply;

/* default socket path: $SMART_B200_SOCK or /tmp/smart_b200.<uid>.sock */
const char *sb_socket_path(char *buf, unsigned long len);
#endif
