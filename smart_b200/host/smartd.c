/*
 * smartd — the resident GPU worker.
 *
 * SMART launches one process per search (smart.c:143 `system()`); a CUDA context and a text upload per
 * search would dwarf the scan itself.  This long-lived worker owns the CUDA context and keeps the text
 * DEVICE-RESIDENT across the driver's #algorithms x #lengths x #patterns loop: the thin per-algorithm
 * executables pass it the SysV keys they received on their command line, the worker attaches the same
 * segments, (re)uploads the text only when its content hash changes, and runs the search.
 *
 *   smartd [-sock path] [-fg]      start (daemonises unless -fg); prints "ready <path>" when listening
 *   smartd -stop [-sock path]      ask a running worker to exit
 *   smartd -stats [-sock path]     print searches served / text uploads
 */
#include "sb_client.h"
#include "sb_proto.h"
#include "smart_b200.h"

#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ipc.h>
#include <sys/shm.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* content hash of the text, every byte (used up to 64 MiB; larger texts without a generation number are
 * re-uploaded instead: at PCIe Gen5 rates the copy is faster than hashing them) */
static uint64_t text_hash(const unsigned char *t, size_t n)
{
    uint64_t a = 0x9E3779B97F4A7C15ull ^ n, c = ~a;
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        uint64_t u, v;
        memcpy(&u, t + i, 8);
        memcpy(&v, t + i + 8, 8);
        a = (a ^ u) * 0xFF51AFD7ED558CCDull;
        a = (a << 29) | (a >> 35);
        c = (c ^ v) * 0xC4CEB9FE1A85EC53ull;
        c = (c << 31) | (c >> 33);
    }
    for (; i < n; i++) a = (a ^ t[i]) * 0x100000001B3ull;
    return a ^ (c * 0x9E3779B97F4A7C15ull);
}
#define SB_HASH_LIMIT ((size_t)64 << 20)

static struct {
    int tkey, shmid;
    size_t n;
    uint64_t hash, gen;
    const unsigned char *host; /* attachment */
    sb_text *dev;
} g_text = {-1, -1, 0, 0, 0, NULL, NULL};

static struct {
    int pkey, shmid;
    const unsigned char *host;
} g_pat = {-1, -1, NULL};

static uint64_t g_searches, g_uploads;

/* Attach segment `key`, reusing the cached attachment (*host_io, *key_io, *shmid_io) when it is the same
 * segment.  On any failure the cache is CLEARED: a detached pointer must never be handed out again. */
static const unsigned char *attach(int key, size_t size, int *shmid_io, const unsigned char **host_io, int *key_io)
{
    int id = shmget((key_t)key, size, 0666);
    if (id >= 0 && *host_io && key == *key_io && id == *shmid_io) return *host_io; /* same segment still attached */
    if (*host_io) shmdt(*host_io);
    *host_io = NULL;
    *key_io = -1;
    *shmid_io = -1;
    if (id < 0) return NULL;
    void *p = shmat(id, NULL, SHM_RDONLY);
    if (p == (void *)-1) return NULL;
    *shmid_io = id;
    *key_io = key;
    *host_io = p;
    return p;
}

static int ensure_text(int tkey, int n, uint64_t gen, double *upload_ms);

static void serve_search(const sb_request *rq, sb_reply *rp)
{
    rp->status = 0;
    if (rq->m < 1 || rq->m > SB_XSIZE || rq->n < 0) {
        rp->status = EINVAL;
        return;
    }
    if (!attach(rq->pkey, (size_t)rq->m, &g_pat.shmid, &g_pat.host, &g_pat.pkey)) {
        rp->status = errno ? errno : ENOENT;
        return;
    }
    if ((rp->status = ensure_text(rq->tkey, rq->n, rq->text_gen, &rp->upload_ms)) != 0) return;
    long long c = sb_search64(rq->algo, g_text.dev, g_pat.host, rq->m, 0, NULL, 0, &rp->pre_ms, &rp->run_ms);
    if (c < -1) {
        fprintf(stderr, "smartd: search failed: %s\n", sb_last_error());
        rp->status = EIO;
        return;
    }
    rp->count = c;
    g_searches++;
}

/* make sure the resident copy of the text segment (tkey, n) is current; returns 0 or an errno */
static int ensure_text(int tkey, int n, uint64_t gen, double *upload_ms)
{
    const unsigned char *before = g_text.host;
    const unsigned char *t = attach(tkey, (size_t)n, &g_text.shmid, &g_text.host, &g_text.tkey);
    if (!t) {
        if (g_text.dev) sb_text_free(g_text.dev); /* the resident copy belonged to the segment just detached */
        g_text.dev = NULL;
        return errno ? errno : ENOENT;
    }
    const double t0 = now_ms();
    /* a sender-provided generation number identifies the content; otherwise every byte of the text is hashed
     * (up to 64 MiB), and a larger text is uploaded again for every search: never a sampled hash, which could
     * serve a stale resident copy of a text that was rewritten in place between the samples */
    const int trust_hash = gen || (size_t)n <= SB_HASH_LIMIT;
    const uint64_t h = gen ? 0 : (trust_hash ? text_hash(t, (size_t)n) : 0);
    if (!g_text.dev || t != before || g_text.n != (size_t)n || g_text.gen != gen || !trust_hash || g_text.hash != h) {
        if (g_text.dev) sb_text_free(g_text.dev);
        g_text.dev = sb_text_upload(t, (size_t)n, NULL, 1);
        g_text.n = (size_t)n;
        g_text.hash = h;
        g_text.gen = gen;
        g_uploads++;
        if (!g_text.dev) {
            fprintf(stderr, "smartd: text upload failed: %s\n", sb_last_error());
            return EIO;
        }
        *upload_ms = now_ms() - t0;
    }
    return 0;
}

/* SEARCH_MULTI: the whole pattern set of one length in one pass (sb_search_multi64) */
static void serve_multi(const sb_request *rq, sb_reply *rp)
{
    rp->status = 0;
    if (rq->m < 1 || rq->m > SB_XSIZE || rq->n < 0 || rq->npat < 1) {
        rp->status = EINVAL;
        return;
    }
    int pid = shmget((key_t)rq->pkey, (size_t)rq->npat * (size_t)rq->m, 0666);
    int cid = shmget((key_t)rq->ckey, (size_t)rq->npat * sizeof(uint64_t), 0666);
    const unsigned char *pats = pid >= 0 ? shmat(pid, NULL, SHM_RDONLY) : (void *)-1;
    unsigned long long *counts = cid >= 0 ? shmat(cid, NULL, 0) : (void *)-1;
    if (pats == (void *)-1 || counts == (void *)-1) {
        rp->status = errno ? errno : ENOENT;
    } else if ((rp->status = ensure_text(rq->tkey, rq->n, rq->text_gen, &rp->upload_ms)) == 0) {
        char name[sizeof rq->algo_name + 1];
        memcpy(name, rq->algo_name, sizeof rq->algo_name);
        name[sizeof rq->algo_name] = 0;
        const int algo = rq->algo >= 0 ? rq->algo : sb_algo_id(name);
        long long c = sb_search_multi64(algo, g_text.dev, pats, rq->npat, rq->m, 0, counts, &rp->pre_ms, &rp->run_ms);
        if (c < -1) {
            fprintf(stderr, "smartd: multi search failed: %s\n", sb_last_error());
            rp->status = EIO;
        } else {
            rp->count = c;
            g_searches += (uint64_t)rq->npat;
        }
    }
    if (pats != (void *)-1) shmdt(pats);
    if (counts != (void *)-1) shmdt(counts);
}

/* SEARCH_SET: npat independent searches of one length, submitted back to back and collected afterwards */
static void serve_set(const sb_request *rq, sb_reply *rp)
{
    rp->status = 0;
    if (rq->m < 1 || rq->m > SB_XSIZE || rq->n < 0 || rq->npat < 1) {
        rp->status = EINVAL;
        return;
    }
    const size_t np = (size_t)rq->npat;
    int pid = shmget((key_t)rq->pkey, np * (size_t)rq->m, 0666);
    int cid = shmget((key_t)rq->ckey, np * 24, 0666);
    const unsigned char *pats = pid >= 0 ? shmat(pid, NULL, SHM_RDONLY) : (void *)-1;
    unsigned char *res = cid >= 0 ? shmat(cid, NULL, 0) : (void *)-1;
    if (pats == (void *)-1 || res == (void *)-1) {
        rp->status = errno ? errno : ENOENT;
    } else if ((rp->status = ensure_text(rq->tkey, rq->n, rq->text_gen, &rp->upload_ms)) == 0) {
        char name[sizeof rq->algo_name + 1];
        memcpy(name, rq->algo_name, sizeof rq->algo_name);
        name[sizeof rq->algo_name] = 0;
        const int algo = rq->algo >= 0 ? rq->algo : sb_algo_id(name);
        unsigned long long *counts = (unsigned long long *)res;
        double *run = (double *)(res + np * 8), *pre = (double *)(res + np * 16);
        enum { WINDOW = 128 };  /* tickets in flight (the library allows 256) */
        long long tickets[WINDOW];
        long long total = 0;
        size_t sent = 0, done = 0;
        const double t0 = now_ms();
        while (done < np && rp->status == 0) {
            while (sent < np && sent - done < WINDOW) {
                tickets[sent % WINDOW] = sb_search64_submit(algo, g_text.dev, pats + sent * (size_t)rq->m, rq->m, 0);
                if (tickets[sent % WINDOW] < 0) {
                    rp->status = EIO;
                    break;
                }
                sent++;
            }
            if (done < sent) {
                const long long c = sb_search_collect(tickets[done % WINDOW], &pre[done], &run[done]);
                if (c < -1) rp->status = EIO;
                counts[done] = (unsigned long long)c; /* -1 (not applicable) is stored as ~0 */
                if (c > 0) total += c;
                if (c == -1) total = -1;
                done++;
            }
        }
        while (done < sent) { /* drain after an error */
            double a, b;
            sb_search_collect(tickets[done % WINDOW], &a, &b);
            done++;
        }
        if (rp->status) fprintf(stderr, "smartd: pipelined search failed: %s\n", sb_last_error());
        rp->count = total;
        rp->run_ms = now_ms() - t0; /* wall time of the whole set */
        g_searches += (uint64_t)np;
    }
    if (pats != (void *)-1) shmdt(pats);
    if (res != (void *)-1) shmdt(res);
}

/* A live worker already listening on `path` keeps its socket (and its CUDA context and resident text): a second
 * worker must not take it over.  Only a dead socket file is removed.  umask 077 makes the socket private from
 * the moment bind() creates it. */
static int listen_on(const char *path)
{
    struct sockaddr_un sa;
    memset(&sa, 0, sizeof sa);
    sa.sun_family = AF_UNIX;
    snprintf(sa.sun_path, sizeof sa.sun_path, "%s", path);
    int probe = socket(AF_UNIX, SOCK_STREAM, 0);
    if (probe >= 0) {
        const int alive = connect(probe, (struct sockaddr *)&sa, sizeof sa) == 0;
        close(probe);
        if (alive) {
            errno = EADDRINUSE;
            return -2;
        }
    }
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    unlink(path); /* nobody answered: a stale file of a worker that died */
    const mode_t old = umask(077);
    const int rc = bind(fd, (struct sockaddr *)&sa, sizeof sa);
    umask(old);
    if (rc != 0 || listen(fd, 64) != 0) {
        close(fd);
        return -1;
    }
    return fd;
}

static int simple_request(int kind)
{
    int fd = sb_worker_connect();
    if (fd < 0) {
        fprintf(stderr, "smartd: no worker is listening\n");
        return 1;
    }
    sb_request rq = {SB_PROTO_MAGIC, kind, 0, 0, 0, 0, 0, 0, 0, 0, "", 0};
    sb_reply rp;
    int rc = sb_worker_call(fd, &rq, &rp);
    close(fd);
    if (rc != 0) return kind == SB_REQ_QUIT ? 0 : 1;
    if (kind == SB_REQ_STATS)
        printf("searches %llu text_uploads %llu\n", (unsigned long long)rp.searches, (unsigned long long)rp.text_uploads);
    return 0;
}

int main(int argc, char **argv)
{
    int fg = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-sock") && i + 1 < argc) setenv("SMART_B200_SOCK", argv[++i], 1);
        else if (!strcmp(argv[i], "-fg")) fg = 1;
        else if (!strcmp(argv[i], "-stop")) return simple_request(SB_REQ_QUIT);
        else if (!strcmp(argv[i], "-stats")) return simple_request(SB_REQ_STATS);
        else {
            fprintf(stderr, "usage: smartd [-sock path] [-fg] | -stop | -stats\n");
            return 2;
        }
    }
    char path[108];
    sb_socket_path(path, sizeof path);
    /* Daemonise BEFORE the first CUDA call: a CUDA context does not survive fork().  The child reports
     * "ok" or an error line to the parent through a pipe once it is listening. */
    int report[2] = {-1, -1};
    if (!fg) {
        if (pipe(report) != 0) return 5;
        pid_t pid = fork();
        if (pid < 0) return 5;
        if (pid > 0) {
            char msg[256];
            close(report[1]);
            ssize_t k = read(report[0], msg, sizeof msg - 1);
            msg[k > 0 ? k : 0] = 0;
            if (k >= 2 && !strncmp(msg, "ok", 2)) {
                printf("ready %s pid %d\n", path, (int)pid);
                return 0;
            }
            fprintf(stderr, "smartd: %s\n", k > 0 ? msg : "worker died during start-up");
            return (k > 0 && strstr(msg, "already running")) ? 7 : 3; /* 7: a live worker owns the socket (shared) */
        }
        close(report[0]);
        setsid();
        /* detach from the caller's stdio so that a capturing parent is not kept waiting */
        int nul = open("/dev/null", O_RDWR);
        if (nul >= 0) {
            dup2(nul, 0);
            dup2(nul, 1);
            if (!getenv("SMART_B200_KEEP_STDERR")) dup2(nul, 2);
            if (nul > 2) close(nul);
        }
    }
#define FAIL_START(code, ...)                                  \
    do {                                                       \
        char m_[256];                                          \
        snprintf(m_, sizeof m_, __VA_ARGS__);                  \
        if (report[1] >= 0) {                                  \
            if (write(report[1], m_, strlen(m_)) < 0) {}       \
            close(report[1]);                                  \
        } else                                                 \
            fprintf(stderr, "smartd: %s\n", m_);               \
        return code;                                           \
    } while (0)
    if (sb_device_count() < 1) FAIL_START(3, "no CUDA device: %s", sb_last_error());
    /* create the CUDA context now, not inside the first timed search */
    {
        unsigned char dummy[64] = {0};
        sb_text *w = sb_text_upload(dummy, sizeof dummy, NULL, 1);
        if (!w) FAIL_START(6, "CUDA initialisation failed: %s", sb_last_error());
        sb_text_free(w);
    }
    int ls = listen_on(path);
    if (ls == -2) FAIL_START(7, "a worker is already running on %s", path);
    if (ls < 0) FAIL_START(4, "cannot bind %s: %s", path, strerror(errno));
    if (report[1] >= 0) {
        if (write(report[1], "ok", 2) < 0) {}
        close(report[1]);
    } else {
        printf("ready %s pid %d\n", path, (int)getpid());
        fflush(stdout);
    }
    signal(SIGPIPE, SIG_IGN);
    for (;;) {
        int fd = accept(ls, NULL, NULL);
        if (fd < 0) {
            if (errno == EINTR) continue;
            break;
        }
        sb_request rq;
        sb_reply rp;
        memset(&rp, 0, sizeof rp);
        rp.magic = SB_PROTO_MAGIC;
        size_t got = 0;
        while (got < sizeof rq) {
            ssize_t k = read(fd, (char *)&rq + got, sizeof rq - got);
            if (k <= 0) break;
            got += (size_t)k;
        }
        int quit = 0;
        if (got == sizeof rq && rq.magic == SB_PROTO_MAGIC) {
            if (rq.kind == SB_REQ_SEARCH) serve_search(&rq, &rp);
            else if (rq.kind == SB_REQ_SEARCH_MULTI) serve_multi(&rq, &rp);
            else if (rq.kind == SB_REQ_SEARCH_SET) serve_set(&rq, &rp);
            else if (rq.kind == SB_REQ_QUIT) quit = 1;
            rp.searches = g_searches;
            rp.text_uploads = g_uploads;
            if (write(fd, &rp, sizeof rp) < 0) { /* client went away */ }
        }
        close(fd);
        if (quit) break;
    }
    if (g_text.dev) sb_text_free(g_text.dev);
    close(ls);
    unlink(path);
    return 0;
}
