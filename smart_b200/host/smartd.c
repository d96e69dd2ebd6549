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

/* content hash of the text: every byte for texts up to 64 MiB, 16384 evenly spaced 4 KiB blocks above */
static uint64_t text_hash(const unsigned char *t, size_t n)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
    const size_t full = (size_t)64 << 20;
    size_t nblk = 1, blk = n, stride = 0;
    if (n > full) {
        nblk = 16384;
        blk = 4096;
        stride = (n - blk) / (nblk - 1);
    }
    for (size_t b = 0; b < nblk; b++) {
        const unsigned char *p = t + b * stride;
        size_t i = 0;
        uint64_t a = h, c = ~h;
        for (; i + 16 <= blk; i += 16) {
            uint64_t u, v;
            memcpy(&u, p + i, 8);
            memcpy(&v, p + i + 8, 8);
            a = (a ^ u) * 0xFF51AFD7ED558CCDull;
            a = (a << 29) | (a >> 35);
            c = (c ^ v) * 0xC4CEB9FE1A85EC53ull;
            c = (c << 31) | (c >> 33);
        }
        for (; i < blk; i++) a = (a ^ p[i]) * 0x100000001B3ull;
        h = (a ^ (c * 0x9E3779B97F4A7C15ull)) + b;
    }
    return h;
}

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

static const unsigned char *attach(int key, size_t size, int *shmid_io, const unsigned char *old, int oldkey)
{
    int id = shmget((key_t)key, size, 0666);
    if (id < 0) return NULL;
    if (old && key == oldkey && id == *shmid_io) return old; /* same segment still attached */
    if (old) shmdt(old);
    void *p = shmat(id, NULL, SHM_RDONLY);
    if (p == (void *)-1) return NULL;
    *shmid_io = id;
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
    g_pat.host = attach(rq->pkey, (size_t)rq->m, &g_pat.shmid, g_pat.host, g_pat.pkey);
    g_pat.pkey = rq->pkey;
    if (!g_pat.host) {
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
    const unsigned char *t = attach(tkey, (size_t)n, &g_text.shmid, g_text.host, g_text.tkey);
    if (!t) return errno ? errno : ENOENT;
    const double t0 = now_ms();
    /* a sender-provided generation number identifies the content; otherwise hash the text */
    const uint64_t h = gen ? 0 : text_hash(t, (size_t)n);
    if (!g_text.dev || t != g_text.host || g_text.tkey != tkey || g_text.n != (size_t)n || g_text.gen != gen ||
        g_text.hash != h) {
        if (g_text.dev) sb_text_free(g_text.dev);
        g_text.dev = sb_text_upload(t, (size_t)n, NULL, 1);
        g_text.host = t;
        g_text.tkey = tkey;
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

static int listen_on(const char *path)
{
    struct sockaddr_un sa;
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    memset(&sa, 0, sizeof sa);
    sa.sun_family = AF_UNIX;
    snprintf(sa.sun_path, sizeof sa.sun_path, "%s", path);
    unlink(path);
    if (bind(fd, (struct sockaddr *)&sa, sizeof sa) != 0 || listen(fd, 64) != 0) {
        close(fd);
        return -1;
    }
    chmod(path, 0600);
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
            return 3;
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
