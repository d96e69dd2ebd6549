/* Client side of the plugin: find the resident worker, else run the search in this process. */
#include "sb_client.h"

#include "sb_proto.h"
#include "smart_b200.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

const char *sb_socket_path(char *buf, unsigned long len)
{
    const char *e = getenv("SMART_B200_SOCK");
    if (e && *e)
        snprintf(buf, len, "%s", e);
    else
        snprintf(buf, len, "/tmp/smart_b200.%u.sock", (unsigned)getuid());
    return buf;
}

int sb_worker_connect(void)
{
    char path[108];
    struct sockaddr_un sa;
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    memset(&sa, 0, sizeof sa);
    sa.sun_family = AF_UNIX;
    sb_socket_path(path, sizeof path);
    snprintf(sa.sun_path, sizeof sa.sun_path, "%s", path);
    if (connect(fd, (struct sockaddr *)&sa, sizeof sa) != 0) {
        close(fd);
        return -1;
    }
    return fd;
}

static int xfer(int fd, void *buf, size_t len, int writing)
{
    char *p = buf;
    while (len) {
        ssize_t k = writing ? write(fd, p, len) : read(fd, p, len);
        if (k < 0 && errno == EINTR) continue;
        if (k <= 0) return -1;
        p += k;
        len -= (size_t)k;
    }
    return 0;
}

int sb_worker_call(int fd, const sb_request *rq, sb_reply *rp)
{
    if (xfer(fd, (void *)rq, sizeof *rq, 1) != 0) return -1;
    if (xfer(fd, rp, sizeof *rp, 0) != 0) return -1;
    return rp->magic == SB_PROTO_MAGIC ? 0 : -1;
}

int sb_client_search(int algo, unsigned char *x, int m, unsigned char *y, int n, int pkey, int tkey,
                     double *pre_ms, double *run_ms)
{
    /* shared mode with a worker listening: hand over the segment keys, the text stays where it is */
    if (pkey >= 0 && tkey >= 0 && !getenv("SMART_B200_NO_WORKER")) {
        int fd = sb_worker_connect();
        if (fd >= 0) {
            sb_request rq = {SB_PROTO_MAGIC, SB_REQ_SEARCH, algo, pkey, m, tkey, n, 0};
            sb_reply rp;
            int rc = sb_worker_call(fd, &rq, &rp);
            close(fd);
            if (rc == 0 && rp.status == 0) {
                if (pre_ms) *pre_ms = rp.pre_ms + rp.upload_ms;
                if (run_ms) *run_ms = rp.run_ms;
                if (rp.count > 2147483647LL) return 2147483647;
                return (int)rp.count;
            }
            fprintf(stderr, "smart_b200: worker call failed (status %d), searching in-process\n", rc ? -1 : rp.status);
        }
    }
    /* no worker: this process owns the CUDA context for one search (pays context creation + H2D) */
    return sb_plugin_search(algo, x, m, y, n, pre_ms, run_ms);
}
