/* Socket plumbing shared by the plugin client, the worker's control commands and the driver's batch mode.
 * No dependency on libsmartb200. */
#include "sb_client.h"
#include "sb_proto.h"

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

