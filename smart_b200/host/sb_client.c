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

int sb_client_search(int algo, unsigned char *x, int m, unsigned char *y, int n, int pkey, int tkey,
                     double *pre_ms, double *run_ms)
{
    /* shared mode with a worker listening: hand over the segment keys, the text stays where it is */
    if (pkey >= 0 && tkey >= 0 && !getenv("SMART_B200_NO_WORKER")) {
        int fd = sb_worker_connect();
        if (fd >= 0) {
            sb_request rq = {SB_PROTO_MAGIC, SB_REQ_SEARCH, algo, pkey, m, tkey, n, 0, 0, 0, "", 0};
            const char *gen = getenv("SMART_B200_TEXT_GEN");
            if (gen) rq.text_gen = strtoull(gen, NULL, 10);
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
