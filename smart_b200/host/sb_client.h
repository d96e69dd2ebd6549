#ifndef SB_CLIENT_H
#define SB_CLIENT_H
#include "sb_proto.h"
/* One search on behalf of a per-algorithm executable.  pkey/tkey are the SysV keys of the pattern and
 * text segments in shared mode (main.h:47-58 of the reference) or -1 in standalone mode.  Returns the
 * count, or -1 ("not applicable" or failure: the driver prints [--], smart.c:377). */
int sb_client_search(int algo, unsigned char *x, int m, unsigned char *y, int n, int pkey, int tkey,
                     double *pre_ms, double *run_ms);
int sb_worker_connect(void);
int sb_worker_call(int fd, const sb_request *rq, sb_reply *rp);
#endif
