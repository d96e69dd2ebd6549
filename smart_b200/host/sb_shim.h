/*
 * Plugin shim: the `main()` that is linked into every per-algorithm executable, written from scratch
 * to the protocol of the reference's src/algos/include/main.h:42-136 so that the reference's `smart`
 * and `test` binaries can drive these executables and our driver can drive the reference's:
 *
 *   shared mode      <algo> shared <pkey> <m> <tkey> <n> <rkey> <ekey> <prekey>
 *                    attaches the pattern / text / run-time / pre-time segments (shmget(key,size,0666) +
 *                    shmat), calls search(), stores the count in the 4-byte result segment, exit 0;
 *                    any failure: message on stderr/stdout and exit 1 (the driver maps that to -1,
 *                    smart.c:143-145)
 *   standalone mode  <algo> <P> <m> <T> <n>   prints "found %d occurrences"
 *
 * Timing side channel: search() must store milliseconds in *pre_time and *run_time (main.h:28-37);
 * the shim itself does not time the call, like the reference (main.h:101-106).
 */
#ifndef SB_SHIM_H
#define SB_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ipc.h>
#include <sys/shm.h>
#include <sys/types.h>

double *run_time, *pre_time; /* same names as main.h:34-35: algorithm code writes through them */
int sb_shim_pkey = -1, sb_shim_tkey = -1;

int search(unsigned char *p, int m, unsigned char *t, int n);

static void *sb_attach(const char *what, int key, size_t size)
{
    int id = shmget((key_t)key, size, 0666);
    if (id < 0) {
        perror("shmget");
        fprintf(stderr, "cannot locate the %s segment (key %d)\n", what, key);
        return NULL;
    }
    void *p = shmat(id, NULL, 0);
    if (p == (void *)-1) {
        perror("shmat");
        return NULL;
    }
    return p;
}

int main(int argc, char *argv[])
{
    static double local_run, local_pre;
    if (argc >= 2 && strcmp(argv[1], "shared") == 0) {
        if (argc < 9) {
            printf("error in input parameter\nfive parameters needed when used with shared memory\n");
            return 1;
        }
        const int pkey = atoi(argv[2]), m = atoi(argv[3]), tkey = atoi(argv[4]), n = atoi(argv[5]);
        const int rkey = atoi(argv[6]), ekey = atoi(argv[7]), prekey = atoi(argv[8]);
        unsigned char *p = sb_attach("pattern", pkey, (size_t)m);
        unsigned char *t = p ? sb_attach("text", tkey, (size_t)n) : NULL;
        run_time = t ? sb_attach("running time", ekey, 8) : NULL;
        pre_time = run_time ? sb_attach("preprocessing time", prekey, 8) : NULL;
        if (!pre_time) return 1;
        sb_shim_pkey = pkey;
        sb_shim_tkey = tkey;
        const int count = search(p, m, t, n);
        int *result = sb_attach("result", rkey, 4);
        if (!result) return 1;
        *result = count;
        return 0;
    }
    if (argc < 5) {
        printf("error in input parameter\nfour parameters needed in standard mode\n");
        return 1;
    }
    run_time = &local_run;
    pre_time = &local_pre;
    const int m = atoi(argv[2]), n = atoi(argv[4]);
    if (m < 0 || n < 0 || (size_t)m > strlen(argv[1]) || (size_t)n > strlen(argv[3])) {
        printf("error in input parameter\nm and n must not exceed the argument lengths\n");
        return 1;
    }
    const int occ = search((unsigned char *)argv[1], m, (unsigned char *)argv[3], n);
    printf("found %d occurrences\n", occ);
    return 0;
}
#endif
