/*
 * textgen — regenerates SMART's random corpus buffers (data/rand<sigma>/rand<sigma>.txt + index.txt).
 *
 * Convention of the reference's src/textgen.c:34-54, which the parity tests rely on: eight files of
 * 5 000 000 bytes, byte = rand() % sigma for sigma = 2,4,8,16,32,64,128 and 250, RAW byte values (not
 * ASCII), one continuous glibc rand() stream that is never seeded (so the files are identical on every
 * glibc machine; md5 of rand2.txt = a8e4cecf43689f3964b4fdfe191aa01b).
 *
 *   usage: textgen [-data DIR] [-only SIGMA] [-bytes N]
 *   -only keeps the stream position right: the skipped files' rand() calls are still consumed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

int main(int argc, char **argv)
{
    static const int sigmas[] = {2, 4, 8, 16, 32, 64, 128, 250};
    const char *data = "data";
    long nbytes = 5000000;
    int only = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-data") && i + 1 < argc) data = argv[++i];
        else if (!strcmp(argv[i], "-only") && i + 1 < argc) only = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-bytes") && i + 1 < argc) nbytes = atol(argv[++i]);
        else {
            fprintf(stderr, "usage: textgen [-data DIR] [-only SIGMA] [-bytes N<=5000000]\n");
            return 2;
        }
    }
    if (nbytes < 1 || nbytes > 5000000) nbytes = 5000000;
    mkdir(data, 0775);
    unsigned char *buf = malloc(5000000);
    if (!buf) return 1;
    for (size_t f = 0; f < sizeof sigmas / sizeof sigmas[0]; f++) {
        const int sigma = sigmas[f];
        for (long i = 0; i < 5000000; i++) buf[i] = (unsigned char)(rand() % sigma); /* always the full stream */
        if (only && only != sigma) continue;
        char dir[512], path[600];
        snprintf(dir, sizeof dir, "%s/rand%d", data, sigma);
        mkdir(dir, 0775);
        snprintf(path, sizeof path, "%s/rand%d.txt", dir, sigma);
        printf("\tGenerating random text over %.3d chars ...........", sigma);
        fflush(stdout);
        FILE *out = fopen(path, "wb");
        if (!out || fwrite(buf, 1, (size_t)nbytes, out) != (size_t)nbytes) {
            printf("[FAILED]\n");
            return 1;
        }
        fclose(out);
        snprintf(path, sizeof path, "%s/index.txt", dir);
        out = fopen(path, "w");
        if (out) {
            fprintf(out, "#rand%d.txt#\n", sigma);
            fclose(out);
        }
        printf("..[OK]\n");
    }
    free(buf);
    return 0;
}
