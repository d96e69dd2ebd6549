/*
 * select — edits a SMART algorithm registry (the "#<0|1> #<name>" lines of source/algorithms.h that the
 * reference's drivers read, function.h:62-77) with the command line of the reference's own tool
 * (src/select.c:33-45): -show, -which, NAME (toggle), -all, -none, -add NAME, -h.
 *
 * Differences, all additive:
 *   -file PATH   registry to edit (default source/algorithms.h, as the reference)
 *   -bin DIR     where -add / -init look for executables (default source/bin)
 *   -test PATH   harness run by -add as "<PATH> <name> -nv" (default ./test next to this program's cwd,
 *                select.c:113); -add refuses an algorithm the harness rejects
 *   -init        start from an empty registry and register every executable found in -bin DIR (deselected)
 * A registry that does not exist yet is treated as empty.  The file is rewritten sorted by name, like
 * select.c:183-194.  Pure host C: no GPU, no libsmartb200.
 */
#include <dirent.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <unistd.h>

#define MAX_ENTRIES 1024

typedef struct {
    char name[64];
    int selected;
} entry;

static entry reg[MAX_ENTRIES];
static int nreg;

static void manual(void)
{
    printf("\n\tSMART UTILITY FOR SELECTING STRING MATCHING ALGORITHMS (smart_b200)\n\n"
           "\t-show           shows the list of all algorithms\n"
           "\t-which          shows the list of all selected algorithms\n"
           "\tALGO            selects/deselects the algorithm ALGO (ex. select bf)\n"
           "\t-all            selects all algorithms\n"
           "\t-none           deselects all algorithms\n"
           "\t-add ALGO       add the new algorithm ALGO to the set (its executable must be in the bin directory\n"
           "\t                and pass the test harness)\n"
           "\t-init           rebuild the set from the executables of the bin directory\n"
           "\t-file PATH      registry file (default source/algorithms.h)\n"
           "\t-bin DIR        bin directory (default source/bin)\n"
           "\t-test PATH      test harness used by -add (default ./test)\n"
           "\t-h              gives this help list\n\n");
}

static int find(const char *name)
{
    for (int i = 0; i < nreg; i++)
        if (!strcasecmp(reg[i].name, name)) return i;
    return -1;
}

static int append(const char *name, int selected)
{
    if (nreg >= MAX_ENTRIES || strlen(name) >= sizeof reg[0].name) return -1;
    snprintf(reg[nreg].name, sizeof reg[nreg].name, "%s", name);
    reg[nreg].selected = selected;
    return nreg++;
}

/* "#1 #name " per line; anything else is ignored (function.h:62-77 reads the same two fields) */
static void load(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) return;
    char line[256];
    while (fgets(line, sizeof line, f)) {
        int sel;
        char name[64];
        if (sscanf(line, "#%d #%63s", &sel, name) == 2 && find(name) < 0) append(name, sel != 0);
    }
    fclose(f);
}

static int by_name(const void *a, const void *b) { return strcmp(((const entry *)a)->name, ((const entry *)b)->name); }

static int save(const char *path)
{
    FILE *f = fopen(path, "w");
    if (!f) {
        printf("\tError in writing file %s\n", path);
        return -1;
    }
    qsort(reg, (size_t)nreg, sizeof reg[0], by_name);
    for (int i = 0; i < nreg; i++) fprintf(f, "#%d #%s \n", reg[i].selected, reg[i].name);
    fclose(f);
    return 0;
}

static int is_executable(const char *dir, const char *name)
{
    char path[1024];
    struct stat st;
    snprintf(path, sizeof path, "%s/%s", dir, name);
    return stat(path, &st) == 0 && S_ISREG(st.st_mode) && access(path, X_OK) == 0;
}

int main(int argc, char **argv)
{
    const char *file = "source/algorithms.h", *bindir = "source/bin", *harness = "./test";
    if (argc == 1) {
        printf("\n\tNo parameter given. Use -h for help.\n\n");
        return 0;
    }
    /* options that say WHERE come first, whatever their position on the command line */
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "-file")) file = argv[i + 1];
        else if (!strcmp(argv[i], "-bin")) bindir = argv[i + 1];
        else if (!strcmp(argv[i], "-test")) harness = argv[i + 1];
    }
    load(file);
    int dirty = 0;
    for (int par = 1; par < argc; par++) {
        const char *a = argv[par];
        if (!strcmp(a, "-h")) {
            manual();
            return 0;
        } else if (!strcmp(a, "-file") || !strcmp(a, "-bin") || !strcmp(a, "-test")) {
            par++;
        } else if (!strcmp(a, "-show")) {
            printf("The list of all string matching algorithms\n");
            for (int i = 0; i < nreg; i++) printf("%s\n", reg[i].name);
        } else if (!strcmp(a, "-which")) {
            printf("\n\tThe list of selected algorithms:\n");
            for (int i = 0; i < nreg; i++)
                if (reg[i].selected) printf("\t-%s\n", reg[i].name);
            printf("\n");
        } else if (!strcmp(a, "-all") || !strcmp(a, "-none")) {
            for (int i = 0; i < nreg; i++) reg[i].selected = a[1] == 'a';
            dirty = 1;
        } else if (!strcmp(a, "-init")) {
            DIR *d = opendir(bindir);
            if (!d) {
                printf("\n\n\tSMART error message\n\tcannot read the directory %s\n\n", bindir);
                return 1;
            }
            nreg = 0;
            for (struct dirent *e; (e = readdir(d));)
                if (e->d_name[0] != '.' && is_executable(bindir, e->d_name)) append(e->d_name, 0);
            closedir(d);
            printf("\t%d algorithms registered from %s\n", nreg, bindir);
            dirty = 1;
        } else if (!strcmp(a, "-add")) {
            if (++par >= argc) {
                printf("\n\n\tError in input parameters. Use -h for help.\n\n");
                return 1;
            }
            const char *name = argv[par];
            if (!is_executable(bindir, name)) {
                printf("\n\n\tSMART error message\n\tError in input parameters....program %s/%s does not exist\n\n", bindir, name);
            } else if (find(name) >= 0) {
                printf("\n\n\tSMART error message\n\tError in input parameters....algorithm %s already in the set\n\n", name);
            } else {
                char cmd[2048];
                printf("\n\n\tAdding the algorithm %s to SMART\n\tTesting the algorithm for correctness....", name);
                fflush(stdout);
                /* the reference's harness only knows "<name> -nv" and looks in ./source/bin; ours takes -bin */
                if (strcmp(bindir, "source/bin")) snprintf(cmd, sizeof cmd, "%s %s -nv -bin %s > /dev/null 2>&1", harness, name, bindir);
                else snprintf(cmd, sizeof cmd, "%s %s -nv > /dev/null 2>&1", harness, name);
                if (system(cmd) != 0) {
                    printf("failed!\n\tThe system is unable to add the algorithm %s to SMART.\n"
                           "\tPlease, check for algorithm's correctness.\n\n", name);
                } else {
                    printf("ok\n");
                    append(name, 0);
                    dirty = 1;
                    printf("\tAlgorithm %s added succesfully.\n\n", name);
                }
            }
        } else if (a[0] != '-' && find(a) >= 0) {
            entry *e = &reg[find(a)];
            e->selected = !e->selected;
            printf("\tThe %s algorithm has been %s\n", e->name, e->selected ? "selected" : "deselected");
            dirty = 1;
        } else {
            printf("\tError in input parameters....no parameter %s\n\n", a);
            return 1;
        }
    }
    return dirty ? (save(file) == 0 ? 0 : 1) : 0;
}
