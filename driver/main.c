/*
 * driver/main.c -- the reference driver's command line on top of libtilespgemm_b200.so.
 *
 *   ./test_b200 -d <device> -aat <0|1> <matrix.mtx | gen:NAME:ARGS> <tile_size_m> <tile_size_n>
 *
 * Same argument order and same call sequence as the reference's src/main.cu:13-359:
 * load -> value[i] = i % 10 (:111-112) -> (aat) matrix_transposition (:114-142) -> nnzCub (:155-160)
 * -> csr2tile_row_major(A) (:168) -> csr2tile_col_major(B) (:191) -> tilespgemm (:261) -> tile2csr(C)
 * (:327) -> matrix_destroy (:351-352). What changed relative to the reference driver:
 *   - the three hot-path #includes are replaced by include/tilespgemm.h (the whole integration);
 *   - the dense tile bitmaps (:195-232, O(tilem*tilen/32) bytes, only read by unreachable kernels) are
 *     not built: NULL/0 is passed;
 *   - gen:lap2d:N, gen:stencil27:N, gen:blockfem:NODES, gen:rmat:SCALE:EF[:a:b:c] replace SuiteSparse
 *     files (not available offline); .mtx input is sorted and de-duplicated (the library's contract);
 *   - the library's error latch is checked after every call and the driver exits non-zero;
 *   - the cuSPARSE structure check (:344) is replaced by an in-driver serial SPA structure check
 *     (-DTSG_DRIVER_CHECK=1, small inputs only);
 *   - an optional trailing "-slabs <max tile pairs per slab>" runs the product slab by slab through the device-resident
 *     API (tsg_spgemm_slabs): for matrices whose C does not fit the 32-bit sizes of SMatrix (R-MAT scale 20: nnz(C) ~ 1e10,
 *     SURVEY.md fact 10). Totals are 64-bit; every slab is checked through C*ones = A*(B*ones) (exact for the driver's
 *     integer values) and its per-row counts are summed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "tilespgemm.h"

#ifndef TSG_DRIVER_CHECK
#define TSG_DRIVER_CHECK 1
#endif

/* -slabs mode: state shared with the slab callback */
typedef struct {
    const double *expect;   /* A*(B*ones), one entry per row of C */
    double *sums;           /* scratch, rows of the largest slab */
    long long *counts;
    long long nnz, bad_rows, rows_seen;
    int slabs;
} slab_check_t;

static int slab_sink(const tsg_dtile *c, const tsg_stats *st, void *user)
{
    slab_check_t *k = (slab_check_t *)user;
    (void)st;
    if (tsg_tile_rowsums(c, k->sums, k->counts) != TSG_OK) return 1;
    const long long r0 = (long long)c->trow0 * 16;
    for (int i = 0; i < c->m; i++) {
        if (k->sums[i] != k->expect[r0 + i]) k->bad_rows++;
        k->nnz += k->counts[i];
    }
    k->rows_seen += c->m;
    k->slabs++;
    return 0;
}

static double now_ms(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec * 1000.0 + t.tv_usec / 1000.0;
}

static void die_on_error(const char *where)
{
    if (tilespgemm_last_error()) {
        fprintf(stderr, "%s failed: error %d: %s\n", where, tilespgemm_last_error(), tilespgemm_last_error_string());
        exit(2);
    }
}

typedef struct { long long key; double v; } coo_t;
static int cmp_coo(const void *a, const void *b)
{
    long long x = ((const coo_t *)a)->key, y = ((const coo_t *)b)->key;
    return (x > y) - (x < y);
}

/* COO (possibly unsorted, with duplicates) -> sorted duplicate-free CSR. */
static void coo_to_csr(int m, int n, long long cnt, coo_t *e, SMatrix *A)
{
    qsort(e, (size_t)cnt, sizeof(coo_t), cmp_coo);
    long long u = 0;
    for (long long i = 0; i < cnt; i++)
        if (i == 0 || e[i].key != e[u - 1].key) e[u++] = e[i];
    if (u >= 2147483647LL) { fprintf(stderr, "nnz does not fit int\n"); exit(2); }
    A->m = m; A->n = n; A->nnz = (int)u; A->isSymmetric = 0;
    A->rowpointer = (int *)calloc((size_t)m + 1, sizeof(int));
    A->columnindex = (int *)malloc((size_t)(u ? u : 1) * sizeof(int));
    A->value = (double *)malloc((size_t)(u ? u : 1) * sizeof(double));
    for (long long i = 0; i < u; i++) {
        int r = (int)(e[i].key / n);
        A->rowpointer[r + 1]++;
        A->columnindex[i] = (int)(e[i].key % n);
        A->value[i] = e[i].v;
    }
    for (int i = 0; i < m; i++) A->rowpointer[i + 1] += A->rowpointer[i];
}

/* File-order CSR, exactly what the reference's mmio_allinone builds (src/mmio_highlevel.h:593-759): rows in the order
 * the entries appear in the file, symmetric entries mirrored on the spot, nothing sorted, duplicates kept. Selected with
 * TSG_MTX_RAW=1: the library's drop-in csr2tile_* then canonicalise on the device (tsg_csr_canonicalize). */
static void coo_to_csr_file_order(int m, int n, long long cnt, const coo_t *e, SMatrix *A)
{
    A->m = m; A->n = n; A->nnz = (int)cnt; A->isSymmetric = 0;
    A->rowpointer = (int *)calloc((size_t)m + 2, sizeof(int));
    A->columnindex = (int *)malloc((size_t)(cnt ? cnt : 1) * sizeof(int));
    A->value = (double *)malloc((size_t)(cnt ? cnt : 1) * sizeof(double));
    for (long long i = 0; i < cnt; i++) A->rowpointer[e[i].key / n + 2]++;
    for (int i = 0; i < m; i++) A->rowpointer[i + 2] += A->rowpointer[i + 1];
    for (long long i = 0; i < cnt; i++) {
        const int r = (int)(e[i].key / n), at = A->rowpointer[r + 1]++;
        A->columnindex[at] = (int)(e[i].key % n);
        A->value[at] = e[i].v;
    }
}

/* Binary cache next to the .mtx file (text parsing of a 57M-entry file takes minutes): the CSR exactly
 * as load_mtx built it. "<file>.tsgcsr" (".tsgraw" for TSG_MTX_RAW=1); ignored when older than the .mtx; TSG_MTX_CACHE=0 disables. */
typedef struct { char magic[8]; int m, n, nnz, sym; } cache_hdr_t;
#include <sys/stat.h>

static int cache_load(const char *cpath, const char *mtx, SMatrix *A)
{
    struct stat sc, sm;
    if (stat(cpath, &sc) || stat(mtx, &sm) || sc.st_mtime < sm.st_mtime) return -1;
    FILE *f = fopen(cpath, "rb");
    if (!f) return -1;
    cache_hdr_t h;
    int ok = fread(&h, sizeof h, 1, f) == 1 && !memcmp(h.magic, "TSGCSR1", 8) && h.m >= 0 && h.n >= 0 && h.nnz >= 0;
    if (ok) {
        A->m = h.m; A->n = h.n; A->nnz = h.nnz; A->isSymmetric = h.sym;
        A->rowpointer = (int *)malloc(((size_t)h.m + 1) * sizeof(int));
        A->columnindex = (int *)malloc((size_t)(h.nnz ? h.nnz : 1) * sizeof(int));
        A->value = (double *)malloc((size_t)(h.nnz ? h.nnz : 1) * sizeof(double));
        ok = fread(A->rowpointer, sizeof(int), (size_t)h.m + 1, f) == (size_t)h.m + 1 &&
             fread(A->columnindex, sizeof(int), (size_t)h.nnz, f) == (size_t)h.nnz &&
             fread(A->value, sizeof(double), (size_t)h.nnz, f) == (size_t)h.nnz && A->rowpointer[h.m] == h.nnz;
        if (!ok) { free(A->rowpointer); free(A->columnindex); free(A->value); }
    }
    fclose(f);
    return ok ? 0 : -1;
}

static void cache_store(const char *cpath, const SMatrix *A)
{
    FILE *f = fopen(cpath, "wb");
    if (!f) return;  /* read-only directory: no cache */
    cache_hdr_t h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "TSGCSR1", 8);
    h.m = A->m; h.n = A->n; h.nnz = A->nnz; h.sym = A->isSymmetric;
    int ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(A->rowpointer, sizeof(int), (size_t)A->m + 1, f) == (size_t)A->m + 1 &&
             fwrite(A->columnindex, sizeof(int), (size_t)A->nnz, f) == (size_t)A->nnz &&
             fwrite(A->value, sizeof(double), (size_t)A->nnz, f) == (size_t)A->nnz;
    fclose(f);
    if (!ok) remove(cpath);
}

/* MatrixMarket coordinate files: real, integer, pattern (value 1) and complex (real part kept, like mmio_allinone's
 * "%lg %lg" read, src/mmio_highlevel.h:662-666); general, symmetric and hermitian (mirrored, :636-640, :701-723). */
static int load_mtx(const char *path, SMatrix *A)
{
    const int raw = getenv("TSG_MTX_RAW") && atoi(getenv("TSG_MTX_RAW"));
    const int use_cache = !(getenv("TSG_MTX_CACHE") && !atoi(getenv("TSG_MTX_CACHE")));
    char cpath[1100];
    snprintf(cpath, sizeof cpath, "%s.%s", path, raw ? "tsgraw" : "tsgcsr");
    if (use_cache && cache_load(cpath, path, A) == 0) { printf("(binary cache %s)\n", cpath); return 0; }
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    if (!fgets(line, sizeof line, f)) { fclose(f); return -1; }
    if (strncmp(line, "%%MatrixMarket", 14) || !strstr(line, "coordinate")) { fclose(f); fprintf(stderr, "%s: not a MatrixMarket coordinate file\n", path); return -1; }
    int pattern = strstr(line, "pattern") != NULL;
    int symmetric = strstr(line, "symmetric") != NULL || strstr(line, "hermitian") != NULL;
    do { if (!fgets(line, sizeof line, f)) { fclose(f); return -1; } } while (line[0] == '%');
    int m, n;
    long long nz;
    if (sscanf(line, "%d %d %lld", &m, &n, &nz) != 3 || m < 0 || n < 0 || nz < 0) { fclose(f); return -1; }
    coo_t *e = (coo_t *)malloc(sizeof(coo_t) * (size_t)(2 * nz + 1));
    long long cnt = 0;
    for (long long i = 0; i < nz; i++) {
        int r, c;
        double v = 1.0;
        if (!fgets(line, sizeof line, f)) break;
        /* "%lf" reads a real, an integer, or the real part of a complex entry alike */
        if ((pattern ? sscanf(line, "%d %d", &r, &c) : sscanf(line, "%d %d %lf", &r, &c, &v)) < 2) continue;
        if (r < 1 || r > m || c < 1 || c > n) { fprintf(stderr, "%s: entry (%d,%d) outside %dx%d\n", path, r, c, m, n); free(e); fclose(f); return -1; }
        e[cnt].key = (long long)(r - 1) * n + (c - 1); e[cnt++].v = v;
        if (symmetric && r != c) { e[cnt].key = (long long)(c - 1) * n + (r - 1); e[cnt++].v = v; }
    }
    fclose(f);
    if (raw) coo_to_csr_file_order(m, n, cnt, e, A); else coo_to_csr(m, n, cnt, e, A);
    A->isSymmetric = symmetric;  /* like mmio_allinone (src/mmio_highlevel.h) */
    free(e);
    if (use_cache) cache_store(cpath, A);
    return 0;
}

static unsigned long long rng_state = 88172645463325252ULL;
static double rng_u(void)
{
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

/* gen:lap2d:N | gen:stencil27:N | gen:blockfem:NODES | gen:rmat:SCALE:EF[:a:b:c] */
static int generate(const char *spec, SMatrix *A)
{
    char name[64];
    double p[6] = {0, 0, 0.57, 0.19, 0.19, 0};
    int np = 0;
    const char *s = spec + 4;
    const char *colon = strchr(s, ':');
    size_t nl = colon ? (size_t)(colon - s) : strlen(s);
    if (nl >= sizeof name) return -1;
    memcpy(name, s, nl); name[nl] = 0;
    while (colon && np < 6) { p[np++] = atof(colon + 1); colon = strchr(colon + 1, ':'); }
    if (!strcmp(name, "lap2d") || !strcmp(name, "stencil27")) {
        int three = name[0] == 's', N = (int)p[0];
        long long n = three ? (long long)N * N * N : (long long)N * N;
        long long cap = n * (three ? 27 : 5);
        coo_t *e = (coo_t *)malloc(sizeof(coo_t) * (size_t)cap);
        long long cnt = 0;
        for (long long i = 0; i < n; i++) {
            int x = (int)(i % N), y = (int)((i / N) % N), z = (int)(i / ((long long)N * N));
            for (int dz = three ? -1 : 0; dz <= (three ? 1 : 0); dz++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!three && dx != 0 && dy != 0) continue; /* 5-point */
                        if (x + dx < 0 || x + dx >= N || y + dy < 0 || y + dy >= N || z + dz < 0 || z + dz >= (three ? N : 1)) continue;
                        e[cnt].key = i * n + (i + dx + (long long)dy * N + (long long)dz * N * N); e[cnt++].v = 1.0;
                    }
        }
        coo_to_csr((int)n, (int)n, cnt, e, A);
        free(e);
        return 0;
    }
    if (!strcmp(name, "blockfem")) {
        int nodes = (int)p[0], dof = 6;
        long long n = (long long)nodes * dof, cnt = 0;
        coo_t *e = (coo_t *)malloc(sizeof(coo_t) * (size_t)(n * 3 * dof));
        for (long long i = 0; i < n; i++) {
            long long b = i / dof;
            for (long long bj = (b > 0 ? b - 1 : 0); bj <= (b < nodes - 1 ? b + 1 : nodes - 1); bj++)
                for (int c = 0; c < dof; c++) { e[cnt].key = i * n + bj * dof + c; e[cnt++].v = 1.0; }
        }
        coo_to_csr((int)n, (int)n, cnt, e, A);
        free(e);
        return 0;
    }
    if (!strcmp(name, "rmat")) {
        int scale = (int)p[0], ef = np > 1 ? (int)p[1] : 16;
        double a = np > 2 ? p[2] : 0.57, b = np > 3 ? p[3] : 0.19, c = np > 4 ? p[4] : 0.19;
        long long n = 1LL << scale, ne = n * ef;
        coo_t *e = (coo_t *)malloc(sizeof(coo_t) * (size_t)ne);
        for (long long i = 0; i < ne; i++) {
            long long r = 0, col = 0;
            for (int bit = 0; bit < scale; bit++) {
                double u = rng_u();
                if (u >= a + b) r |= 1LL << bit;
                if ((u >= a && u < a + b) || u >= a + b + c) col |= 1LL << bit;
            }
            e[i].key = r * n + col; e[i].v = 1.0;
        }
        coo_to_csr((int)n, (int)n, ne, e, A);
        free(e);
        return 0;
    }
    return -1;
}

#if TSG_DRIVER_CHECK
/* serial SPA, structure + values (the reference checks structure against cuSPARSE, main.cu:344) */
static int check_small(const SMatrix *A, const SMatrix *B, const SMatrix *C)
{
    int nB = B->n, bad = 0;
    char *flag = (char *)calloc((size_t)nB, 1);
    double *acc = (double *)calloc((size_t)nB, sizeof(double));
    for (int i = 0; i < A->m && !bad; i++) {
        for (int ja = A->rowpointer[i]; ja < A->rowpointer[i + 1]; ja++) {
            int k = A->columnindex[ja];
            for (int jb = B->rowpointer[k]; jb < B->rowpointer[k + 1]; jb++) {
                flag[B->columnindex[jb]] = 1;
                acc[B->columnindex[jb]] += A->value[ja] * B->value[jb];
            }
        }
        int pos = C->rowpointer[i];
        for (int c = 0; c < nB; c++)
            if (flag[c]) {
                if (pos >= C->rowpointer[i + 1] || C->columnindex[pos] != c || C->value[pos] != acc[c]) { bad = 1; break; }
                pos++; flag[c] = 0; acc[c] = 0.0;
            }
        if (pos != C->rowpointer[i + 1]) bad = 1;
    }
    free(flag); free(acc);
    return bad;
}
#endif

int main(int argc, char **argv)
{
    if (argc < 7) {
        printf("Run the code by './test_b200 -d 0 -aat 0 matrix.mtx|gen:NAME:ARGS tile_size_m tile_size_n'.\n");
        return 0;
    }
    if (strcmp(argv[1], "-d") != 0 || strcmp(argv[3], "-aat") != 0) return 0;
    int device_id = atoi(argv[2]), aat = atoi(argv[4]);
    char *filename = argv[5];
    /* -aat 2 A B tile_m tile_n: the general product C = A*B of two files, the CLI shape of the reference's cuSPARSE harness
     * (src/external/cusparse/main.cu:83-126) */
    char *filenameB = NULL;
    if (aat == 2) {
        if (argc < 9) { printf("-aat 2 needs two matrices: ./test_b200 -d 0 -aat 2 A.mtx B.mtx tile_size_m tile_size_n\n"); return 0; }
        filenameB = argv[6];
        argv++; argc--;
    }
    int tile_size_m = atoi(argv[6]), tile_size_n = argc > 7 ? atoi(argv[7]) : tile_size_m;
    printf("device_id = %i\n", device_id);
    /* the reference calls cudaSetDevice(device_id) (main.cu:49); the library owns its CUDA runtime, so select there */
    if (!getenv("TSG_DRIVER_PARSE_ONLY") && tsg_init(device_id) != TSG_OK) {
        fprintf(stderr, "tsg_init(%d) failed: %s\n", device_id, tilespgemm_last_error_string());
        return 2;
    }

    SMatrix *matrixA = (SMatrix *)calloc(1, sizeof(SMatrix));
    SMatrix *matrixB = (SMatrix *)calloc(1, sizeof(SMatrix));
    SMatrix *matrixC = (SMatrix *)calloc(1, sizeof(SMatrix));
    printf("MAT: -------------- %s --------------\n", filename);
    double t0 = now_ms();
    int rc = strncmp(filename, "gen:", 4) == 0 ? generate(filename, matrixA) : load_mtx(filename, matrixA);
    if (rc) { fprintf(stderr, "cannot load %s\n", filename); return 2; }
    printf("input matrix A: ( %i, %i ) nnz = %i\n loadfile time    = %4.5f sec\n", matrixA->m, matrixA->n, matrixA->nnz,
           (now_ms() - t0) / 1000.0);
    if (!aat && matrixA->m != matrixA->n) { printf("matrix squaring must have rowA == colA. Exit.\n"); return 0; }
    if (getenv("TSG_DRIVER_PARSE_ONLY")) {  /* loader check without a device (tests/test_driver_loader.py) */
        long long s0 = 0; double s1 = 0;
        int sorted = 1;
        for (int i = 0; i < matrixA->m; i++)
            for (int p = matrixA->rowpointer[i]; p < matrixA->rowpointer[i + 1]; p++) {
                s0 += (long long)(i + 1) * (matrixA->columnindex[p] + 1); s1 += matrixA->value[p];
                if (p > matrixA->rowpointer[i] && matrixA->columnindex[p - 1] >= matrixA->columnindex[p]) sorted = 0;
            }
        printf("parsed: m=%d n=%d nnz=%d symmetric=%d sorted=%d index_checksum=%lld value_sum=%.6f\n", matrixA->m, matrixA->n, matrixA->nnz,
               matrixA->isSymmetric, sorted, s0, s1);
        return 0;
    }
    printf("the tile_size_m = %d\nthe tile_size_n = %d\n", tile_size_m, tile_size_n);
    for (int i = 0; i < matrixA->nnz; i++) matrixA->value[i] = i % 10;       /* main.cu:111-112 */

    if (aat && matrixA->m == matrixA->n && matrixA->isSymmetric) {           /* main.cu:120-124 */
        printf("matrix AAT does not do symmetric matrix. Exit.\n");
        return 0;
    }
    if (aat == 2) {                                                          /* general C = A*B: B from its own file */
        int rcb = strncmp(filenameB, "gen:", 4) == 0 ? generate(filenameB, matrixB) : load_mtx(filenameB, matrixB);
        if (rcb) { fprintf(stderr, "cannot load %s\n", filenameB); return 2; }
        printf("input matrix B: ( %i, %i ) nnz = %i\n", matrixB->m, matrixB->n, matrixB->nnz);
        if (matrixA->n != matrixB->m) { printf("C = A*B needs colA == rowB. Exit.\n"); return 0; }
        for (int i = 0; i < matrixB->nnz; i++) matrixB->value[i] = i % 10;
    } else if (aat) {                                                        /* main.cu:114-142 */
        matrixB->m = matrixA->n; matrixB->n = matrixA->m; matrixB->nnz = matrixA->nnz;
        matrixB->rowpointer = (int *)malloc(((size_t)matrixA->n + 1) * sizeof(int));
        matrixB->columnindex = (int *)malloc((size_t)(matrixA->nnz ? matrixA->nnz : 1) * sizeof(int));
        matrixB->value = (double *)malloc((size_t)(matrixA->nnz ? matrixA->nnz : 1) * sizeof(double));
        matrix_transposition(matrixA->m, matrixA->n, matrixA->nnz, matrixA->rowpointer, matrixA->columnindex, matrixA->value,
                             matrixB->columnindex, matrixB->rowpointer, matrixB->value);
        die_on_error("matrix_transposition");
    } else {                                                                 /* main.cu:145-151: B aliases A */
        matrixB->m = matrixA->m; matrixB->n = matrixA->n; matrixB->nnz = matrixA->nnz;
        matrixB->rowpointer = matrixA->rowpointer; matrixB->columnindex = matrixA->columnindex; matrixB->value = matrixA->value;
    }
    unsigned long long nnzCub = 0;                                           /* main.cu:155-160 */
    for (int i = 0; i < matrixA->nnz; i++) {
        int rowidx = matrixA->columnindex[i];
        nnzCub += (unsigned long long)(matrixB->rowpointer[rowidx + 1] - matrixB->rowpointer[rowidx]);
    }
    printf("SpGEMM nnzCub = %llu\n", nnzCub);

    if (argc > 9 && strcmp(argv[8], "-slabs") == 0) {
        /* slab-wise product through the device-resident API: C never exists whole, totals are 64-bit */
        const long long max_pairs = atoll(argv[9]);
        if (tile_size_m != 16 || tile_size_n != 16) { fprintf(stderr, "-slabs runs the 16 x 16 kernels (tsg_spgemm_slabs); other tile sizes take the whole-matrix path\n"); return 2; }
        tsg_dcsr dA, dB;
        tsg_dtile tA, tB;
        tsg_stats tot;
        memset(&dA, 0, sizeof dA); memset(&dB, 0, sizeof dB); memset(&tA, 0, sizeof tA); memset(&tB, 0, sizeof tB);
        if (tsg_csr_upload(matrixA->m, matrixA->n, matrixA->rowpointer, matrixA->columnindex, matrixA->value, &dA) ||
            (aat ? tsg_csr_upload(matrixB->m, matrixB->n, matrixB->rowpointer, matrixB->columnindex, matrixB->value, &dB) : 0) ||
            tsg_csr2tile(&dA, 0, &tA) || tsg_csr2tile(aat ? &dB : &dA, 1, &tB))
            die_on_error("slab set-up");
        /* expected row sums of C on the host: A * (B * ones) */
        double *bones = (double *)calloc((size_t)matrixB->m + 1, sizeof(double)), *expect = (double *)calloc((size_t)matrixA->m + 1, sizeof(double));
        for (int i = 0; i < matrixB->m; i++)
            for (int p = matrixB->rowpointer[i]; p < matrixB->rowpointer[i + 1]; p++) bones[i] += matrixB->value[p];
        for (int i = 0; i < matrixA->m; i++)
            for (int p = matrixA->rowpointer[i]; p < matrixA->rowpointer[i + 1]; p++) expect[i] += matrixA->value[p] * bones[matrixA->columnindex[p]];
        slab_check_t chk;
        memset(&chk, 0, sizeof chk);
        chk.expect = expect;
        chk.sums = (double *)malloc(((size_t)matrixA->m + 16) * sizeof(double));
        chk.counts = (long long *)malloc(((size_t)matrixA->m + 16) * sizeof(long long));
        int nslabs = 0;
        for (int rep = 0; rep < 2; rep++) { /* first sweep warms the memory pool, second is reported */
            memset(&tot, 0, sizeof tot);
            chk.nnz = chk.bad_rows = chk.rows_seen = 0; chk.slabs = 0;
            if (tsg_spgemm_slabs(&tA, &tB, 0, -1, max_pairs, slab_sink, &chk, &tot, &nslabs)) die_on_error("tsg_spgemm_slabs");
        }
        printf("slabs = %d (<= %lld tile pairs each)\nC tiles listed = %lld\nnnzC = %lld\ntile pairs = %lld\n", nslabs, max_pairs, tot.numblkC,
               tot.nnzC, tot.pairs);
        printf("step1 %.3f ms, step2 %.3f ms, step3 %.3f ms, alloc %.3f ms, compression rate %.3f\n", tot.ms_step1, tot.ms_step2, tot.ms_step3,
               tot.ms_alloc, tot.nnzC ? (double)nnzCub / (double)tot.nnzC : 0.0);
        printf("CUDA  TileSpGEMM runtime is %4.2f ms, gflops = %4.2f\n", tot.ms_total, tot.ms_total > 0 ? 2.0 * (double)nnzCub / (tot.ms_total * 1e6) : 0.0);
        printf("-------------------------------check----------------------------------------\n");
        const int ok = chk.bad_rows == 0 && chk.nnz == tot.nnzC && chk.rows_seen == matrixA->m;
        printf("C*ones == A*(B*ones) on %lld rows, per-row counts sum to nnzC: [%s]\n", chk.rows_seen, ok ? "PASSED" : "NOT PASSED");
        tsg_tile_free(&tA); tsg_tile_free(&tB); tsg_csr_free(&dA); tsg_csr_free(&dB);
        return ok ? 0 : 3;
    }

    t0 = now_ms();
    csr2tile_row_major(matrixA, tile_size_m, tile_size_n);
    die_on_error("csr2tile_row_major");
    double time_conversion = now_ms() - t0;
    printf("CSR to Tile conversion uses %.2f ms\n", time_conversion);
    /* format footprint, same formula as the reference (src/main.cu:178-185) */
    double tile_mb = ((matrixA->tilem + 1) * 4.0 + matrixA->numtile * 4.0 + (matrixA->numtile + 1) * 4.0 + matrixA->nnz * 8.0 +
                      matrixA->nnz * 1.0 + matrixA->numtile * 16.0 * 1.0 + matrixA->numtile * 16.0 * 2.0) / 1024 / 1024;
    double csr_mb = ((matrixA->m + 1) * 4.0 + matrixA->nnz * 4.0 + matrixA->nnz * 8.0) / 1024 / 1024;
    printf("tile space overhead = %.2f MB\n", tile_mb);
    csr2tile_col_major(matrixB, tile_size_m, tile_size_n);
    die_on_error("csr2tile_col_major");

    unsigned long long nnzC_computed = 0;
    double compression_rate = 0, time_tile = 0, gflops_tile = 0, ts1 = 0, ts2 = 0, ts3 = 0, tmalloc = 0;
    for (int rep = 0; rep < 2; rep++) { /* first call warms the memory pool, second is reported */
        if (rep) { matrix_destroy(matrixC); free(matrixC->tile_rowidx); }
        tilespgemm(matrixA, matrixB, matrixC, NULL, NULL, 0, 0.0, 0.0, nnzCub, &nnzC_computed, &compression_rate, &time_tile,
                   &gflops_tile, filename, &ts1, &ts2, &ts3, &tmalloc, tile_size_m, tile_size_n);
        die_on_error("tilespgemm");
    }
    printf("step1 %.3f ms, step2 %.3f ms, step3 %.3f ms, alloc %.3f ms, compression rate %.3f\n", ts1, ts2, ts3, tmalloc, compression_rate);
    /* result logs, same files and column order as the reference (src/main.cu:283-318); written only when the
     * reference's ../data directory exists next to the working directory, or under $TSG_CSV_DIR */
    {
        const char *dir = getenv("TSG_CSV_DIR") ? getenv("TSG_CSV_DIR") : "../data";
        char path[1024];
        FILE *f;
        snprintf(path, sizeof path, "%s/results_tile.csv", dir);
        if ((f = fopen(path, "a"))) {
            fprintf(f, "%s,%i,%i,%i,%llu,%llu,%f,%f,%f\n", filename, matrixA->m, matrixA->n, matrixA->nnz, nnzCub, nnzC_computed,
                    compression_rate, time_tile, gflops_tile);
            fclose(f);
            snprintf(path, sizeof path, "%s/step_runtime.csv", dir);
            if ((f = fopen(path, "a"))) {
                fprintf(f, "%s,%i,%i,%i,%llu,%llu,%f,%f,%f,%f,%f\n", filename, matrixA->m, matrixA->n, matrixA->nnz, nnzCub, nnzC_computed,
                        compression_rate, ts1, ts2, ts3, tmalloc);
                fclose(f);
            }
            snprintf(path, sizeof path, "%s/mem-cost.csv", dir);
            if ((f = fopen(path, "a"))) {
                fprintf(f, "%s,%i,%i,%i,%llu,%llu,%f,%f,%f\n", filename, matrixA->m, matrixA->n, matrixA->nnz, nnzCub, nnzC_computed,
                        compression_rate, csr_mb, tile_mb);
                fclose(f);
            }
            snprintf(path, sizeof path, "%s/preprocessing.csv", dir);
            if ((f = fopen(path, "a"))) {
                fprintf(f, "%s,%i,%i,%i,%llu,%llu,%f,%f,%f\n", filename, matrixA->m, matrixA->n, matrixA->nnz, nnzCub, nnzC_computed,
                        compression_rate, time_conversion, time_tile);
                fclose(f);
            }
        }
    }

    printf("-------------------------------check----------------------------------------\n");
    t0 = now_ms();
    tile2csr(matrixC, tile_size_m, tile_size_m);                             /* main.cu:327 */
    die_on_error("tile2csr");
    printf("tile to CSR conversion complete! (%.2f ms, nnzC = %d)\n", now_ms() - t0, matrixC->nnz);
#if TSG_DRIVER_CHECK
    if (nnzCub <= 400000000ULL) {
        int bad = check_small(matrixA, matrixB, matrixC);
        printf("serial SPA check (structure and values): %s\n", bad ? "[NOT PASSED]" : "[PASSED]");
        if (bad) return 3;
    }
#endif
    matrix_destroy(matrixA);
    matrix_destroy(matrixB);
    matrix_destroy(matrixC);
    free(matrixA->rowpointer); free(matrixA->columnindex); free(matrixA->value);
    return 0;
}
