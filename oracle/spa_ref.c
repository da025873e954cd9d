/*
 * oracle/spa_ref.c -- CPU restatement of the TileSpGEMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under spgemm_b200/ (the product) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle_vs_ref.py,
 * tests/test_golden.py) against
 *   - the reference's own CPU code compiled in place from /root/reference/src
 *     (oracle/_ref/libref_cpu.so, built by oracle/Makefile) on the 36x36
 *     UnitTest/CSR2TILE fixtures and small synthetic matrices,
 *   - UnitTest/CSR2TILE/bitmask.h (golden row masks of random_0.1_36x36),
 *   - the probe values recorded in SURVEY.md 4.3.
 * FP64 values of C are not pinned by any reference test (the cuSPARSE value
 * compare is commented out, external/cusparse/spgemm_cusparse.h:282); they are
 * pinned here by the dense-row SPA of external/cusparse/spgemm_serialref_spa.h
 * run through oracle/_ref and by scipy as a third opinion.
 *
 * All citations are relative to /root/reference/.
 * Every function is linear in its input/output size (the reference's CPU code is
 * O(tilem*tilen) / O(m*n/32)); results are identical.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_max_threads(void) { return 1; }
static int omp_get_thread_num(void) { return 0; }
static void omp_set_num_threads(int n) { (void)n; }
#endif

#define T16 16 /* the paper's tile edge; src/common.h:36 BLOCK_SIZE. MaskBits = 16 (src/common.h:146): one mask word covers 16 columns */

typedef struct {
    int m, n, tilem, tilen, numtile;
    int64_t nnz;
    int *tile_ptr;         /* [tilem+1] row-major tile structure            */
    int *tile_columnidx;   /* [numtile] ascending per tile-row              */
    int *tile_rowidx;      /* [numtile] (A: tile row; B: reference leaves 0) */
    int64_t *tile_nnz;     /* [numtile+1] exclusive offsets, storage order  */
    double *val;           /* [nnz]                                         */
    uint16_t *col;         /* [nnz]  A: r*16+c, B and C: c                  */
    uint16_t *ptr;         /* [numtile*16] per-tile exclusive row offsets   */
    uint16_t *mask;        /* [numtile*16] bit (15-c) <-> column c          */
    int *csc_tile_ptr;     /* [tilen+1]  (B only)                           */
    int *csc_tile_rowidx;  /* [numtile]  (B only)                           */
    int tr, tc;            /* rows x columns of one tile (16 x 16 unless a *_g entry point made it):
                              ptr holds tr slots per tile, mask tr*(tc/16) words per tile */
} orc_tiled;

typedef struct {
    int m, n;
    int64_t nnz;
    int64_t *rowptr; /* 64-bit: src/common.h:26-28 uses int and overflows */
    int *colidx;
    double *val;
} orc_csr;

void orc_tiled_free(orc_tiled *t)
{
    free(t->tile_ptr); free(t->tile_columnidx); free(t->tile_rowidx); free(t->tile_nnz);
    free(t->val); free(t->col); free(t->ptr); free(t->mask);
    free(t->csc_tile_ptr); free(t->csc_tile_rowidx);
    memset(t, 0, sizeof(*t));
}

void orc_csr_free(orc_csr *c)
{
    free(c->rowptr); free(c->colidx); free(c->val);
    memset(c, 0, sizeof(*c));
}

int orc_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline legs of bench.py ask for all host cores back */
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------
 * Row-major tile structure of a CSR matrix: tile_ptr / tile_columnidx.
 * Follows step1_kernel (src/csr2tile.h:6-39) + exclusive_scan (:219) and the
 * ascending tile-col emission of step2_kernel (:91-105). The per-thread flag
 * array is reset through a touched list instead of a tilen-byte memset.
 * ---------------------------------------------------------------------- */
static void tile_structure(int m, int n, const int64_t *rowptr, const int *colidx, int TR, int TC,
                           int *tilem_out, int *tilen_out, int **tile_ptr_out, int **tile_col_out)
{
    int tilem = (m + TR - 1) / TR, tilen = (n + TC - 1) / TC; /* src/csr2tile.h:210-211 */
    int *tile_ptr = (int *)calloc((size_t)tilem + 1, sizeof(int));
    int nth = omp_get_max_threads();
    char *flag_g = (char *)calloc((size_t)nth * (tilen > 0 ? tilen : 1), 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = 0; bi < tilem; bi++) {
        char *flag = flag_g + (size_t)omp_get_thread_num() * tilen;
        int r0 = bi * TR, r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        int cnt = 0;
        for (int64_t j = rowptr[r0]; j < rowptr[r1]; j++) {
            int jc = colidx[j] / TC;
            if (!flag[jc]) { flag[jc] = 1; cnt++; }
        }
        for (int64_t j = rowptr[r0]; j < rowptr[r1]; j++) flag[colidx[j] / TC] = 0;
        tile_ptr[bi + 1] = cnt;
    }
    for (int bi = 0; bi < tilem; bi++) tile_ptr[bi + 1] += tile_ptr[bi];
    int numtile = tile_ptr[tilem];
    int *tile_col = (int *)malloc(sizeof(int) * (size_t)(numtile > 0 ? numtile : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = 0; bi < tilem; bi++) {
        char *flag = flag_g + (size_t)omp_get_thread_num() * tilen;
        int r0 = bi * TR, r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        int *out = tile_col + tile_ptr[bi];
        int cnt = 0;
        for (int64_t j = rowptr[r0]; j < rowptr[r1]; j++) {
            int jc = colidx[j] / TC;
            if (!flag[jc]) { flag[jc] = 1; out[cnt++] = jc; }
        }
        for (int k = 0; k < cnt; k++) flag[out[k]] = 0;
        qsort(out, cnt, sizeof(int), cmp_int); /* quick_sort_key, src/csr2tile.h:235-239 */
    }
    free(flag_g);
    *tilem_out = tilem; *tilen_out = tilen; *tile_ptr_out = tile_ptr; *tile_col_out = tile_col;
}

static int find_tile(const int *cols, int len, int key)
{
    int lo = 0, hi = len - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (cols[mid] == key) return mid;
        if (cols[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* ------------------------------------------------------------------------
 * csr2tile_row_major (src/csr2tile.h:205-277): tiles enumerated (I asc, J asc);
 * in-tile order row-major and, within a row, CSR order (:152-168,177-184);
 * Col = r*16 + c (:192); Ptr = exclusive scan over 16 slots (:243-246), rows
 * past the matrix edge repeat the tile total; mask bit (15-c) (:195).
 * ---------------------------------------------------------------------- */
int orc_csr2tile_row_major_g(int m, int n, const int64_t *rowptr, const int *colidx, const double *val,
                             int TR, int TC, orc_tiled *out)
{
    /* general tiles (the fork's runtime tile_size_m x tile_size_n, src/csr2tile.h:205-277 with tile_size_m = TR rows and
     * tile_size_n = TC columns; TC a multiple of MaskBits, :193): Ptr has TR slots per tile, the mask of row r is
     * W = TC/16 words, column c <-> word c/16, bit 15 - c%16 (:194-195), Col = r*TC + c (:192) */
    if (TR <= 0 || TC <= 0 || TC % 16 || (int64_t)TR * TC > 65536) return -2;
    const int W = TC / 16;
    memset(out, 0, sizeof(*out));
    out->m = m; out->n = n; out->nnz = rowptr[m]; out->tr = TR; out->tc = TC;
    tile_structure(m, n, rowptr, colidx, TR, TC, &out->tilem, &out->tilen, &out->tile_ptr, &out->tile_columnidx);
    int tilem = out->tilem, numtile = out->tile_ptr[tilem];
    out->numtile = numtile;
    size_t nt = (size_t)(numtile > 0 ? numtile : 1);
    out->tile_rowidx = (int *)calloc(nt, sizeof(int));
    out->tile_nnz = (int64_t *)calloc(nt + 1, sizeof(int64_t));
    out->ptr = (uint16_t *)calloc(nt * TR, sizeof(uint16_t));
    out->mask = (uint16_t *)calloc(nt * TR * W, sizeof(uint16_t));
    out->val = (double *)calloc((size_t)(out->nnz > 0 ? out->nnz : 1), sizeof(double));
    out->col = (uint16_t *)calloc((size_t)(out->nnz > 0 ? out->nnz : 1), sizeof(uint16_t));

    /* per-tile, per-row counts (step2_kernel, src/csr2tile.h:79-89) */
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = 0; bi < tilem; bi++) {
        int t0 = out->tile_ptr[bi], nt_row = out->tile_ptr[bi + 1] - t0;
        const int *cols = out->tile_columnidx + t0;
        int r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        for (int k = 0; k < nt_row; k++) out->tile_rowidx[t0 + k] = bi; /* :97 */
        for (int row = bi * TR; row < r1; row++)
            for (int64_t j = rowptr[row]; j < rowptr[row + 1]; j++) {
                int k = find_tile(cols, nt_row, colidx[j] / TC);
                out->ptr[(size_t)(t0 + k) * TR + (row - bi * TR)]++;
                out->tile_nnz[t0 + k + 1]++;
            }
    }
    for (int t = 0; t < numtile; t++) out->tile_nnz[t + 1] += out->tile_nnz[t]; /* :241 */
#pragma omp parallel for
    for (int t = 0; t < numtile; t++) { /* exclusive_scan_uint16 over 16 slots, :243-246 */
        uint16_t *p = out->ptr + (size_t)t * TR, run = 0;
        for (int r = 0; r < TR; r++) { uint16_t c = p[r]; p[r] = run; run += c; }
    }
    /* scatter (step3_kernel, :112-203) */
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = 0; bi < tilem; bi++) {
        int t0 = out->tile_ptr[bi], nt_row = out->tile_ptr[bi + 1] - t0;
        const int *cols = out->tile_columnidx + t0;
        int r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        uint16_t *cursor = (uint16_t *)calloc((size_t)(nt_row > 0 ? nt_row : 1) * TR, sizeof(uint16_t));
        for (int row = bi * TR; row < r1; row++) {
            int r = row - bi * TR;
            for (int64_t j = rowptr[row]; j < rowptr[row + 1]; j++) {
                int jc = colidx[j] / TC, c = colidx[j] - jc * TC;
                int k = find_tile(cols, nt_row, jc);
                size_t t = (size_t)(t0 + k);
                int64_t dst = out->tile_nnz[t] + out->ptr[t * TR + r] + cursor[(size_t)k * TR + r]++;
                out->val[dst] = val[j];
                out->col[dst] = (uint16_t)(r * TC + c);
                out->mask[(t * TR + r) * W + (c >> 4)] |= (uint16_t)(1u << (15 - (c & 15)));
            }
        }
        free(cursor);
    }
    return 0;
}

int orc_csr2tile_row_major(int m, int n, const int64_t *rowptr, const int *colidx, const double *val,
                           orc_tiled *out)
{
    return orc_csr2tile_row_major_g(m, n, rowptr, colidx, val, T16, T16, out);
}

/* CSR -> CSC, stable (matrix_transposition, src/utils.h:161-198). */
void orc_transpose(int m, int n, const int64_t *rowptr, const int *colidx, const double *val,
                   int64_t *colptr, int *rowidx, double *cscval)
{
    memset(colptr, 0, sizeof(int64_t) * ((size_t)n + 1));
    int64_t nnz = rowptr[m];
    for (int64_t i = 0; i < nnz; i++) colptr[colidx[i] + 1]++;
    for (int j = 0; j < n; j++) colptr[j + 1] += colptr[j];
    int64_t *incr = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n + 1));
    memcpy(incr, colptr, sizeof(int64_t) * ((size_t)n + 1));
    for (int row = 0; row < m; row++)
        for (int64_t j = rowptr[row]; j < rowptr[row + 1]; j++) {
            int64_t d = incr[colidx[j]]++;
            rowidx[d] = row;
            if (cscval) cscval[d] = val[j];
        }
    free(incr);
}

/* ------------------------------------------------------------------------
 * csr2tile_col_major (src/csr2tile.h:279-506): tile_ptr/tile_columnidx as the
 * row-major structure (:331-333,349-388); everything else in CSC-tile order
 * (J asc, I asc): csc_tile_ptr / csc_tile_rowidx / tile_nnz (:345-347); tiles
 * internally row-major with ascending columns (double transposition :455-457),
 * Col = c (:475), Ptr padded with the tile total past the edge (:477-483),
 * mask (:474). tile_rowidx is allocated and left zero (:336-337).
 * ---------------------------------------------------------------------- */
int orc_csr2tile_col_major_g(int m, int n, const int64_t *rowptr, const int *colidx, const double *val,
                             int TR, int TC, orc_tiled *out)
{
    /* general tiles: a tile of B has TR rows and TC columns. The reference's csr2tile_col_major(B, tile_size_m,
     * tile_size_n) makes them tile_size_n x tile_size_m ("the tile of A is m x n, and the tile of B is n x m",
     * src/main.cu:84; tilem = ceil(m/tile_size_n), tilen = ceil(n/tile_size_m), src/csr2tile.h:282-283), so the caller
     * passes TR = tile_size_n, TC = tile_size_m. Ptr: TR slots per tile (:477-483), mask: TR rows of TC/16 words (:472-474). */
    if (TR <= 0 || TC <= 0 || TC % 16 || (int64_t)TR * TC > 65536) return -2;
    const int W = TC / 16;
    memset(out, 0, sizeof(*out));
    out->m = m; out->n = n; out->nnz = rowptr[m]; out->tr = TR; out->tc = TC;
    tile_structure(m, n, rowptr, colidx, TR, TC, &out->tilem, &out->tilen, &out->tile_ptr, &out->tile_columnidx);
    int tilem = out->tilem, tilen = out->tilen, numtile = out->tile_ptr[tilem];
    out->numtile = numtile;
    size_t nt = (size_t)(numtile > 0 ? numtile : 1);
    int64_t nnz = out->nnz;
    out->tile_rowidx = (int *)calloc(nt, sizeof(int));
    out->tile_nnz = (int64_t *)calloc(nt + 1, sizeof(int64_t));
    out->ptr = (uint16_t *)calloc(nt * TR, sizeof(uint16_t));
    out->mask = (uint16_t *)calloc(nt * TR * W, sizeof(uint16_t));
    out->val = (double *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(double));
    out->col = (uint16_t *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(uint16_t));
    out->csc_tile_ptr = (int *)calloc((size_t)tilen + 1, sizeof(int));
    out->csc_tile_rowidx = (int *)calloc(nt, sizeof(int));

    /* tile-level transposition: stable counting sort of (I,J) by J */
    for (int t = 0; t < numtile; t++) out->csc_tile_ptr[out->tile_columnidx[t] + 1]++;
    for (int j = 0; j < tilen; j++) out->csc_tile_ptr[j + 1] += out->csc_tile_ptr[j];
    int *incr = (int *)malloc(sizeof(int) * ((size_t)tilen + 1));
    memcpy(incr, out->csc_tile_ptr, sizeof(int) * ((size_t)tilen + 1));
    int *rm2csc = (int *)malloc(sizeof(int) * nt); /* row-major tile index -> CSC tile id */
    for (int bi = 0; bi < tilem; bi++)
        for (int t = out->tile_ptr[bi]; t < out->tile_ptr[bi + 1]; t++) {
            int d = incr[out->tile_columnidx[t]]++;
            out->csc_tile_rowidx[d] = bi;
            rm2csc[t] = d;
        }
    free(incr);

    /* counts per tile and per tile row (CSC ids) */
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = 0; bi < tilem; bi++) {
        int t0 = out->tile_ptr[bi], nt_row = out->tile_ptr[bi + 1] - t0;
        const int *cols = out->tile_columnidx + t0;
        int r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        for (int row = bi * TR; row < r1; row++)
            for (int64_t j = rowptr[row]; j < rowptr[row + 1]; j++) {
                size_t t = (size_t)rm2csc[t0 + find_tile(cols, nt_row, colidx[j] / TC)];
                out->ptr[t * TR + (row - bi * TR)]++;
                out->tile_nnz[t + 1]++;
            }
    }
    for (int t = 0; t < numtile; t++) out->tile_nnz[t + 1] += out->tile_nnz[t];
#pragma omp parallel for
    for (int t = 0; t < numtile; t++) {
        uint16_t *p = out->ptr + (size_t)t * TR, run = 0;
        for (int r = 0; r < TR; r++) { uint16_t c = p[r]; p[r] = run; run += c; }
    }
    /* scatter; columns inside a (tile,row) run end up ascending and stable, which is
     * what transposing the tile twice produces (:455-457). */
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = 0; bi < tilem; bi++) {
        int t0 = out->tile_ptr[bi], nt_row = out->tile_ptr[bi + 1] - t0;
        const int *cols = out->tile_columnidx + t0;
        int r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        uint16_t *cursor = (uint16_t *)calloc((size_t)(nt_row > 0 ? nt_row : 1) * TR, sizeof(uint16_t));
        for (int row = bi * TR; row < r1; row++) {
            int r = row - bi * TR;
            for (int64_t j = rowptr[row]; j < rowptr[row + 1]; j++) {
                int jc = colidx[j] / TC, c = colidx[j] - jc * TC;
                int k = find_tile(cols, nt_row, jc);
                size_t t = (size_t)rm2csc[t0 + k];
                int64_t base = out->tile_nnz[t] + out->ptr[t * TR + r];
                int pos = cursor[(size_t)k * TR + r]++;
                /* stable insertion by column keeps duplicates in CSR order */
                while (pos > 0 && out->col[base + pos - 1] > (uint16_t)c) {
                    out->col[base + pos] = out->col[base + pos - 1];
                    out->val[base + pos] = out->val[base + pos - 1];
                    pos--;
                }
                out->col[base + pos] = (uint16_t)c;
                out->val[base + pos] = val[j];
                out->mask[(t * TR + r) * W + (c >> 4)] |= (uint16_t)(1u << (15 - (c & 15)));
            }
        }
        free(cursor);
    }
    free(rm2csc);
    return 0;
}

int orc_csr2tile_col_major(int m, int n, const int64_t *rowptr, const int *colidx, const double *val,
                           orc_tiled *out)
{
    return orc_csr2tile_col_major_g(m, n, rowptr, colidx, val, T16, T16, out);
}

/* nnzCub loop of the driver (src/main.cu:155-160). */
uint64_t orc_nnzcub(int64_t nnzA, const int *colidxA, const int64_t *rowptrB)
{
    uint64_t s = 0;
#pragma omp parallel for reduction(+ : s)
    for (int64_t i = 0; i < nnzA; i++) s += (uint64_t)(rowptrB[colidxA[i] + 1] - rowptrB[colidxA[i]]);
    return s;
}

/* ------------------------------------------------------------------------
 * Row-wise SPA SpGEMM with values. Value semantics = compute_dense_row
 * (src/external/cusparse/spgemm_serialref_spa.h:7-31): for row i, for each A
 * entry in CSR order, for each B entry of that row in CSR order,
 * dense[col] += a*b; structure = flag set by any product, explicit zeros kept,
 * columns emitted ascending (:89-112; src/spgemm_serialref_spa_new.h:63-103).
 * Restated with a per-thread dense accumulator reset through a touched list
 * (O(products)) and 64-bit row pointers. Rows [row0,row1) only (slab runs).
 * ---------------------------------------------------------------------- */
int orc_spgemm_spa(int mA, int nB, const int64_t *rpA, const int *ciA, const double *vA,
                   const int64_t *rpB, const int *ciB, const double *vB,
                   int row0, int row1, orc_csr *C)
{
    memset(C, 0, sizeof(*C));
    if (row0 < 0) row0 = 0;
    if (row1 > mA) row1 = mA;
    int rows = row1 - row0;
    C->m = rows; C->n = nB;
    C->rowptr = (int64_t *)calloc((size_t)rows + 1, sizeof(int64_t));
    int nth = omp_get_max_threads();
    char *flag_g = (char *)calloc((size_t)nth * (nB > 0 ? nB : 1), 1);
    double *acc_g = (double *)calloc((size_t)nth * (nB > 0 ? nB : 1), sizeof(double));
    int **touch_g = (int **)calloc(nth, sizeof(int *));
    size_t *tcap = (size_t *)calloc(nth, sizeof(size_t));

    /* pass 1: counts (get_nnzC_only, src/spgemm_serialref_spa_new.h:29-62) */
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = row0; i < row1; i++) {
        int tid = omp_get_thread_num();
        char *flag = flag_g + (size_t)tid * nB;
        size_t cnt = 0;
        for (int64_t ja = rpA[i]; ja < rpA[i + 1]; ja++) {
            int k = ciA[ja];
            for (int64_t jb = rpB[k]; jb < rpB[k + 1]; jb++) {
                int c = ciB[jb];
                if (!flag[c]) {
                    flag[c] = 1;
                    if (cnt == tcap[tid]) {
                        tcap[tid] = tcap[tid] ? tcap[tid] * 2 : 1024;
                        touch_g[tid] = (int *)realloc(touch_g[tid], tcap[tid] * sizeof(int));
                    }
                    touch_g[tid][cnt++] = c;
                }
            }
        }
        for (size_t k = 0; k < cnt; k++) flag[touch_g[tid][k]] = 0;
        C->rowptr[i - row0 + 1] = (int64_t)cnt;
    }
    int64_t maxrow = 1;
    for (int i = 0; i < rows; i++) if (C->rowptr[i + 1] > maxrow) maxrow = C->rowptr[i + 1];
    for (int t = 0; t < nth; t++) { /* pass 2 may give a row to a different thread */
        touch_g[t] = (int *)realloc(touch_g[t], (size_t)maxrow * sizeof(int));
        tcap[t] = (size_t)maxrow;
    }
    for (int i = 0; i < rows; i++) C->rowptr[i + 1] += C->rowptr[i];
    C->nnz = C->rowptr[rows];
    C->colidx = (int *)malloc(sizeof(int) * (size_t)(C->nnz > 0 ? C->nnz : 1));
    C->val = (double *)malloc(sizeof(double) * (size_t)(C->nnz > 0 ? C->nnz : 1));

    /* pass 2: fill */
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = row0; i < row1; i++) {
        int tid = omp_get_thread_num();
        char *flag = flag_g + (size_t)tid * nB;
        double *acc = acc_g + (size_t)tid * nB;
        int *touch = touch_g[tid];
        size_t cnt = 0;
        for (int64_t ja = rpA[i]; ja < rpA[i + 1]; ja++) {
            int k = ciA[ja];
            double a = vA[ja];
            for (int64_t jb = rpB[k]; jb < rpB[k + 1]; jb++) {
                int c = ciB[jb];
                if (!flag[c]) { flag[c] = 1; touch[cnt++] = c; }
                acc[c] += a * vB[jb];
            }
        }
        qsort(touch, cnt, sizeof(int), cmp_int);
        int64_t o = C->rowptr[i - row0];
        for (size_t k = 0; k < cnt; k++) {
            int c = touch[k];
            C->colidx[o + (int64_t)k] = c;
            C->val[o + (int64_t)k] = acc[c];
            acc[c] = 0.0; flag[c] = 0;
        }
    }
    for (int t = 0; t < nth; t++) free(touch_g[t]);
    free(touch_g); free(tcap); free(flag_g); free(acc_g);
    return 0;
}

/* ------------------------------------------------------------------------
 * Count-only SPA: nnz of every row of C = A*B for rows [row0,row1), the
 * get_nnzC_only pass of the reference protocol
 * (src/spgemm_serialref_spa_new.h:29-62; call protocol
 * src/external/cusparse/main.cu:196-212). Used where C itself is too large to
 * build on the host (R-MAT scale 20+: nnz(C) ~ 1e10): bench.py compares these
 * counts and A*(B*1) with the device's per-row counts and sums.
 * ---------------------------------------------------------------------- */
int orc_spgemm_rowcounts(int mA, int nB, const int64_t *rpA, const int *ciA, const int64_t *rpB, const int *ciB,
                         int row0, int row1, int64_t *counts)
{
    if (row0 < 0) row0 = 0;
    if (row1 > mA) row1 = mA;
    int nth = omp_get_max_threads();
    char *flag_g = (char *)calloc((size_t)nth * (nB > 0 ? nB : 1), 1);
    int **touch_g = (int **)calloc(nth, sizeof(int *));
    size_t *tcap = (size_t *)calloc(nth, sizeof(size_t));
    if (!flag_g || !touch_g || !tcap) return 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = row0; i < row1; i++) {
        int tid = omp_get_thread_num();
        char *flag = flag_g + (size_t)tid * nB;
        size_t cnt = 0;
        for (int64_t ja = rpA[i]; ja < rpA[i + 1]; ja++) {
            int k = ciA[ja];
            for (int64_t jb = rpB[k]; jb < rpB[k + 1]; jb++) {
                int c = ciB[jb];
                if (!flag[c]) {
                    flag[c] = 1;
                    if (cnt == tcap[tid]) {
                        tcap[tid] = tcap[tid] ? tcap[tid] * 2 : 1024;
                        touch_g[tid] = (int *)realloc(touch_g[tid], tcap[tid] * sizeof(int));
                    }
                    touch_g[tid][cnt++] = c;
                }
            }
        }
        for (size_t k = 0; k < cnt; k++) flag[touch_g[tid][k]] = 0;
        counts[i - row0] = (int64_t)cnt;
    }
    for (int t = 0; t < nth; t++) free(touch_g[t]);
    free(touch_g); free(tcap); free(flag_g);
    return 0;
}

/* ------------------------------------------------------------------------
 * C in tiled form, as tilespgemm must return it (SURVEY Appendix A "C"):
 *  - tile list = tile-level boolean product of A's and B's tile patterns
 *    (step 1, src/tilespgemm-cuda.h:298-311), (I asc, J asc), INCLUDING tiles
 *    that end up with zero entries;
 *  - per tile: mask / Ptr (exclusive; all zero for an empty tile) / Col = c
 *    ascending (:1365-1369) / Val, derived from CSR(C);
 *  - tile_nnz exclusive offsets (:2602).
 * Tile rows [trow0,trow1) of A only; csrC must hold rows [trow0*16, min(trow1*16,m)).
 * ---------------------------------------------------------------------- */
int orc_ctiles_from_csr_g(int m, int n, int tilemA, const int *tile_ptrA, const int *tile_colA,
                          int tilenB, const int *tile_ptrB, const int *tile_colB,
                          int trow0, int trow1,
                          const int64_t *rpC, const int *ciC, const double *vC, int TR, int TC, orc_tiled *out)
{
    /* general tiles: a tile of C has TR = tile_size_m rows (A's tile rows) and TC = tile_size_m columns (B's tile
     * columns); the driver hands tile2csr(C, tile_size_m, tile_size_m), src/main.cu:327 */
    if (TR <= 0 || TC <= 0 || TC % 16 || (int64_t)TR * TC > 65536) return -2;
    const int W = TC / 16;
    memset(out, 0, sizeof(*out));
    out->tr = TR; out->tc = TC;
    if (trow0 < 0) trow0 = 0;
    if (trow1 > tilemA) trow1 = tilemA;
    int trows = trow1 - trow0;
    out->n = n; out->tilem = trows; out->tilen = tilenB;
    out->m = (trow1 * TR < m ? trow1 * TR : m) - trow0 * TR; /* a slab is a matrix of its own */
    if (out->m < 0) out->m = 0;
    out->tile_ptr = (int *)calloc((size_t)trows + 1, sizeof(int));
    int nth = omp_get_max_threads();
    char *flag_g = (char *)calloc((size_t)nth * (tilenB > 0 ? tilenB : 1), 1);
    int **touch_g = (int **)calloc(nth, sizeof(int *));
    size_t *tcap = (size_t *)calloc(nth, sizeof(size_t));
    for (int pass = 0; pass < 2; pass++) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int bi = trow0; bi < trow1; bi++) {
            int tid = omp_get_thread_num();
            char *flag = flag_g + (size_t)tid * tilenB;
            size_t cnt = 0;
            for (int ta = tile_ptrA[bi]; ta < tile_ptrA[bi + 1]; ta++) {
                int k = tile_colA[ta];
                for (int tb = tile_ptrB[k]; tb < tile_ptrB[k + 1]; tb++) {
                    int j = tile_colB[tb];
                    if (!flag[j]) {
                        flag[j] = 1;
                        if (cnt == tcap[tid]) {
                            tcap[tid] = tcap[tid] ? tcap[tid] * 2 : 1024;
                            touch_g[tid] = (int *)realloc(touch_g[tid], tcap[tid] * sizeof(int));
                        }
                        touch_g[tid][cnt++] = j;
                    }
                }
            }
            for (size_t k = 0; k < cnt; k++) flag[touch_g[tid][k]] = 0;
            if (pass == 0) out->tile_ptr[bi - trow0 + 1] = (int)cnt;
            else {
                qsort(touch_g[tid], cnt, sizeof(int), cmp_int);
                int o = out->tile_ptr[bi - trow0];
                for (size_t k = 0; k < cnt; k++) {
                    out->tile_columnidx[o + (int)k] = touch_g[tid][k];
                    out->tile_rowidx[o + (int)k] = bi;
                }
            }
        }
        if (pass == 0) {
            for (int i = 0; i < trows; i++) out->tile_ptr[i + 1] += out->tile_ptr[i];
            out->numtile = out->tile_ptr[trows];
            size_t nt = (size_t)(out->numtile > 0 ? out->numtile : 1);
            out->tile_columnidx = (int *)calloc(nt, sizeof(int));
            out->tile_rowidx = (int *)calloc(nt, sizeof(int));
            out->tile_nnz = (int64_t *)calloc(nt + 1, sizeof(int64_t));
            out->ptr = (uint16_t *)calloc(nt * TR, sizeof(uint16_t));
            out->mask = (uint16_t *)calloc(nt * TR * W, sizeof(uint16_t));
        }
    }
    for (int t = 0; t < nth; t++) free(touch_g[t]);
    free(touch_g); free(tcap); free(flag_g);

    int row_base = trow0 * TR;
    int row_end = trow1 * TR < m ? trow1 * TR : m;
    int64_t nnz = rpC[row_end - row_base];
    out->nnz = nnz;
    out->val = (double *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(double));
    out->col = (uint16_t *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(uint16_t));
    int bad = 0;
    /* counts */
#pragma omp parallel for schedule(dynamic, 64) reduction(| : bad)
    for (int bi = trow0; bi < trow1; bi++) {
        int t0 = out->tile_ptr[bi - trow0], ntr = out->tile_ptr[bi - trow0 + 1] - t0;
        const int *cols = out->tile_columnidx + t0;
        int r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        for (int row = bi * TR; row < r1; row++)
            for (int64_t j = rpC[row - row_base]; j < rpC[row - row_base + 1]; j++) {
                int k = find_tile(cols, ntr, ciC[j] / TC);
                if (k < 0) { bad = 1; continue; } /* C entry outside the tile-level product: impossible */
                out->ptr[(size_t)(t0 + k) * TR + (row - bi * TR)]++;
                out->tile_nnz[t0 + k + 1]++;
            }
    }
    if (bad) return -1;
    for (int t = 0; t < out->numtile; t++) out->tile_nnz[t + 1] += out->tile_nnz[t];
#pragma omp parallel for
    for (int t = 0; t < out->numtile; t++) {
        uint16_t *p = out->ptr + (size_t)t * TR, run = 0;
        for (int r = 0; r < TR; r++) { uint16_t c = p[r]; p[r] = run; run += c; }
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int bi = trow0; bi < trow1; bi++) {
        int t0 = out->tile_ptr[bi - trow0], ntr = out->tile_ptr[bi - trow0 + 1] - t0;
        const int *cols = out->tile_columnidx + t0;
        int r1 = (bi + 1) * TR < m ? (bi + 1) * TR : m;
        uint16_t *cursor = (uint16_t *)calloc((size_t)(ntr > 0 ? ntr : 1) * TR, sizeof(uint16_t));
        for (int row = bi * TR; row < r1; row++) {
            int r = row - bi * TR;
            for (int64_t j = rpC[row - row_base]; j < rpC[row - row_base + 1]; j++) {
                int jc = ciC[j] / TC, c = ciC[j] - jc * TC;
                int k = find_tile(cols, ntr, jc);
                size_t t = (size_t)(t0 + k);
                int64_t dst = out->tile_nnz[t] + out->ptr[t * TR + r] + cursor[(size_t)k * TR + r]++;
                out->val[dst] = vC[j];
                out->col[dst] = (uint16_t)c;
                out->mask[(t * TR + r) * W + (c >> 4)] |= (uint16_t)(1u << (15 - (c & 15)));
            }
        }
        free(cursor);
    }
    return 0;
}

int orc_ctiles_from_csr(int m, int n, int tilemA, const int *tile_ptrA, const int *tile_colA,
                        int tilenB, const int *tile_ptrB, const int *tile_colB,
                        int trow0, int trow1,
                        const int64_t *rpC, const int *ciC, const double *vC, orc_tiled *out)
{
    return orc_ctiles_from_csr_g(m, n, tilemA, tile_ptrA, tile_colA, tilenB, tile_ptrB, tile_colB, trow0, trow1, rpC, ciC, vC,
                                 T16, T16, out);
}

/* ------------------------------------------------------------------------
 * tile2csr (src/tile2csr.h:72-140): pass 1 adds per-tile row counts (:8-32),
 * exclusive scan (:101), pass 2 appends (tile_col*16 + Col, Val) per row with a
 * running cursor visiting tiles in (tile-row, ascending tile-col) order
 * (:118-139, :34-68). Explicit zeros are kept (:27-28,59-60).
 * Col is added as stored (:57), so this is meaningful for B/C-style Col = c only.
 * A slab produced by orc_ctiles_from_csr is a matrix of its own (m = rows of the slab).
 * ---------------------------------------------------------------------- */
int orc_tile2csr(const orc_tiled *t, orc_csr *out)
{
    /* tile size from the tiled matrix (0 = a caller that never set it: 16 x 16) */
    const int TR = t->tr > 0 ? t->tr : T16, TC = t->tc > 0 ? t->tc : T16;
    memset(out, 0, sizeof(*out));
    int m = t->m, tilem = t->tilem;
    const int row_base = 0;
    out->m = m; out->n = t->n;
    out->rowptr = (int64_t *)calloc((size_t)m + 1, sizeof(int64_t));
    for (int bi = 0; bi < tilem; bi++) {
        int rowlen = (bi == tilem - 1) ? m - (tilem - 1) * TR : TR; /* :19 */
        for (int tt = t->tile_ptr[bi]; tt < t->tile_ptr[bi + 1]; tt++) {
            int64_t tnnz = t->tile_nnz[tt + 1] - t->tile_nnz[tt];
            const uint16_t *p = t->ptr + (size_t)tt * TR;
            for (int r = 0; r < rowlen; r++) {
                int64_t end = (r == rowlen - 1) ? tnnz : p[r + 1]; /* :22 */
                out->rowptr[row_base + bi * TR + r + 1] += end - p[r];
            }
        }
    }
    for (int i = 0; i < m; i++) out->rowptr[i + 1] += out->rowptr[i];
    out->nnz = out->rowptr[m];
    out->colidx = (int *)calloc((size_t)(out->nnz > 0 ? out->nnz : 1), sizeof(int));
    out->val = (double *)calloc((size_t)(out->nnz > 0 ? out->nnz : 1), sizeof(double));
    int64_t *cursor = (int64_t *)calloc((size_t)(m > 0 ? m : 1), sizeof(int64_t));
    for (int bi = 0; bi < tilem; bi++) {
        int rowlen = (bi == tilem - 1) ? m - (tilem - 1) * TR : TR;
        for (int tt = t->tile_ptr[bi]; tt < t->tile_ptr[bi + 1]; tt++) {
            int64_t tnnz = t->tile_nnz[tt + 1] - t->tile_nnz[tt], base = t->tile_nnz[tt];
            const uint16_t *p = t->ptr + (size_t)tt * TR;
            int tc = t->tile_columnidx[tt];
            for (int r = 0; r < rowlen; r++) {
                int64_t end = (r == rowlen - 1) ? tnnz : p[r + 1];
                for (int64_t j = p[r]; j < end; j++) {
                    int row = bi * TR + r;
                    int64_t d = out->rowptr[row] + cursor[row]++;
                    out->colidx[d] = tc * TC + t->col[base + j];
                    out->val[d] = t->val[base + j];
                }
            }
        }
    }
    free(cursor);
    return 0;
}

/* Per-tile-row step-1 weight w_I = sum over A tiles (I,K) of |B tile-row K|
 * (the number of matched tile pairs of tile-row I; nsparse set_intprod_num,
 * src/spgemm_nsparse_kernel.h:135-151). Used by the multi-GPU partitioner tests. */
void orc_tilerow_weights(int tilemA, const int *tile_ptrA, const int *tile_colA,
                         const int *tile_ptrB, int64_t *w)
{
#pragma omp parallel for
    for (int bi = 0; bi < tilemA; bi++) {
        int64_t s = 0;
        for (int ta = tile_ptrA[bi]; ta < tile_ptrA[bi + 1]; ta++) {
            int k = tile_colA[ta];
            s += tile_ptrB[k + 1] - tile_ptrB[k];
        }
        w[bi] = s;
    }
}
