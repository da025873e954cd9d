// gentile_emu.cpp -- TEST INFRASTRUCTURE: spgemm_b200/csrc/gentile.cu compiled as plain C++ (see gentile_emu.h) behind
// the same three host-buffer calls the library's drop-in entry points make for a general tile size
// (csrc/api.cu: csr2tile_host_general, tilespgemm, tile2csr). Built by tests/test_gentile_emu.py with g++.
#define GT_EMULATE 1
#include "gentile_emu.h"
#include "../../spgemm_b200/csrc/gentile.cu"

using namespace tsg;

extern "C" {

int emu_last_error(void) { return g_emu_err; }
const char *emu_last_error_string(void) { return g_emu_msg; }
void emu_clear_error(void) { g_emu_err = 0; g_emu_msg[0] = 0; }
long long emu_launches(void) { return g_emu_ctx.launches; }
void emu_free(void *p) { free(p); }

// csr2tile_row_major / csr2tile_col_major(matrix, tile_size_m, tile_size_n) for a general tile size
int emu_csr2tile(SMatrix *mat, int tm, int tn, int col_major)
{
    const int TR = col_major ? tn : tm, TC = col_major ? tm : tn;
    tsg_dcsr A;
    memset(&A, 0, sizeof(A));
    A.m = mat->m; A.n = mat->n; A.nnz = mat->rowpointer[mat->m];
    A.rowptr = mat->rowpointer; A.colidx = mat->columnindex; A.val = mat->value;
    tsg_gtile t;
    int rc = gtile_csr2tile_device(&A, col_major, TR, TC, &t);
    if (rc == TSG_OK) rc = gtile_download(&t, mat);
    gtile_free(&t);
    return rc ? rc : g_emu_err;
}

int emu_tilespgemm(SMatrix *A, SMatrix *B, SMatrix *C, int tm, int tn, tsg_stats *st)
{
    tsg_gtile gA, gB, gC;
    memset(&gA, 0, sizeof(gA)); memset(&gB, 0, sizeof(gB)); memset(&gC, 0, sizeof(gC));
    int rc = gtile_upload(A, 0, tm, tn, &gA);
    if (!rc) rc = gtile_upload(B, 1, tn, tm, &gB);
    if (!rc) rc = gtile_spgemm_device(&gA, &gB, &gC, st);
    if (!rc) rc = gtile_download(&gC, C);
    gtile_free(&gA); gtile_free(&gB); gtile_free(&gC);
    return rc ? rc : g_emu_err;
}

int emu_tile2csr(SMatrix *mat, int TR, int TC)
{
    tsg_gtile g;
    tsg_dcsr c;
    memset(&g, 0, sizeof(g)); memset(&c, 0, sizeof(c));
    int rc = gtile_upload(mat, 0, TR, TC, &g);
    if (!rc) rc = gtile_tile2csr_device(&g, &c);
    if (!rc) {
        mat->rowpointer = (int *)malloc(((size_t)c.m + 1) * 4);
        mat->columnindex = (int *)malloc((size_t)(c.nnz > 0 ? c.nnz : 1) * 4);
        mat->value = (double *)malloc((size_t)(c.nnz > 0 ? c.nnz : 1) * 8);
        memcpy(mat->rowpointer, c.rowptr, ((size_t)c.m + 1) * 4);
        if (c.nnz) { memcpy(mat->columnindex, c.colidx, (size_t)c.nnz * 4); memcpy(mat->value, c.val, (size_t)c.nnz * 8); }
        mat->nnz = (int)c.nnz;
    }
    if (c.owner) dfree(c.owner);
    gtile_free(&g);
    return rc ? rc : g_emu_err;
}

}  // extern "C"
