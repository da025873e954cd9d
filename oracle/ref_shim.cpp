/*
 * oracle/ref_shim.cpp -- thin extern "C" wrappers around the reference's OWN CPU code,
 * compiled from the sources where they lie under /root/reference/src (nothing is copied).
 * TEST INFRASTRUCTURE ONLY: builds into oracle/_ref/libref_cpu.so (git-ignored), used by
 * tests/ to validate the restatement in oracle/spa_ref.c and by tests/golden/make_golden.py
 * to generate the committed golden vectors. Never on the product path.
 *
 * Two tricks are needed to compile the reference headers with plain g++:
 *  - csr2tile.h uses assert() without including <assert.h>  (src/csr2tile.h:193);
 *  - common.h:18 pulls a vendored CUDA-11 cuda_fp16.h that does not compile under CUDA 12.9 /
 *    host-only; defining its include guard turns it into an empty header.
 */
#include <assert.h>
#define __CUDA_FP16_H__
#ifndef DEBUG_PRINT_ENABLE
#define DEBUG_PRINT_ENABLE 0
#endif
#include "common.h"
#include "utils.h"
#include "csr2tile.h"
#include "tile2csr.h"
#include "spgemm_serialref_spa_new.h"

extern "C" {

int ref_sizeof_smatrix(void) { return (int)sizeof(SMatrix); }

/* src/csr2tile.h:205 / :279. The struct is caller-allocated (sizeof from ref_sizeof_smatrix)
 * with m,n,nnz,rowpointer,columnindex,value filled, exactly like src/main.cu:77-152. */
void ref_csr2tile_row_major(void *mat) { csr2tile_row_major((SMatrix *)mat, 16, 16); }
void ref_csr2tile_col_major(void *mat) { csr2tile_col_major((SMatrix *)mat, 16, 16); }
/* src/tile2csr.h:72 (main.cu:327 passes (tile_size_m, tile_size_m)) */
void ref_tile2csr(void *mat) { tile2csr((SMatrix *)mat, 16, 16); }
void ref_matrix_destroy(void *mat) { matrix_destroy((SMatrix *)mat); }
/* the fork's runtime tile sizes (src/main.cu:84-91: "the tile of A is m x n, and the tile of B is n x m") */
void ref_csr2tile_row_major_g(void *mat, int tile_size_m, int tile_size_n) { csr2tile_row_major((SMatrix *)mat, tile_size_m, tile_size_n); }
void ref_csr2tile_col_major_g(void *mat, int tile_size_m, int tile_size_n) { csr2tile_col_major((SMatrix *)mat, tile_size_m, tile_size_n); }
void ref_tile2csr_g(void *mat, int tile_size_m, int tile_size_n) { tile2csr((SMatrix *)mat, tile_size_m, tile_size_n); }

/* src/utils.h:161 */
void ref_matrix_transposition(int m, int n, int nnz, const int *rp, const int *ci, const double *v,
                              int *cscRowIdx, int *cscColPtr, double *cscVal)
{
    matrix_transposition(m, n, nnz, rp, ci, v, cscRowIdx, cscColPtr, cscVal);
}

/* src/spgemm_serialref_spa_new.h:7, two-pass protocol of external/cusparse/main.cu:196-212 */
void ref_spgemm_spa(const int *rpA, const int *ciA, const double *vA, int mA, int nA, int nnzA,
                    const int *rpB, const int *ciB, const double *vB, int mB, int nB, int nnzB,
                    int *rpC, int *ciC, double *vC, int mC, int nC, int *nnzC, int get_nnzC_only)
{
    spgemm_spa(rpA, ciA, vA, mA, nA, nnzA, rpB, ciB, vB, mB, nB, nnzB, rpC, ciC, vC, mC, nC, nnzC, get_nnzC_only);
}

void ref_free(void *p) { free(p); }

} /* extern "C" */
