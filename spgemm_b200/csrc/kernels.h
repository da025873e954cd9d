// kernels.h -- internal host-side entry points of the translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace tsg {

// csr2tile.cu
int tile_alloc_layout(int m, int n, int numtile, long long nnz, int col_major, tsg_dtile *out);
int csr2tile_device(const tsg_dcsr *A, int col_major, tsg_dtile *out);
int transpose_device(const tsg_dcsr *A, tsg_dcsr *AT);
int nnzcub_device(const tsg_dcsr *A, const tsg_dcsr *B, unsigned long long *out);
int masks_from_tiles_device(tsg_dtile *T);
int csr_check_device(const tsg_dcsr *A);
int csr_row_slice_device(const tsg_dcsr *A, int r0, int r1, tsg_dcsr *out);
int csr_canonicalize_device(const tsg_dcsr *A, int dup_policy, tsg_dcsr *out);
int last_input_flags();
// stable LSD radix sort of (key, value) pairs by the low key_bits bits of the key (radix_sort.cuh); the result is in
// the a or the b pair of buffers, *keys_res / *vals_res say which. n < 2^31.
int sort_pairs_device(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, long long n, int key_bits,
                      uint32_t **keys_res, uint32_t **vals_res);

// spgemm.cu
int tilerow_weights_device(const tsg_dtile *A, const tsg_dtile *B, int **d_w, int **d_jlo, int **d_jhi);
int spgemm_device(const tsg_dtile *A, const tsg_dtile *B, int trow0, int trow1, tsg_dtile *C, tsg_stats *stats);
int build_rm2csc_device(tsg_dtile *B);

// numeric.cu (step 3)
// pairs of C tile t: (a[p], b[p]) for p in [ptr[t], end[t]); b = storage id of the B tile. slot (light tile-rows only):
// for the pairs of a tile-row enumerated in A-major order (A tiles ascending, then the B tiles of B's tile-row K),
// starting at wptr[row], the position of the pair's C tile within the tile-row.
struct PairLists { const int *ptr, *end, *a, *b; const uint16_t *slot; };
struct NumericBufs {
    uint8_t *row_kind;  // [ntr]     which kernel computes the (non-dense) tiles of each C tile-row
    int *dense_list;    // [numblkC] C tiles that take the dense accumulator
    int dense_th, smem_cap;
};
size_t numeric_scratch_bytes(int ntr, long long numblkC);
int numeric_classify_device(const tsg_dtile *A, const tsg_dtile *C, int trow0, int ntr, const int *wptr, const uint8_t *light, NumericBufs *nb,
                            int *d_ns);
int numeric_device(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, int trow0, int ntr, const int *wptr, const PairLists &pl,
                   const NumericBufs &nb, const int *h_ns, bool heavy_rows, tsg_stats *stats);

// plans.cu (recipe plans: bit-exact fast path of steps 2 and 3 for matrices made of few distinct tiles)
struct PlanTable;
// tile-row templates (rowplans.cu): what k_rows_instantiate needs besides the tiled matrices
struct RowTemplates {
    int n, trow0, ntr;                    // distinct tile-row signatures; the slab
    const int *rep_list, *rep_of, *w, *wptr, *bclass;
    int *pair_ptr, *pair_end, *pair_a, *pair_b;
    const unsigned *pair_src;             // beside the representative's pairs: (A tile of the row << 16) | (tile of B's tile-row)
};
bool rowplans_env_on();
int rowplans_signatures(const tsg_dtile *A, const tsg_dtile *B, int trow0, int ntr, int *w, int *sig_slot, int *rep_of, int *bclass, int *sc_err,
                        const int **rep_list, int *nsig);
int rowplans_expand_counts(int ntr, const int *rep_of, int *cnt, uint8_t *light);
int rowplans_instantiate(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, const RowTemplates &rt, int *recipe_id);
int *rowplans_fail_ptr();
void rowplans_shutdown();
int tile_patterns_device(tsg_dtile *T);
bool plans_wanted(const tsg_dtile *A, const tsg_dtile *B);
int plans_begin(PlanTable *out);
int plans_symbolic_device(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, const PairLists &pl, const unsigned *pair_pat, const int *rslot, int *recipe_id,
                          const RowTemplates *rt, const int **d_fail);
int plans_numeric_device(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, const PairLists &pl, const int *recipe_id,
                         int trow0, int ntr, const int *wptr, int max_need, tsg_stats *stats, const RowTemplates *rt, const int *tau,
                         int max_nnzA_row, int wmax);
int plans_rowplan_build(const tsg_dtile *C, const RowTemplates &rt, const int *recipe_id, const int *pair_ptr, int **d_tau, const int **d_fail);
size_t plans_rows_need_bound(int max_nnzA_row, int maxJ, int wmax);
const int *plans_recipe_count_ptr();
void plans_shutdown();

// gentile.cu (general tile sizes: csr2tile, steps 1-3, tile2csr)
bool gtile_size_ok(int tile_rows, int tile_cols);
int gtile_csr2tile_device(const tsg_dcsr *A, int col_major, int tile_rows, int tile_cols, tsg_gtile *out);
int gtile_spgemm_device(const tsg_gtile *A, const tsg_gtile *B, tsg_gtile *C, tsg_stats *stats);
int gtile_tile2csr_device(const tsg_gtile *T, tsg_dcsr *out);
int gtile_upload(const SMatrix *h, int col_major, int tile_rows, int tile_cols, tsg_gtile *out);
int gtile_download(const tsg_gtile *t, SMatrix *h);
void gtile_free(tsg_gtile *t);
int gtile_last_input_flags();

// tile2csr.cu
int tile2csr_device(const tsg_dtile *T, tsg_dcsr *out);
int tile2csr_into(const tsg_dtile *T, int *rowptr, int *colidx, double *val, int base);
int tile_rowsums_device(const tsg_dtile *T, double *d_out, long long *d_cnt);

}  // namespace tsg
