/*
 * tilespgemm.h -- C ABI of libtilespgemm_b200.so, the B200-native (sm_100a) drop-in for the
 * TileSpGEMM hot path of for-the-juan/SpGEMM.
 *
 * Part 1 re-declares, name for name and argument for argument, the functions the reference
 * driver (src/main.cu) gets by #including csr2tile.h / tilespgemm-cuda.h / tile2csr.h / utils.h;
 * replacing those includes by this header and linking the library is the whole integration
 * (INTEGRATION.md). Part 2 is an additive device-resident API (prefix tsg_) that keeps the
 * matrices in HBM between csr2tile -> SpGEMM -> tile2csr, runs C tile-row slabs, and is what
 * the multi-GPU path and bench.py drive.
 *
 * Plain C: pointers and sizes only. Every kernel behind these entry points is hand-written CUDA
 * for sm_100a; there is no CPU fallback -- without a usable CUDA device every call fails loudly
 * (error code + message, see tilespgemm_last_error()).
 *
 * All "reference" citations are relative to the reference repository root.
 */
#ifndef TILESPGEMM_B200_H
#define TILESPGEMM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Types. SMatrix is bit-identical to the reference struct (src/common.h:150-172):
 * 20 fields, int sizes, raw host pointers, MAT_VAL_TYPE = double, MAT_PTR_TYPE = int,
 * TILE_CSR_{PTR,COL}_TYPE = TILE_MASK_TYPE = uint16_t (src/common.h:140-143).
 * ---------------------------------------------------------------------------------------- */
#ifndef MAT_VAL_TYPE
#define MAT_VAL_TYPE double
#endif
#ifndef MAT_PTR_TYPE
#define MAT_PTR_TYPE int
#endif
typedef uint16_t TILE_CSR_PTR_TYPE;
typedef uint16_t TILE_CSR_COL_TYPE;
typedef uint16_t TILE_MASK_TYPE;

#ifndef SMATRIX
#define SMATRIX
typedef struct
{
    int m;
    int n;
    int nnz;
    int isSymmetric;
    MAT_VAL_TYPE *value;            /* CSR values        [nnz]   */
    int *columnindex;               /* CSR column index  [nnz]   */
    MAT_PTR_TYPE *rowpointer;       /* CSR row pointer   [m+1]   */
    int tilem;                      /* ceil(m/16)                */
    int tilen;                      /* ceil(n/16)                */
    MAT_PTR_TYPE *tile_ptr;         /* [tilem+1] row-major tile structure */
    int *tile_columnidx;            /* [numtile] ascending per tile-row   */
    int *tile_rowidx;               /* [numtile]                          */
    int *tile_nnz;                  /* [numtile+1] exclusive offsets in storage order */
    int numtile;
    MAT_VAL_TYPE *tile_csr_Value;   /* [nnz]                              */
    TILE_CSR_COL_TYPE *tile_csr_Col;/* [nnz] A: row*16+col; B, C: col     */
    TILE_CSR_PTR_TYPE *tile_csr_Ptr;/* [numtile*16] per-tile exclusive row offsets (no 17th slot) */
    TILE_MASK_TYPE *mask;           /* [numtile*16] row r, col c <-> bit (15-c) */
    int *csc_tile_ptr;              /* [tilen+1]  B only: tile columns    */
    int *csc_tile_rowidx;           /* [numtile]  B only                  */
} SMatrix;
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1 -- drop-in entry points (host buffers in, host buffers out; malloc()-owned results so
 * the driver's free()/matrix_destroy() keep working). tile_size_m / tile_size_n: 16 x 16 runs the tuned kernels;
 * any other pair of multiples of 16 up to 128 runs the general-tile path (Part 3: tiles of A are tile_size_m x
 * tile_size_n, of B tile_size_n x tile_size_m, of C tile_size_m x tile_size_m, as in src/main.cu:84-91); anything
 * else sets TSG_ERR_UNSUPPORTED.
 * Input contract: CSR rows sorted by column and duplicate-free. csr2tile_row_major / csr2tile_col_major accept rows
 * that are not (they sort and merge on the device first, see tsg_csr_canonicalize); the tsg_* API reports TSG_ERR_INPUT.
 * ---------------------------------------------------------------------------------------- */

/* Replaces reference src/csr2tile.h:205. Fills tilem, tilen, numtile, tile_ptr, tile_columnidx,
 * tile_rowidx, tile_nnz, tile_csr_Value, tile_csr_Col (= row*16+col, :192), tile_csr_Ptr, mask. */
void csr2tile_row_major(SMatrix *matrix, int tile_size_m, int tile_size_n);

/* Replaces reference src/csr2tile.h:279. As above but tile data in CSC-tile order with
 * csc_tile_ptr / csc_tile_rowidx, tile_csr_Col = col (:475); tile_ptr / tile_columnidx keep the
 * row-major tile structure (:331-388); tile_rowidx is allocated and zero (:336-337). */
void csr2tile_col_major(SMatrix *matrix, int tile_size_m, int tile_size_n);

/* Replaces reference src/tilespgemm-cuda.h:2220-2235 (same argument list). Steps 1-3 on the
 * device. The dense tile bitmaps (src/main.cu:195-232) are accepted and ignored; pass NULL/0.
 * Fills C->{m,n,tilem,tilen,numtile,nnz,tile_ptr,tile_columnidx,tile_nnz,tile_csr_Value,
 * tile_csr_Col,tile_csr_Ptr} as the reference does (:2750-2775) plus C->mask and C->tile_rowidx.
 * Empty C tiles are kept with Ptr = mask = 0. A->mask / B->mask may be NULL
 * (tiles built elsewhere): the row masks are then rebuilt on the device from Ptr / Col. *gflops_tile =
 * 2*nnzCub/(ms*1e6), *compression_rate = nnzCub/nnzC, times in ms (:2804-2808). */
void tilespgemm(SMatrix *matrixA, SMatrix *matrixB, SMatrix *matrixC,
                unsigned int *blk_intersec_bitmask_A, unsigned int *blk_intersec_bitmask_B,
                int blk_intersec_bitmask_len, double densityA, double densityB,
                unsigned long long int nnzCub, unsigned long long int *nnzC_computed,
                double *compression_rate, double *time_tile, double *gflops_tile, char *filename,
                double *time_step1, double *time_step2, double *time_step3, double *time_malloc,
                int tile_size_m, int tile_size_n);

/* Replaces reference src/tile2csr.h:72. Fills rowpointer, columnindex, value, nnz from the tile
 * arrays (explicit zeros kept, columns ascending per row). */
void tile2csr(SMatrix *matrix, int tile_size_m, int tile_size_n);

/* Replaces reference src/csr2tile.h:509 (frees the same seven arrays). */
void matrix_destroy(SMatrix *matrix);

/* Replaces reference src/utils.h:161 (CSR -> CSC, stable; caller allocates the outputs). The
 * driver uses it to materialise B = A^T for -aat 1 (src/main.cu:130-139). Runs on the device. */
void matrix_transposition(const int m, const int n, const MAT_PTR_TYPE nnz,
                          const MAT_PTR_TYPE *csrRowPtr, const int *csrColIdx, const MAT_VAL_TYPE *csrVal,
                          int *cscRowIdx, MAT_PTR_TYPE *cscColPtr, MAT_VAL_TYPE *cscVal);

/* The reference functions are void and ignore failures; these report them. 0 = no error. The
 * error is sticky until tilespgemm_clear_error(). A driver should exit non-zero on error. */
enum {
    TSG_OK = 0,
    TSG_ERR_CUDA = 1,          /* a CUDA runtime call or kernel failed / no device            */
    TSG_ERR_UNSUPPORTED = 2,   /* tile size not a multiple of 16 in [16,128], or a shape the path cannot take */
    TSG_ERR_OVERFLOW = 3,      /* a 32-bit size of SMatrix would overflow (use the slab API)   */
    TSG_ERR_INPUT = 4,         /* CSR rows not sorted / duplicate columns / index out of range */
    TSG_ERR_NOMEM = 5
};
int tilespgemm_last_error(void);
const char *tilespgemm_last_error_string(void);
void tilespgemm_clear_error(void);

/* ------------------------------------------------------------------------------------------
 * Part 2 -- device-resident API. All pointers inside tsg_dcsr / tsg_dtile are DEVICE pointers on
 * the current CUDA device; the structs themselves live on the host. Work is enqueued on one
 * library-owned non-blocking stream per device; every function returns after its results are
 * complete unless stated otherwise. Return value: TSG_OK or an error code (also latched).
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    int m, n;
    long long nnz;
    int *rowptr;          /* [m+1]  */
    int *colidx;          /* [nnz]  */
    double *val;          /* [nnz]  */
    void *owner;          /* allocation to release in tsg_csr_free (NULL = borrowed pointers) */
} tsg_dcsr;

typedef struct {
    int m, n, tilem, tilen, numtile;
    int col_major;        /* 1: B layout (CSC-tile order, Col = col); 0: A / C layout           */
    int trow0;            /* first tile-row this object covers (slabs of C; 0 otherwise)        */
    int cached;           /* 1: slab[0..1] are C slab buffers that tsg_tile_free hands back to the library's cache */
    long long nnz;
    int *tile_ptr;        /* [tilem+1]                                                         */
    int *tile_columnidx;  /* [numtile]                                                         */
    int *tile_rowidx;     /* [numtile] global tile-row of each tile (row-major order)          */
    int *tile_nnz;        /* [numtile+1]                                                       */
    double *val;          /* [nnz]                                                             */
    uint16_t *col;        /* [nnz]                                                             */
    uint16_t *ptr;        /* [numtile*16]                                                      */
    uint16_t *mask;       /* [numtile*16]                                                      */
    int *csc_tile_ptr;    /* [tilen+1]  col_major only                                         */
    int *csc_tile_rowidx; /* [numtile]  col_major only                                         */
    int *rm2csc;          /* [numtile]  col_major only: row-major tile index -> storage id     */
    int *pat;             /* [numtile]  pattern id of every tile, in ROW-MAJOR tile order (storage order for a row-major
                             matrix, the order of rm2csc for a column-major one): tiles with equal ids have identical
                             row masks, i.e. the same sparsity pattern. Filled by csr2tile / tile upload. */
    int npat;             /* number of distinct patterns; -1: not available (too many -- the recipe plans of
                             csrc/plans.cu are then not attempted for this matrix) */
    void *slab[4];        /* device allocations owned by this object (tsg_tile_free)           */
    size_t slab_bytes[4];
} tsg_dtile;

typedef struct {
    long long numblkC;        /* C tiles listed (incl. empty ones)                   */
    long long nnzC;
    long long pairs;          /* matched (A tile, B tile) pairs = sum of step-1 weights */
    double ms_step1, ms_step2, ms_step3, ms_alloc, ms_total; /* device time, CUDA events  */
    long long algorithmic_bytes; /* bytes(A)+bytes(B)+bytes(C written)  */
    int launches;             /* kernels launched by this call                       */
    /* step 3 picks the accumulator per C tile / tile-row (csrc/numeric.cu): */
    int rows_staged;          /* C tile-rows computed by k_step3_rows (sparse accumulator in shared memory) */
    int rows_gather;          /* C tile-rows computed by k_step3_gather (lane per nonzero)                  */
    int tiles_dense;          /* C tiles computed by k_step3_dense (dense accumulator in registers)         */
    int rows_smem;            /* dynamic shared memory of k_step3_rows, bytes                               */
    int tiles_nonempty;       /* C tiles holding at least one entry (numblkC counts the empty ones too)     */
    int plan_recipes;         /* > 0: steps 2 and 3 ran from recipe plans (csrc/plans.cu), this many distinct recipes;
                                 0: generic kernels; -1: plans were attempted and fell back                     */
    int row_templates;        /* > 0: step 1 ran on this many representative tile-rows and the others were instantiated
                                 from them (csrc/rowplans.cu); 0: every tile-row expanded; -1: attempted, redone without */
} tsg_stats;

/* Select the device (like the driver's cudaSetDevice, reference src/main.cu:49) and create the
 * library context (stream, stream-ordered memory pool, scan workspace). Idempotent per device. */
int tsg_init(int device);
void tsg_shutdown(void);
/* The library's stream as a cudaStream_t (void* to keep this header free of CUDA includes). */
void *tsg_stream(void);
int tsg_sync(void);
/* Number of kernels this library has launched since tsg_init (for bench.py's gpu_launches). */
long long tsg_launch_count(void);
/* CUDA-event stopwatch on the library stream (torch.cuda.Event only sees torch's streams):
 * start records an event; stop records a second one, waits for it and returns the milliseconds
 * the stream spent between the two, host gaps included. */
int tsg_timer_start(void);
int tsg_timer_stop(double *ms);

/* CSR on the device. upload copies from host (pinned or pageable) memory; wrap borrows device
 * pointers the caller owns (e.g. a torch tensor's data_ptr()). */
int tsg_csr_upload(int m, int n, const int *rowptr, const int *colidx, const double *val, tsg_dcsr *out);
int tsg_csr_wrap(int m, int n, long long nnz, int *d_rowptr, int *d_colidx, double *d_val, tsg_dcsr *out);
int tsg_csr_download(const tsg_dcsr *a, int *rowptr, int *colidx, double *val);
void tsg_csr_free(tsg_dcsr *a);
/* Checks the input contract (sorted, duplicate-free, in-range columns). upload and wrap already check the index
 * ranges (column in [0,n), row pointer monotone and ending at nnz: TSG_ERR_INPUT) before anything indexes by them. */
int tsg_csr_validate(const tsg_dcsr *a);
/* Rows [row0, row1) of a device CSR as a CSR of its own: the row pointer is a rebased copy, colidx / val are BORROWED
 * from `a`, which must outlive the slice (free the slice with tsg_csr_free). No host round trip: this is how a rank of
 * the multi-GPU path takes its tile-rows out of the broadcast matrix. */
int tsg_csr_row_slice(const tsg_dcsr *a, int row0, int row1, tsg_dcsr *out);
/* Canonical form of a CSR whose rows are unsorted and / or hold duplicate columns -- what the reference's MatrixMarket
 * loader produces (src/mmio_highlevel.h:593-759: neither sorted nor merged): entries ordered by (row, column), duplicates
 * merged (dup_policy 0: summed in their original order; 1: first one kept). The drop-in csr2tile_* entry points do this
 * by themselves when they meet such input; the device API is strict (TSG_ERR_INPUT) and leaves the choice to the caller. */
int tsg_csr_canonicalize(const tsg_dcsr *a, int dup_policy, tsg_dcsr *out);

/* Device matrix_transposition (reference src/utils.h:161): AT = A^T as CSR, stable. */
int tsg_transpose(const tsg_dcsr *a, tsg_dcsr *at);
/* nnzCub = sum over entries (i,k) of A of nnz(B row k) (reference src/main.cu:155-160). */
int tsg_nnzcub(const tsg_dcsr *a, const tsg_dcsr *b, unsigned long long *out);

/* CSR -> tiles on the device (col_major = 0: csr2tile_row_major, 1: csr2tile_col_major). */
int tsg_csr2tile(const tsg_dcsr *a, int col_major, tsg_dtile *out);
/* Host SMatrix tile arrays -> device, and back (download malloc()s the host arrays). */
int tsg_tile_upload(const SMatrix *host, int col_major, tsg_dtile *out);
int tsg_tile_download(const tsg_dtile *t, SMatrix *host);
/* Allocate an empty tiled matrix of known sizes as ONE contiguous device slab (slab[0]) whose
 * layout is a pure function of the sizes: the multi-GPU path broadcasts B as that single buffer. */
int tsg_tile_alloc(int m, int n, int numtile, long long nnz, int col_major, tsg_dtile *out);
void tsg_tile_free(tsg_dtile *t);

/* Per-tile-row step-1 weights w[I] = number of matched tile pairs of C tile-row I (= sum over A
 * tiles (I,K) of |B tile-row K|), written to a HOST array of A->tilem entries. The multi-GPU
 * partitioner and the slab planner cut on its prefix sums. */
int tsg_tilerow_weights(const tsg_dtile *a, const tsg_dtile *b, long long *w_host);

/* Limits of one call / slab (TSG_ERR_OVERFLOW, "use smaller slabs"): < 2^31 matched tile pairs, < 2^30 C tiles, < 2^31 C
 * nonzeros. Every C entry is summed in a fixed order (pair lists ascending by A tile, also those of hub tile-rows, which are
 * collected with atomics and sorted afterwards): values are reproducible run to run.
 * SpGEMM steps 1-3 for C tile-rows [trow0, trow1) (trow1 < 0: to the end). C is a slab: a tiled
 * matrix of its own with tilem = trow1-trow0 and m = its row count. stats may be NULL. */
int tsg_spgemm(const tsg_dtile *a, const tsg_dtile *b, int trow0, int trow1, tsg_dtile *c, tsg_stats *stats);

/* Steps 1-3 over C tile-rows [trow0, trow1) in SLABS of at most max_pairs matched tile pairs each (<= 0: 2^28; a single
 * tile-row heavier than that is a slab of its own): what C = A*B needs when nnz(C) or its tile count does not fit the
 * 32-bit sizes of SMatrix or one GPU's memory (R-MAT scale 20 has nnz(C) ~ 1e10). Every slab -- a
 * tiled matrix of its own, trow0 = its first tile-row -- is handed to sink(slab, stats, user) and freed afterwards; a
 * non-zero return stops the run. totals (may be NULL) sums the slabs' stats in 64 bits; *nslabs_out = slabs run. */
typedef int (*tsg_slab_sink)(const tsg_dtile *c_slab, const tsg_stats *stats, void *user);
int tsg_spgemm_slabs(const tsg_dtile *a, const tsg_dtile *b, int trow0, int trow1, long long max_pairs, tsg_slab_sink sink,
                     void *user, tsg_stats *totals, int *nslabs_out);

/* tiles -> CSR on the device (reference src/tile2csr.h:72). */
int tsg_tile2csr(const tsg_dtile *t, tsg_dcsr *out);

/* Verification helper for slabs too large to download: per-row sum of the stored values (= C * ones)
 * and per-row entry count, written to HOST arrays of t->m entries (either may be NULL). */
int tsg_tile_rowsums(const tsg_dtile *t, double *sums_host, long long *counts_host);

/* Whole pipeline with HOST buffers, the end-to-end call bench.py times: H2D CSR(A) [and CSR(B);
 * b_* = NULL means B = A, aat != 0 means B = A^T built on the device], csr2tile x2, steps 1-3,
 * tile2csr, D2H CSR(C). Outputs are malloc()ed; 64-bit row pointers are never needed because the
 * call fails with TSG_ERR_OVERFLOW when nnz(C) >= 2^31. */
int tsg_spgemm_csr_host(int m, int k, int n, const int *a_rowptr, const int *a_colidx, const double *a_val,
                        const int *b_rowptr, const int *b_colidx, const double *b_val, int aat,
                        int **c_rowptr, int **c_colidx, double **c_val, long long *c_nnz, tsg_stats *stats);

/* Steps 1-3 + tile2csr + D2H for C tile-rows [trow0, trow1) of A*B, overlapped: the range is cut into
 * nslabs slabs of about equal step-1 weight (nslabs <= 0: up to 16, each >= 2^20 tile pairs and <= 2^28),
 * and the CSR of slab s is copied to the host on a second stream while slab s+1 is computed (the path is
 * PCIe-bound: 12 bytes per C nonzero leave the device). Output goes to CALLER buffers -- page-locked
 * memory for the copies to overlap: c_rowptr holds rows+1 entries (rows = rows of the range, numbered
 * from 0), c_colidx / c_val hold c_cap entries. *c_nnz = nnz(C) of the range. If c_cap is too small the
 * call still counts, sets *c_nnz to the capacity needed and returns TSG_ERR_NOMEM; TSG_ERR_OVERFLOW when
 * nnz(C) >= 2^31 (the reference's CSR, src/tile2csr.h:72, has 32-bit row pointers too). stats (may be
 * NULL) are summed over the slabs. */
int tsg_spgemm_to_host(const tsg_dtile *a, const tsg_dtile *b, int trow0, int trow1, int nslabs, int *c_rowptr,
                       int *c_colidx, double *c_val, long long c_cap, long long *c_nnz, tsg_stats *stats);

/* The slab plan tsg_spgemm_to_host uses (pure host code, no device needed): boundaries of the slabs of
 * tile-rows [trow0, trow1) given the step-1 weights (tsg_tilerow_weights; indexed by absolute tile-row).
 * Writes nslabs+1 ascending boundaries cuts[0] = trow0 .. cuts[nslabs] = trow1 and returns the number of
 * slabs (0 for an empty range), or -1 (error latched) if cuts_cap is too small (trow1-trow0+1 always fits). */
int tsg_plan_slabs(const long long *weights, int trow0, int trow1, int nslabs, int *cuts, int cuts_cap);

/* tsg_spgemm_csr_host with caller-provided output buffers and the overlap above: H2D CSR(A) [CSR(B)],
 * csr2tile x2, then tsg_spgemm_to_host over all tile-rows. */
int tsg_spgemm_csr_host_into(int m, int k, int n, const int *a_rowptr, const int *a_colidx, const double *a_val,
                             const int *b_rowptr, const int *b_colidx, const double *b_val, int aat, int *c_rowptr,
                             int *c_colidx, double *c_val, long long c_cap, long long *c_nnz, tsg_stats *stats);

/* ------------------------------------------------------------------------------------------
 * Part 3 -- general tile sizes (the feature this fork adds: runtime tile_size_m x tile_size_n,
 * reference src/main.cu:84-91, src/common.h:138-146, src/csr2tile.h:192-195,255,472-474, src/tilespgemm-cuda.h:495-705).
 * With the driver's (tile_size_m, tile_size_n): a tile of A is tile_size_m x tile_size_n, a tile of B tile_size_n x
 * tile_size_m, a tile of C tile_size_m x tile_size_m. Both must be multiples of 16 (MaskBits) between 16 and 128.
 * Format per tile of R rows x Q columns: Ptr R slots, mask R rows of Q/16 words (column c <-> word c/16, bit 15 - c%16),
 * Col = r*Q + c for A and c for B and C. Everything else is as for 16 x 16. The drop-in entry points of Part 1 take any
 * such size and route 16 x 16 to the tuned kernels and every other size here (csrc/gentile.cu); at 16 x 16 this path
 * produces bit-identical arrays. Sizes of one call: < 2^31 tile pairs, < 2^31 nonzeros of C.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int m, n;
    int tile_rows, tile_cols;   /* rows x columns of ONE tile of this matrix                              */
    int tilem, tilen, numtile;
    int col_major;              /* 1: B layout (CSC-tile order, Col = col); 0: A / C layout               */
    long long nnz;
    int *tile_ptr;              /* [tilem+1]                                                              */
    int *tile_columnidx;        /* [numtile]                                                              */
    int *tile_rowidx;           /* [numtile] tile row of each tile, row-major order                       */
    int *tile_nnz;              /* [numtile+1] storage order                                              */
    double *val;                /* [nnz]                                                                  */
    uint16_t *col;              /* [nnz]                                                                  */
    uint16_t *ptr;              /* [numtile*tile_rows]                                                    */
    uint16_t *mask;             /* [numtile*tile_rows*(tile_cols/16)]                                     */
    int *csc_tile_ptr;          /* [tilen+1]  col_major only                                              */
    int *csc_tile_rowidx;       /* [numtile]  col_major only                                              */
    int *rm2csc;                /* [numtile]  col_major only: row-major tile index -> storage id          */
    void *slab[2];              /* device allocations owned by this object (tsg_gtile_free)               */
    size_t slab_bytes[2];
} tsg_gtile;

/* 1 if tiles of tile_rows x tile_cols are representable (multiples of 16 in [16, 128]). */
int tsg_gtile_size_ok(int tile_rows, int tile_cols);
/* CSR -> tiles of tile_rows x tile_cols on the device (col_major = 0: csr2tile_row_major, Col = r*tile_cols + c;
 * 1: csr2tile_col_major, CSC-tile order, Col = c). For the reference's csr2tile_col_major(B, tile_size_m, tile_size_n)
 * pass tile_rows = tile_size_n, tile_cols = tile_size_m. */
int tsg_gtile_csr2tile(const tsg_dcsr *a, int col_major, int tile_rows, int tile_cols, tsg_gtile *out);
/* Host SMatrix tile arrays of that tile size -> device, and back (download malloc()s the host arrays). */
int tsg_gtile_upload(const SMatrix *host, int col_major, int tile_rows, int tile_cols, tsg_gtile *out);
int tsg_gtile_download(const tsg_gtile *t, SMatrix *host);
void tsg_gtile_free(tsg_gtile *t);
/* Steps 1-3: C = A*B with A row-major (R x Q tiles), B column-major (Q x P tiles); C comes out row-major with R x P
 * tiles, empty tiles kept. stats (may be NULL): numblkC, nnzC, pairs, step times, launches, algorithmic_bytes. */
int tsg_gtile_spgemm(const tsg_gtile *a, const tsg_gtile *b, tsg_gtile *c, tsg_stats *stats);
/* tiles -> CSR on the device (row-major storage, Col = c: what tsg_gtile_spgemm returns). */
int tsg_gtile_tile2csr(const tsg_gtile *t, tsg_dcsr *out);

#ifdef __cplusplus
}
#endif
#endif /* TILESPGEMM_B200_H */
