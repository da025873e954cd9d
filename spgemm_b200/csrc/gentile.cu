// gentile.cu -- the whole path for GENERAL tile sizes: csr2tile, SpGEMM steps 1-3 and tile2csr for tiles of
// tile_size_m x tile_size_n (A), tile_size_n x tile_size_m (B) and tile_size_m x tile_size_m (C), both multiples of 16
// up to 128 -- the feature this fork adds to TileSpGEMM:
//   runtime tile sizes        src/main.cu:84-91 ("the tile of A is m x n, and the tile of B is n x m")
//   multi-word row masks      src/common.h:138-146 (MaskBits = 16), src/csr2tile.h:192-195,255,472-474
//   per-row MaskNum loops     src/tilespgemm-cuda.h:495-705
//   tile2csr(C, m, m)         src/main.cu:327, src/tile2csr.h:72-140
// The 16 x 16 case keeps its own, tuned kernels (csr2tile.cu, spgemm.cu, numeric.cu, plans.cu, tile2csr.cu); this file
// is what every other size runs, and at 16 x 16 it produces bit-identical arrays (tests/test_gentile_gpu.py).
//
// Design. Every kernel here is ONE THREAD PER ITEM (a CSR entry, a run, a tile pair, a (tile,row), a C nonzero) with no
// communication between threads other than integer atomics whose result does not depend on the order; everything that
// needs an order comes from the library's stable radix sort (radix_sort.cuh) and its look-back scan (scan.cuh):
//   csr2tile   the maximal runs of a CSR row inside one tile column are found per entry, sorted stably by (tile row, tile
//              column) -- which is the tile list, the (tile,row) counts (-> Ptr), the row masks and, per run, the copy
//              of its entries; the CSC-tile order of B is one more stable sort of the tile list by tile column.
//   step 1     every (A tile, B tile) pair is expanded once (its position by a scan of |B tile-row K| over A's tiles),
//              sorted stably by (tile row, tile column): runs of equal keys are C's tiles, and inside a run the pairs are
//              in ascending A-tile order -- the serial SPA's summation order. No bitmaps, no heavy / light split.
//   step 2     thread per (C tile, row): OR of B's row masks over the row's A entries, W = tile_size_m/16 words.
//   step 3     thread per C nonzero: the products of its (row, column) in pair order, fma() like the oracle.
//   tile2csr   thread per matrix row, twice (count, fill).
// Being free of warp-level code the file also compiles as plain C++ (GT_EMULATE, tests/emu/): the CPU tests run every
// kernel and all of the host orchestration below against the oracle, thread by thread, before a GPU ever sees it.
#ifdef GT_EMULATE
#include "gentile_emu.h"
#else
#include "common.cuh"
#include "scan.cuh"
#include "kernels.h"
#define GT_KERNEL __global__ void
#define GT_DEVICE __device__ __forceinline__
#define GT_TID ((long long)blockIdx.x * blockDim.x + threadIdx.x)
#define GT_POPC(x) __popc(x)
// thread per item: n items, 256 threads per CTA
#define GT_LAUNCH(kern, n, ...)                                                            \
    do {                                                                                   \
        long long n_ = (long long)(n);                                                     \
        if (n_ > 0) {                                                                      \
            kern<<<(unsigned)((n_ + 255) / 256), 256, 0, tsg::ctx().stream>>>(__VA_ARGS__); \
            CK_LAUNCH();                                                                   \
        }                                                                                  \
    } while (0)
#endif

namespace tsg {

// ------------------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------------------

// largest i in [0, n) with a[i] <= key (a ascending, a[0] <= key): the owner of position `key` in an offsets array;
// among equal offsets (empty owners) the LAST one, which is the one that is not empty
GT_DEVICE int gt_owner(const int *__restrict__ a, int n, long long key)
{
    int lo = 0, hi = n;  // invariant: a[lo] <= key, a[hi] > key (hi = n: virtual +inf)
    while (hi - lo > 1) {
        int mid = (int)(((long long)lo + hi) >> 1);
        if ((long long)a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

GT_DEVICE int gt_owner_u16(const uint16_t *__restrict__ a, int n, int key)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((int)a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------------------------
// csr2tile, general tile size (reference src/csr2tile.h:205-277 and :279-506)
// ------------------------------------------------------------------------------------------------------------------

// A run = the maximal stretch of one CSR row inside one tile column. flag[p] = 1 iff entry p starts a run.
// Also the input contract (columns ascending and distinct inside a row): *err |= 2.
GT_KERNEL k_g_run_flags(long long nnz, int m, int TC, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                        int *__restrict__ flag, int *__restrict__ err)
{
    const long long p = GT_TID;
    if (p >= nnz) return;
    const int i = gt_owner(rowptr, m + 1, p);
    const int c = colidx[p];
    int head = 1;
    if (p > rowptr[i]) {
        const int cp = colidx[p - 1];
        if (cp >= c) atomicOr(err, 2);
        head = (cp / TC) != (c / TC);
    }
    flag[p] = head;
}

// per run (in CSR order): first entry, matrix row, tile column
GT_KERNEL k_g_run_emit(long long nnz, int m, int TC, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                       const int *__restrict__ flag, const int *__restrict__ runidx, int *__restrict__ run_start,
                       int *__restrict__ run_row, uint32_t *__restrict__ keyJ, uint32_t *__restrict__ ident)
{
    const long long p = GT_TID;
    if (p >= nnz) return;
    if (!flag[p]) return;
    const int r = runidx[p];
    run_start[r] = (int)p;
    run_row[r] = gt_owner(rowptr, m + 1, p);
    keyJ[r] = (uint32_t)(colidx[p] / TC);
    ident[r] = (uint32_t)r;
}

// key of the second sort: tile row of the run at sorted position q
GT_KERNEL k_g_gather_tilerow(long long n, int TR, const uint32_t *__restrict__ perm, const int *__restrict__ run_row,
                             uint32_t *__restrict__ keyI)
{
    const long long q = GT_TID;
    if (q >= n) return;
    keyI[q] = (uint32_t)(run_row[perm[q]] / TR);
}

// runs sorted by (tile row, tile column, matrix row): head[q] = 1 iff the run at q opens a new tile
GT_KERNEL k_g_tile_heads(long long n, int TR, int TC, const uint32_t *__restrict__ perm, const int *__restrict__ run_row,
                         const int *__restrict__ run_start, const int *__restrict__ colidx, int *__restrict__ head)
{
    const long long q = GT_TID;
    if (q >= n) return;
    int h = 1;
    if (q > 0) {
        const int r1 = (int)perm[q], r0 = (int)perm[q - 1];
        h = (run_row[r1] / TR != run_row[r0] / TR) || (colidx[run_start[r1]] / TC != colidx[run_start[r0]] / TC);
    }
    head[q] = h;
}

// tile list in row-major order; tile_of_run[r] = row-major tile index of run r
GT_KERNEL k_g_tile_struct(long long n, int TR, int TC, const uint32_t *__restrict__ perm, const int *__restrict__ run_row,
                          const int *__restrict__ run_start, const int *__restrict__ colidx, const int *__restrict__ head,
                          const int *__restrict__ headscan, int *__restrict__ tile_col, int *__restrict__ tile_row,
                          int *__restrict__ tile_of_run)
{
    const long long q = GT_TID;
    if (q >= n) return;
    const int r = (int)perm[q];
    const int t = headscan[q] + head[q] - 1;  // headscan = exclusive scan of head
    tile_of_run[r] = t;
    if (head[q]) {
        tile_col[t] = colidx[run_start[r]] / TC;
        tile_row[t] = run_row[r] / TR;
    }
}

// ptr_out[j] = first position q of `sorted` (ascending, n entries) with sorted[q] >= j, for j in [0, domain]
GT_KERNEL k_g_boundaries(long long n, int domain, const int *__restrict__ sorted, int *__restrict__ ptr_out)
{
    const long long q = GT_TID;
    if (q > n) return;
    const long long prev = q == 0 ? -1 : (long long)sorted[q - 1];
    const long long cur = q == n ? (long long)domain : (long long)sorted[q];
    for (long long j = prev + 1; j <= cur; j++) ptr_out[j] = (int)q;
}

GT_KERNEL k_g_copy_u32(long long n, const int *__restrict__ in, uint32_t *__restrict__ keys, uint32_t *__restrict__ ident)
{
    const long long i = GT_TID;
    if (i >= n) return;
    keys[i] = (uint32_t)in[i];
    ident[i] = (uint32_t)i;
}

// CSC-tile order: perm[d] = row-major index of the d-th stored tile
GT_KERNEL k_g_csc_order(long long numtile, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ sorted_cols,
                        const int *__restrict__ tile_row, int *__restrict__ csc_rowidx, int *__restrict__ rm2csc,
                        int *__restrict__ sorted_cols_i)
{
    const long long d = GT_TID;
    if (d >= numtile) return;
    const int t = (int)perm[d];
    csc_rowidx[d] = tile_row[t];
    rm2csc[t] = (int)d;
    sorted_cols_i[d] = (int)sorted_cols[d];
}

// per run: its (tile,row) count, the row's mask words, and its share of the tile's entry count.
// newid: row-major tile index -> storage id (nullptr: identity). ptr/mask/tile_cnt are zero before this kernel.
GT_KERNEL k_g_run_counts(long long nruns, long long nnz, int m, int TR, int TC, const int *__restrict__ rowptr,
                         const int *__restrict__ colidx, const int *__restrict__ run_start, const int *__restrict__ run_row,
                         const int *__restrict__ tile_of_run, const int *__restrict__ newid, uint16_t *__restrict__ ptr,
                         uint16_t *__restrict__ mask, int *__restrict__ tile_cnt)
{
    const long long r = GT_TID;
    if (r >= nruns) return;
    const int i = run_row[r], p0 = run_start[r];
    int p1 = rowptr[i + 1];
    if (r + 1 < nruns && run_row[r + 1] == i) p1 = run_start[r + 1];
    const int t = tile_of_run[r];
    const size_t sid = (size_t)(newid ? newid[t] : t);
    const int lr = i % TR, W = TC >> 4;
    const int c0 = (colidx[p0] / TC) * TC;
    ptr[sid * TR + lr] = (uint16_t)(p1 - p0);
    atomicAdd(&tile_cnt[sid], p1 - p0);
    uint16_t *mrow = mask + (sid * TR + lr) * W;
    int wcur = -1;
    unsigned acc = 0;
    for (int p = p0; p < p1; p++) {
        const int c = colidx[p] - c0, w = c >> 4;
        if (w != wcur) {
            if (wcur >= 0) mrow[wcur] = (uint16_t)acc;
            wcur = w; acc = 0;
        }
        acc |= 0x8000u >> (c & 15);  // column c <-> word c/16, bit 15 - c%16 (src/csr2tile.h:194-195)
    }
    if (wcur >= 0) mrow[wcur] = (uint16_t)acc;
    (void)nnz; (void)m;
}

// per tile: counts -> exclusive offsets over the TR slots (rows past the matrix edge repeat the total)
GT_KERNEL k_g_ptr_scan(long long numtile, int TR, uint16_t *__restrict__ ptr)
{
    const long long t = GT_TID;
    if (t >= numtile) return;
    uint16_t *p = ptr + (size_t)t * TR;
    unsigned run = 0;
    for (int r = 0; r < TR; r++) {
        const unsigned c = p[r];
        p[r] = (uint16_t)run;
        run += c;
    }
}

// per run: copy its entries to the tile. PACKED: Col = r*TC + c (A, src/csr2tile.h:192), else Col = c (B, :475)
GT_KERNEL k_g_run_scatter(long long nruns, int TR, int TC, int packed, const int *__restrict__ rowptr,
                          const int *__restrict__ colidx, const double *__restrict__ val, const int *__restrict__ run_start,
                          const int *__restrict__ run_row, const int *__restrict__ tile_of_run, const int *__restrict__ newid,
                          const int *__restrict__ tile_nnz, const uint16_t *__restrict__ ptr, double *__restrict__ val_out,
                          uint16_t *__restrict__ col_out)
{
    const long long r = GT_TID;
    if (r >= nruns) return;
    const int i = run_row[r], p0 = run_start[r];
    int p1 = rowptr[i + 1];
    if (r + 1 < nruns && run_row[r + 1] == i) p1 = run_start[r + 1];
    const int t = tile_of_run[r];
    const size_t sid = (size_t)(newid ? newid[t] : t);
    const int lr = i % TR;
    const int c0 = (colidx[p0] / TC) * TC;
    size_t dst = (size_t)tile_nnz[sid] + ptr[sid * TR + lr];
    for (int p = p0; p < p1; p++, dst++) {
        const int c = colidx[p] - c0;
        val_out[dst] = val[p];
        col_out[dst] = (uint16_t)(packed ? lr * TC + c : c);
    }
}

// row masks from Ptr / Col (tiles uploaded without their mask array)
GT_KERNEL k_g_masks_from_tiles(long long nrows_total, int TR, int TC, int packed, const int *__restrict__ tile_nnz,
                               const uint16_t *__restrict__ ptr, const uint16_t *__restrict__ col, uint16_t *__restrict__ mask)
{
    const long long g = GT_TID;
    if (g >= nrows_total) return;
    const size_t t = (size_t)(g / TR);
    const int r = (int)(g % TR), W = TC >> 4;
    const int base = tile_nnz[t], tn = tile_nnz[t + 1] - base;
    const int s = ptr[t * TR + r], e = r + 1 < TR ? ptr[t * TR + r + 1] : tn;
    uint16_t *mrow = mask + (t * TR + r) * W;
    for (int w = 0; w < W; w++) mrow[w] = 0;
    for (int x = s; x < e; x++) {
        const int c = packed ? (int)col[base + x] - r * TC : (int)col[base + x];
        mrow[c >> 4] = (uint16_t)(mrow[c >> 4] | (0x8000u >> (c & 15)));
    }
}

// tile row of every tile of a row-major tile list (the reference leaves B's tile_rowidx zero, src/csr2tile.h:336-337)
GT_KERNEL k_g_tile_rows(long long numtile, int tilem, const int *__restrict__ tile_ptr, int *__restrict__ tile_row)
{
    const long long t = GT_TID;
    if (t >= numtile) return;
    tile_row[t] = gt_owner(tile_ptr, tilem + 1, t);
}

// rm2csc of an uploaded column-major matrix: the tile (I, J) is the entry I of tile column J
GT_KERNEL k_g_rm2csc(long long numtile, const int *__restrict__ tile_col, const int *__restrict__ tile_row,
                     const int *__restrict__ csc_ptr, const int *__restrict__ csc_rowidx, int *__restrict__ rm2csc,
                     int *__restrict__ err)
{
    const long long t = GT_TID;
    if (t >= numtile) return;
    const int J = tile_col[t], I = tile_row[t];
    int lo = csc_ptr[J], hi = csc_ptr[J + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (csc_rowidx[mid] < I) lo = mid + 1; else hi = mid;
    }
    if (lo >= csc_ptr[J + 1] || csc_rowidx[lo] != I) { atomicOr(err, 4); rm2csc[t] = 0; return; }
    rm2csc[t] = lo;
}

// ------------------------------------------------------------------------------------------------------------------
// step 1, general: C's tile list and the pair list of every C tile
// ------------------------------------------------------------------------------------------------------------------

// weight of an A tile (I,K): the number of tiles of B's tile-row K (reference: nsparse set_intprod_num,
// src/spgemm_nsparse_kernel.h:135-151)
GT_KERNEL k_g_pair_weights(long long ntA, const int *__restrict__ tile_colA, const int *__restrict__ tile_ptrB, int *__restrict__ w)
{
    const long long ta = GT_TID;
    if (ta >= ntA) return;
    const int K = tile_colA[ta];
    w[ta] = tile_ptrB[K + 1] - tile_ptrB[K];
}

// every (A tile, B tile) pair once, in (A tile asc, B tile asc) order
GT_KERNEL k_g_pair_expand(long long npairs, int ntA, const int *__restrict__ woff, const int *__restrict__ tile_colA,
                          const int *__restrict__ tile_ptrB, const int *__restrict__ tile_colB, const int *__restrict__ rm2cscB,
                          int *__restrict__ pa, int *__restrict__ pb, uint32_t *__restrict__ keyJ, uint32_t *__restrict__ ident)
{
    const long long p = GT_TID;
    if (p >= npairs) return;
    const int ta = gt_owner(woff, ntA + 1, p);
    const int K = tile_colA[ta];
    const int tb = tile_ptrB[K] + (int)(p - woff[ta]);
    pa[p] = ta;
    pb[p] = rm2cscB[tb];
    keyJ[p] = (uint32_t)tile_colB[tb];
    ident[p] = (uint32_t)p;
}

GT_KERNEL k_g_pair_tilerow(long long npairs, const uint32_t *__restrict__ perm, const int *__restrict__ pa,
                           const int *__restrict__ tile_rowA, uint32_t *__restrict__ keyI)
{
    const long long q = GT_TID;
    if (q >= npairs) return;
    keyI[q] = (uint32_t)tile_rowA[pa[perm[q]]];
}

// pairs sorted by (tile row, tile column): (keyI, keyJ2) are the sorted keys; head[q] = 1 iff q opens a C tile
GT_KERNEL k_g_pair_heads(long long npairs, const uint32_t *__restrict__ keyI, const uint32_t *__restrict__ perm,
                         const uint32_t *__restrict__ keyJ_orig, int *__restrict__ head)
{
    const long long q = GT_TID;
    if (q >= npairs) return;
    int h = 1;
    if (q > 0) h = keyI[q] != keyI[q - 1] || keyJ_orig[perm[q]] != keyJ_orig[perm[q - 1]];
    head[q] = h;
}

GT_KERNEL k_g_pair_emit(long long npairs, int trow0, const uint32_t *__restrict__ keyI, const uint32_t *__restrict__ perm,
                        const uint32_t *__restrict__ keyJ_orig, const int *__restrict__ pa, const int *__restrict__ pb,
                        const int *__restrict__ head, const int *__restrict__ headscan, int *__restrict__ c_tile_col,
                        int *__restrict__ c_tile_row, int *__restrict__ pair_ptr, int *__restrict__ pair_a, int *__restrict__ pair_b)
{
    const long long q = GT_TID;
    if (q >= npairs) return;
    const uint32_t src = perm[q];
    pair_a[q] = pa[src];
    pair_b[q] = pb[src];
    if (head[q]) {
        const int t = headscan[q];
        c_tile_col[t] = (int)keyJ_orig[src];
        c_tile_row[t] = (int)keyI[q] - trow0;
        pair_ptr[t] = (int)q;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// step 2, general: row masks, per-row counts of every C tile (reference K4, src/tilespgemm-cuda.h:394-773; the
// per-row MaskNum loops are :495-705)
// ------------------------------------------------------------------------------------------------------------------
#define GT_MAXW 8  // mask words per row: tile columns <= 128

GT_KERNEL k_g_symbolic(long long nrows_total, int TRc, int TCa, int Wc, const int *__restrict__ pair_ptr,
                       const int *__restrict__ pair_a, const int *__restrict__ pair_b, const int *__restrict__ tile_nnzA,
                       const uint16_t *__restrict__ ptrA, const uint16_t *__restrict__ colA, const uint16_t *__restrict__ maskB,
                       uint16_t *__restrict__ ptrC, uint16_t *__restrict__ maskC, int *__restrict__ tile_cnt)
{
    const long long g = GT_TID;
    if (g >= nrows_total) return;
    const size_t t = (size_t)(g / TRc);
    const int r = (int)(g % TRc);
    unsigned acc[GT_MAXW];
#pragma unroll
    for (int w = 0; w < GT_MAXW; w++) acc[w] = 0;
    const int TRb = TCa;  // rows of a B tile = columns of an A tile
    for (int q = pair_ptr[t]; q < pair_ptr[t + 1]; q++) {
        const size_t a = (size_t)pair_a[q], b = (size_t)pair_b[q];
        const int baseA = tile_nnzA[a], nA = tile_nnzA[a + 1] - baseA;
        const int s = ptrA[a * TRc + r], e = r + 1 < TRc ? ptrA[a * TRc + r + 1] : nA;
        for (int x = s; x < e; x++) {
            const int k = (int)colA[baseA + x] - r * TCa;  // A's Col = r*TCa + k (src/csr2tile.h:192)
            const uint16_t *mb = maskB + (b * TRb + k) * Wc;
#pragma unroll
            for (int w = 0; w < GT_MAXW; w++)
                if (w < Wc) acc[w] |= mb[w];
        }
    }
    int cnt = 0;
    uint16_t *mc = maskC + (t * TRc + r) * Wc;
#pragma unroll
    for (int w = 0; w < GT_MAXW; w++)
        if (w < Wc) { mc[w] = (uint16_t)acc[w]; cnt += GT_POPC(acc[w]); }
    ptrC[t * TRc + r] = (uint16_t)cnt;
    if (cnt) atomicAdd(&tile_cnt[t], cnt);
}

// ------------------------------------------------------------------------------------------------------------------
// step 3, general: one thread per C nonzero (reference K6/K7, src/tilespgemm-cuda.h:1273-1952). The products of a
// nonzero are added in pair order and, inside a pair, in the order of A's row: the serial SPA's order
// (src/external/cusparse/spgemm_serialref_spa.h:7-31), with fma() as the -mfma build of the reference contracts it.
// ------------------------------------------------------------------------------------------------------------------
GT_KERNEL k_g_numeric(long long nnzC, int numblkC, int TRc, int TCa, int Wc, const int *__restrict__ pair_ptr,
                      const int *__restrict__ pair_a, const int *__restrict__ pair_b, const int *__restrict__ tile_nnzA,
                      const uint16_t *__restrict__ ptrA, const uint16_t *__restrict__ colA, const double *__restrict__ valA,
                      const int *__restrict__ tile_nnzB, const uint16_t *__restrict__ ptrB, const uint16_t *__restrict__ maskB,
                      const double *__restrict__ valB, const int *__restrict__ tile_nnzC, const uint16_t *__restrict__ ptrC,
                      const uint16_t *__restrict__ maskC, uint16_t *__restrict__ colC, double *__restrict__ valC)
{
    const long long g = GT_TID;
    if (g >= nnzC) return;
    const size_t t = (size_t)gt_owner(tile_nnzC, numblkC + 1, g);
    const int x0 = (int)(g - tile_nnzC[t]);
    const int r = gt_owner_u16(ptrC + t * TRc, TRc, x0);
    int j = x0 - (int)ptrC[t * TRc + r];
    // the j-th set bit of the row mask, columns ascending (word 0 bit 15 = column 0)
    const uint16_t *mc = maskC + (t * TRc + r) * Wc;
    int c = -1;
    for (int w = 0; w < Wc && c < 0; w++) {
        const unsigned mw = mc[w];
        const int pc = GT_POPC(mw);
        if (j >= pc) { j -= pc; continue; }
        for (int bit = 15; bit >= 0; bit--)
            if ((mw >> bit) & 1u) {
                if (j == 0) { c = w * 16 + (15 - bit); break; }
                j--;
            }
    }
    const int cw = c >> 4, cbit = 15 - (c & 15);
    const int TRb = TCa;
    double sum = 0.0;
    for (int q = pair_ptr[t]; q < pair_ptr[t + 1]; q++) {
        const size_t a = (size_t)pair_a[q], b = (size_t)pair_b[q];
        const int baseA = tile_nnzA[a], nA = tile_nnzA[a + 1] - baseA;
        const int s = ptrA[a * TRc + r], e = r + 1 < TRc ? ptrA[a * TRc + r + 1] : nA;
        for (int x = s; x < e; x++) {
            const int k = (int)colA[baseA + x] - r * TCa;
            const uint16_t *mb = maskB + (b * TRb + k) * Wc;
            const unsigned mw = mb[cw];
            if (!((mw >> cbit) & 1u)) continue;
            int rank = GT_POPC(mw >> (cbit + 1));  // columns of this word before c
            for (int w = 0; w < cw; w++) rank += GT_POPC((unsigned)mb[w]);
            const size_t ib = (size_t)tile_nnzB[b] + ptrB[b * TRb + k] + rank;
            sum = fma(valA[baseA + x], valB[ib], sum);
        }
    }
    colC[g] = (uint16_t)c;
    valC[g] = sum;
}

#ifndef GT_EMULATE
// ------------------------------------------------------------------------------------------------------------------
// step 3 for 32-row C tiles that are well filled: dense accumulator in REGISTERS (the only warp-level kernel of this
// file; CUDA only -- the emulated build always takes k_g_numeric, and the GPU tests run both on the same inputs).
// A warp owns a C tile of 32 x 32, lane = row, 32 accumulators per lane. Per pair the B tile (TRb rows x 32 columns) is
// expanded to a dense tile in shared memory (row stride 33 doubles: lanes reading different rows hit different banks),
// then every lane walks the A entries (r, k) of its row and adds a * Bd[k][0..31] -- 32 LDS + 32 DFMA per A entry,
// no search, no mask test. Absent B entries contribute a * 0 = 0 to an accumulator, which leaves it unchanged, so every
// C entry is still the fma() chain of its products in pair order, then in the order of A's row: bit-identical to
// k_g_numeric and the oracle. Compaction through C's row mask at the end.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GD_WARPS = 4;
constexpr int GD_STRIDE = 33;
constexpr int GD_MAXTRB = 64;  // rows of a B tile this kernel takes (columns of an A tile)

__global__ void __launch_bounds__(GD_WARPS * 32, 4)
k_g_numeric_dense32(int numblkC, int TCa, const int *__restrict__ pair_ptr, const int *__restrict__ pair_a, const int *__restrict__ pair_b,
                    const int *__restrict__ tile_nnzA, const uint16_t *__restrict__ ptrA, const uint16_t *__restrict__ colA,
                    const double *__restrict__ valA, const int *__restrict__ tile_nnzB, const uint16_t *__restrict__ ptrB,
                    const uint16_t *__restrict__ colB, const double *__restrict__ valB, const int *__restrict__ tile_nnzC,
                    const uint16_t *__restrict__ ptrC, const uint16_t *__restrict__ maskC, uint16_t *__restrict__ colC,
                    double *__restrict__ valC)
{
    extern __shared__ double gd_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int TRb = TCa;
    double *Bd = gd_smem + (size_t)warp * (TRb * GD_STRIDE + GD_MAXTRB / 2);
    int *sp = (int *)(Bd + TRb * GD_STRIDE);  // row offsets of the B tile being expanded
    const int r = lane;
    // a warp walks C tiles with a stride of the grid's warps: 40 % of a block-FEM product's listed tiles are empty,
    // one warp per tile would leave the CTAs half idle
    for (int t = blockIdx.x * GD_WARPS + warp; t < numblkC; t += gridDim.x * GD_WARPS) {
        const int outbase = tile_nnzC[t];
        if (tile_nnzC[t + 1] == outbase) continue;  // empty tile
        double acc[32];
#pragma unroll
        for (int c = 0; c < 32; c++) acc[c] = 0.0;
        const int q0 = pair_ptr[t], q1 = pair_ptr[t + 1];
        // software pipeline over the pairs: the tile ids of pair q+1 are requested at the top of pair q, what depends on
        // them (offsets, row ranges) after B's expansion, so that both fly under the FMAs of pair q
        size_t a = (size_t)pair_a[q0], b = (size_t)pair_b[q0];
        int baseA = tile_nnzA[a], nA = tile_nnzA[a + 1] - baseA;
        int baseB = tile_nnzB[b], nB = tile_nnzB[b + 1] - baseB;
        int s = ptrA[a * 32 + r], e = r + 1 < 32 ? (int)ptrA[a * 32 + r + 1] : nA;
        int pk0 = lane < TRb ? (int)ptrB[b * TRb + lane] : 0x7fffffff;
        int pk1 = lane + 32 < TRb ? (int)ptrB[b * TRb + lane + 32] : 0x7fffffff;
        for (int q = q0; q < q1; q++) {
            const bool more = q + 1 < q1;
            size_t an = 0, bn = 0;
            if (more) { an = (size_t)pair_a[q + 1]; bn = (size_t)pair_b[q + 1]; }
            int kc = 0;
            double av = 0.0;
            if (s < e) { kc = colA[baseA + s]; av = valA[baseA + s]; }
            for (int i = lane; i < TRb * GD_STRIDE; i += 32) Bd[i] = 0.0;
            sp[lane] = pk0; sp[lane + 32] = pk1;
            __syncwarp();
            // B's entries, lane per entry (coalesced), two per lane in flight; the row of entry x = last k with sp[k] <= x
            for (int x0 = 0; x0 < nB; x0 += 64) {
                const int xa = x0 + lane, xb = x0 + 32 + lane;
                int ca = 0, cb = 0;
                double va = 0.0, vb = 0.0;
                if (xa < nB) { ca = colB[baseB + xa]; va = valB[baseB + xa]; }
                if (xb < nB) { cb = colB[baseB + xb]; vb = valB[baseB + xb]; }
                if (xa < nB) {
                    int lo = 0, hi = TRb;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sp[mid] <= xa) lo = mid; else hi = mid; }
                    Bd[lo * GD_STRIDE + ca] = va;
                }
                if (xb < nB) {
                    int lo = 0, hi = TRb;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sp[mid] <= xb) lo = mid; else hi = mid; }
                    Bd[lo * GD_STRIDE + cb] = vb;
                }
            }
            __syncwarp();
            int baseAn = 0, nAn = 0, baseBn = 0, nBn = 0, sn = 0, en = 0, pk0n = 0x7fffffff, pk1n = 0x7fffffff;
            if (more) {
                baseAn = tile_nnzA[an]; nAn = tile_nnzA[an + 1] - baseAn;
                baseBn = tile_nnzB[bn]; nBn = tile_nnzB[bn + 1] - baseBn;
                sn = ptrA[an * 32 + r]; en = r + 1 < 32 ? (int)ptrA[an * 32 + r + 1] : nAn;
                if (lane < TRb) pk0n = (int)ptrB[bn * TRb + lane];
                if (lane + 32 < TRb) pk1n = (int)ptrB[bn * TRb + lane + 32];
            }
            for (int x = s; x < e; x++) {
                int kn = 0;
                double avn = 0.0;
                if (x + 1 < e) { kn = colA[baseA + x + 1]; avn = valA[baseA + x + 1]; }  // next entry's loads fly under this entry's FMAs
                const double *brow = Bd + (kc - r * TCa) * GD_STRIDE;
#pragma unroll
                for (int c = 0; c < 32; c++) acc[c] = fma(av, brow[c], acc[c]);
                kc = kn; av = avn;
            }
            __syncwarp();
            baseA = baseAn; nA = nAn; baseB = baseBn; nB = nBn; s = sn; e = en; pk0 = pk0n; pk1 = pk1n;
        }
        const unsigned m0 = maskC[((size_t)t * 32 + r) * 2], m1 = maskC[((size_t)t * 32 + r) * 2 + 1];
        int pos = outbase + (int)ptrC[(size_t)t * 32 + r];
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const unsigned mw = c < 16 ? m0 : m1;
            if ((mw >> (15 - (c & 15))) & 1u) {
                valC[pos] = acc[c];
                colC[pos] = (uint16_t)c;
                pos++;
            }
        }
    }
}
#endif

// ------------------------------------------------------------------------------------------------------------------
// tile2csr, general (reference src/tile2csr.h:72-140): thread per matrix row, count then fill
// ------------------------------------------------------------------------------------------------------------------
GT_KERNEL k_g_row_counts(long long m, int TR, const int *__restrict__ tile_ptr, const int *__restrict__ tile_nnz,
                         const uint16_t *__restrict__ ptr, int *__restrict__ rowcnt)
{
    const long long i = GT_TID;
    if (i >= m) return;
    const int I = (int)(i / TR), r = (int)(i % TR);
    int cnt = 0;
    for (int t = tile_ptr[I]; t < tile_ptr[I + 1]; t++) {
        const int tn = tile_nnz[t + 1] - tile_nnz[t];
        const int s = ptr[(size_t)t * TR + r], e = r + 1 < TR ? ptr[(size_t)t * TR + r + 1] : tn;
        cnt += e - s;
    }
    rowcnt[i] = cnt;
}

GT_KERNEL k_g_row_fill(long long m, int TR, int TC, const int *__restrict__ tile_ptr, const int *__restrict__ tile_col,
                       const int *__restrict__ tile_nnz, const uint16_t *__restrict__ ptr, const uint16_t *__restrict__ col,
                       const double *__restrict__ val, const int *__restrict__ rowptr, int *__restrict__ colidx,
                       double *__restrict__ valout)
{
    const long long i = GT_TID;
    if (i >= m) return;
    const int I = (int)(i / TR), r = (int)(i % TR);
    size_t dst = (size_t)rowptr[i];
    for (int t = tile_ptr[I]; t < tile_ptr[I + 1]; t++) {
        const int base = tile_nnz[t], tn = tile_nnz[t + 1] - base;
        const int s = ptr[(size_t)t * TR + r], e = r + 1 < TR ? ptr[(size_t)t * TR + r + 1] : tn;
        const int c0 = tile_col[t] * TC;
        for (int x = s; x < e; x++, dst++) {
            colidx[dst] = c0 + (int)col[base + x];  // Col as stored (src/tile2csr.h:57): C / B style, Col = c
            valout[dst] = val[base + x];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------

static int g_bits_for(long long domain)
{
    int b = 1;
    while (b < 32 && (1ll << b) < domain) b++;
    return b;
}

bool gtile_size_ok(int tile_rows, int tile_cols)
{
    return tile_rows >= 16 && tile_cols >= 16 && tile_rows <= 128 && tile_cols <= 128 && tile_rows % 16 == 0 && tile_cols % 16 == 0;
}

static int gtile_check_size(int tile_rows, int tile_cols)
{
    if (gtile_size_ok(tile_rows, tile_cols)) return TSG_OK;
    set_error(TSG_ERR_UNSUPPORTED, "tile size %d x %d: rows and columns of a tile must be multiples of 16 (one mask word covers 16 columns, "
              "reference src/common.h:146) between 16 and 128", tile_rows, tile_cols);
    return last_error();
}

// one device allocation, 256-byte aligned sub-arrays (same idea as tile_alloc_layout of the 16 x 16 path)
int gtile_alloc(int m, int n, int tile_rows, int tile_cols, int numtile, long long nnz, int col_major, tsg_gtile *out)
{
    memset(out, 0, sizeof(*out));
    out->m = m; out->n = n; out->tile_rows = tile_rows; out->tile_cols = tile_cols;
    out->tilem = (m + tile_rows - 1) / tile_rows; out->tilen = (n + tile_cols - 1) / tile_cols;
    out->numtile = numtile; out->nnz = nnz; out->col_major = col_major;
    const size_t W = (size_t)tile_cols / 16;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t nt = (size_t)(numtile > 0 ? numtile : 1), nz = (size_t)(nnz > 0 ? nnz : 1);
    const size_t o_tp = take(((size_t)out->tilem + 1) * 4), o_tc = take(nt * 4), o_tr = take(nt * 4), o_tn = take((nt + 1) * 4);
    const size_t o_v = take(nz * 8), o_c = take(nz * 2), o_p = take(nt * tile_rows * 2), o_m = take(nt * tile_rows * W * 2);
    size_t o_cp = 0, o_cr = 0, o_rm = 0;
    if (col_major) { o_cp = take(((size_t)out->tilen + 1) * 4); o_cr = take(nt * 4); o_rm = take(nt * 4); }
    char *base = (char *)dalloc(off);
    if (!base) return last_error();
    out->slab[0] = base; out->slab_bytes[0] = off;
    out->tile_ptr = (int *)(base + o_tp); out->tile_columnidx = (int *)(base + o_tc); out->tile_rowidx = (int *)(base + o_tr);
    out->tile_nnz = (int *)(base + o_tn); out->val = (double *)(base + o_v); out->col = (uint16_t *)(base + o_c);
    out->ptr = (uint16_t *)(base + o_p); out->mask = (uint16_t *)(base + o_m);
    if (col_major) {
        out->csc_tile_ptr = (int *)(base + o_cp); out->csc_tile_rowidx = (int *)(base + o_cr); out->rm2csc = (int *)(base + o_rm);
    }
    return TSG_OK;
}

void gtile_free(tsg_gtile *t)
{
    for (int k = 0; k < 2; k++)
        if (t->slab[k]) dfree(t->slab[k]);
    memset(t, 0, sizeof(*t));
}

static int g_gt_input_flags = 0;  // flags of the last contract violation seen by gtile_csr2tile_device (the drop-in retries on 2)
int gtile_last_input_flags() { return g_gt_input_flags; }

#ifndef GT_EMULATE
// CUDA-event stopwatch of one call (destroyed on every return path)
struct GtTimer {
    cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ~GtTimer() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    void mark(int k) { if (!e[k]) cudaEventCreate(&e[k]); if (e[k]) cudaEventRecord(e[k], ctx().stream); }
    double ms(int a, int b) { float f = 0.f; if (e[a] && e[b]) cudaEventElapsedTime(&f, e[a], e[b]); return f; }
};
#endif

// pair_ptr[numblkC] = number of pairs: written by the thread of the last pair
GT_KERNEL k_g_pair_ptr_end(long long npairs, const int *__restrict__ head, const int *__restrict__ headscan, int *__restrict__ pair_ptr)
{
    const long long q = GT_TID;
    if (q != npairs - 1) return;
    pair_ptr[headscan[q] + head[q]] = (int)npairs;
}

// frees a list of scratch blocks on every return path
struct GtScratch {
    void *p[32];
    int n = 0;
    template <typename T> T *take(size_t count)
    {
        T *q = dalloc_n<T>(count);
        if (q && n < 32) p[n++] = q;
        return q;
    }
    ~GtScratch() { for (int k = 0; k < n; k++) dfree(p[k]); }
};

int gtile_csr2tile_device(const tsg_dcsr *A, int col_major, int TR, int TC, tsg_gtile *out)
{
    memset(out, 0, sizeof(*out));
    if (gtile_check_size(TR, TC)) return last_error();
    const int m = A->m, n = A->n;
    const long long nnz = A->nnz;
    if (nnz >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "csr2tile: nnz %lld does not fit int32", nnz); return last_error(); }
    const int tilem = (m + TR - 1) / TR, tilen = (n + TC - 1) / TC;
    GtScratch S;
    int *d_err = S.take<int>(1);
    if (!d_err) return last_error();
    CK(cudaMemsetAsync(d_err, 0, 4, ctx().stream));

    // runs of the CSR rows inside one tile column
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    int *flag = S.take<int>(nz + 1), *runidx = S.take<int>(nz + 1);
    if (!flag || !runidx) return last_error();
    GT_LAUNCH(k_g_run_flags, nnz, nnz, m, TC, A->rowptr, A->colidx, flag, d_err);
    int rc = exclusive_scan<int>(flag, runidx, nnz);
    if (rc) return rc;
    int nruns = 0;
    rc = read_back_i32(runidx + nnz, &nruns);
    if (rc) return rc;
    int eflag = 0;
    rc = read_back_i32(d_err, &eflag);
    if (rc) return rc;
    if (eflag) {
        g_gt_input_flags = eflag;
        set_error(TSG_ERR_INPUT, "csr2tile: CSR input violates the contract (flags=%d: 2=row not sorted/duplicate)", eflag);
        return last_error();
    }
    const size_t nr = (size_t)(nruns > 0 ? nruns : 1);
    int *run_start = S.take<int>(nr), *run_row = S.take<int>(nr), *tile_of_run = S.take<int>(nr);
    uint32_t *ka = S.take<uint32_t>(nr), *va = S.take<uint32_t>(nr), *kb = S.take<uint32_t>(nr), *vb = S.take<uint32_t>(nr);
    int *head = S.take<int>(nr + 1), *headscan = S.take<int>(nr + 1);
    if (!run_start || !run_row || !tile_of_run || !ka || !va || !kb || !vb || !head || !headscan) return last_error();
    GT_LAUNCH(k_g_run_emit, nnz, nnz, m, TC, A->rowptr, A->colidx, flag, runidx, run_start, run_row, ka, va);
    // stable sort by tile column, then by tile row: runs end up ordered by (tile row, tile column, matrix row)
    uint32_t *ks = ka, *perm = va;
    rc = sort_pairs_device(ka, va, kb, vb, nruns, g_bits_for(tilen), &ks, &perm);
    if (rc) return rc;
    {
        uint32_t *k2 = ks, *other_k = ks == ka ? kb : ka, *other_v = perm == va ? vb : va;
        GT_LAUNCH(k_g_gather_tilerow, nruns, nruns, TR, perm, run_row, k2);
        uint32_t *ks2 = k2, *perm2 = perm;
        rc = sort_pairs_device(k2, perm, other_k, other_v, nruns, g_bits_for(tilem), &ks2, &perm2);
        if (rc) return rc;
        ks = ks2; perm = perm2;
    }
    GT_LAUNCH(k_g_tile_heads, nruns, nruns, TR, TC, perm, run_row, run_start, A->colidx, head);
    rc = exclusive_scan<int>(head, headscan, nruns);
    if (rc) return rc;
    int numtile = 0;
    rc = read_back_i32(headscan + nruns, &numtile);
    if (rc) return rc;

    rc = gtile_alloc(m, n, TR, TC, numtile, nnz, col_major, out);
    if (rc) return rc;
    const size_t nt = (size_t)(numtile > 0 ? numtile : 1), W = (size_t)TC / 16;
    GT_LAUNCH(k_g_tile_struct, nruns, nruns, TR, TC, perm, run_row, run_start, A->colidx, head, headscan, out->tile_columnidx,
              out->tile_rowidx, tile_of_run);
    GT_LAUNCH(k_g_boundaries, (long long)numtile + 1, numtile, tilem, out->tile_rowidx, out->tile_ptr);
    if (col_major) {
        // CSC-tile order = stable sort of the row-major tile list by tile column (reference: tiling of B^T, :300-347)
        uint32_t *tk = S.take<uint32_t>(nt), *tv = S.take<uint32_t>(nt), *tk2 = S.take<uint32_t>(nt), *tv2 = S.take<uint32_t>(nt);
        int *sorted_cols = S.take<int>(nt);
        if (!tk || !tv || !tk2 || !tv2 || !sorted_cols) return last_error();
        GT_LAUNCH(k_g_copy_u32, numtile, numtile, out->tile_columnidx, tk, tv);
        uint32_t *tks = tk, *tperm = tv;
        rc = sort_pairs_device(tk, tv, tk2, tv2, numtile, g_bits_for(tilen), &tks, &tperm);
        if (rc) return rc;
        GT_LAUNCH(k_g_csc_order, numtile, numtile, tperm, tks, out->tile_rowidx, out->csc_tile_rowidx, out->rm2csc, sorted_cols);
        GT_LAUNCH(k_g_boundaries, (long long)numtile + 1, numtile, tilen, sorted_cols, out->csc_tile_ptr);
    }
    // per-tile data, at the storage id
    int *tile_cnt = S.take<int>(nt + 1);
    if (!tile_cnt) return last_error();
    CK(cudaMemsetAsync(tile_cnt, 0, (nt + 1) * 4, ctx().stream));
    CK(cudaMemsetAsync(out->ptr, 0, nt * TR * 2, ctx().stream));
    CK(cudaMemsetAsync(out->mask, 0, nt * TR * W * 2, ctx().stream));
    GT_LAUNCH(k_g_run_counts, nruns, nruns, nnz, m, TR, TC, A->rowptr, A->colidx, run_start, run_row, tile_of_run,
              col_major ? out->rm2csc : nullptr, out->ptr, out->mask, tile_cnt);
    rc = exclusive_scan<int>(tile_cnt, out->tile_nnz, numtile);
    if (rc) return rc;
    GT_LAUNCH(k_g_ptr_scan, numtile, numtile, TR, out->ptr);
    GT_LAUNCH(k_g_run_scatter, nruns, nruns, TR, TC, col_major ? 0 : 1, A->rowptr, A->colidx, A->val, run_start, run_row, tile_of_run,
              col_major ? out->rm2csc : nullptr, out->tile_nnz, out->ptr, out->val, out->col);
    CK(cudaStreamSynchronize(ctx().stream));
    return TSG_OK;
}

// Steps 1-3 for general tiles. A: row-major, tiles TRa x TCa; B: column-major, tiles TRb x TCb with TRb = TCa;
// C: row-major, tiles TRa x TCb.
int gtile_spgemm_device(const tsg_gtile *A, const tsg_gtile *B, tsg_gtile *C, tsg_stats *stats)
{
    memset(C, 0, sizeof(*C));
    if (stats) memset(stats, 0, sizeof(*stats));
    if (A->n != B->m) { set_error(TSG_ERR_UNSUPPORTED, "spgemm: A is %dx%d but B is %dx%d", A->m, A->n, B->m, B->n); return last_error(); }
    if (A->col_major || !B->col_major || !B->rm2csc) {
        set_error(TSG_ERR_UNSUPPORTED, "spgemm: A must be row-major tiled and B col-major tiled (csr2tile_col_major)");
        return last_error();
    }
    if (A->tile_cols != B->tile_rows) {
        set_error(TSG_ERR_UNSUPPORTED, "spgemm: tiles of A have %d columns but tiles of B have %d rows (A: tile_size_m x tile_size_n, B: tile_size_n x tile_size_m)",
                  A->tile_cols, B->tile_rows);
        return last_error();
    }
    const int TRc = A->tile_rows, TCc = B->tile_cols, TCa = A->tile_cols, Wc = TCc / 16;
    if (gtile_check_size(TRc, TCc) || gtile_check_size(TRc, TCa)) return last_error();
    const long long launches0 = ctx().launches;
    GtTimer tm;
    GtScratch S;
    tm.mark(0);

    // ---------------- step 1 ----------------
    const size_t ntA = (size_t)(A->numtile > 0 ? A->numtile : 1);
    int *w = S.take<int>(ntA + 1), *woff = S.take<int>(ntA + 1);
    long long *d_tot = S.take<long long>(2);
    if (!w || !woff || !d_tot) return last_error();
    GT_LAUNCH(k_g_pair_weights, A->numtile, A->numtile, A->tile_columnidx, B->tile_ptr, w);
    int rc = exclusive_scan<int>(w, woff, A->numtile, d_tot);
    if (rc) return rc;
    long long npairs = 0;
    rc = read_back_i64(d_tot, &npairs);
    if (rc) return rc;
    if (npairs >= (1ll << 31)) {
        set_error(TSG_ERR_OVERFLOW, "spgemm (general tiles): %lld tile pairs exceed 32-bit indexing", npairs);
        return last_error();
    }
    const size_t np = (size_t)(npairs > 0 ? npairs : 1);
    int *pa = S.take<int>(np), *pb = S.take<int>(np), *pair_a = S.take<int>(np), *pair_b = S.take<int>(np);
    uint32_t *keyJ = S.take<uint32_t>(np), *ka = S.take<uint32_t>(np), *va = S.take<uint32_t>(np), *kb = S.take<uint32_t>(np), *vb = S.take<uint32_t>(np);
    int *head = S.take<int>(np + 1), *headscan = S.take<int>(np + 1);
    if (!pa || !pb || !pair_a || !pair_b || !keyJ || !ka || !va || !kb || !vb || !head || !headscan) return last_error();
    GT_LAUNCH(k_g_pair_expand, npairs, npairs, A->numtile, woff, A->tile_columnidx, B->tile_ptr, B->tile_columnidx, B->rm2csc, pa, pb, keyJ, va);
    CK(cudaMemcpyAsync(ka, keyJ, np * 4, cudaMemcpyDeviceToDevice, ctx().stream));
    uint32_t *ks = ka, *perm = va;
    rc = sort_pairs_device(ka, va, kb, vb, npairs, g_bits_for(B->tilen), &ks, &perm);
    if (rc) return rc;
    {
        uint32_t *k2 = ks, *other_k = ks == ka ? kb : ka, *other_v = perm == va ? vb : va;
        GT_LAUNCH(k_g_pair_tilerow, npairs, npairs, perm, pa, A->tile_rowidx, k2);
        uint32_t *ks2 = k2, *perm2 = perm;
        rc = sort_pairs_device(k2, perm, other_k, other_v, npairs, g_bits_for(A->tilem), &ks2, &perm2);
        if (rc) return rc;
        ks = ks2; perm = perm2;
    }
    GT_LAUNCH(k_g_pair_heads, npairs, npairs, ks, perm, keyJ, head);
    rc = exclusive_scan<int>(head, headscan, npairs);
    if (rc) return rc;
    int numblkC = 0;
    rc = read_back_i32(headscan + npairs, &numblkC);
    if (rc) return rc;
    if ((long long)numblkC * TRc * Wc >= (1ll << 31)) {
        set_error(TSG_ERR_OVERFLOW, "spgemm (general tiles): %d C tiles of %d rows do not fit 32-bit indexing", numblkC, TRc);
        return last_error();
    }
    tm.mark(1);

    // C's metadata (the payload arrays follow once nnz(C) is known: a second allocation, like the 16 x 16 path)
    tsg_gtile meta;
    rc = gtile_alloc(A->m, B->n, TRc, TCc, numblkC, 0, 0, &meta);
    if (rc) return rc;
    const size_t nb = (size_t)(numblkC > 0 ? numblkC : 1);
    int *pair_ptr = S.take<int>(nb + 1), *tile_cnt = S.take<int>(nb + 1);
    if (!pair_ptr || !tile_cnt) { gtile_free(&meta); return last_error(); }
    struct MetaGuard { tsg_gtile *g; ~MetaGuard() { if (g) gtile_free(g); } } guard{&meta};
    CK(cudaMemsetAsync(pair_ptr, 0, (nb + 1) * 4, ctx().stream));  // no pairs at all: pair_ptr[0] = 0
    GT_LAUNCH(k_g_pair_emit, npairs, npairs, 0, ks, perm, keyJ, pa, pb, head, headscan, meta.tile_columnidx, meta.tile_rowidx, pair_ptr,
              pair_a, pair_b);
    GT_LAUNCH(k_g_pair_ptr_end, npairs, npairs, head, headscan, pair_ptr);
    GT_LAUNCH(k_g_boundaries, (long long)numblkC + 1, numblkC, meta.tilem, meta.tile_rowidx, meta.tile_ptr);
    tm.mark(2);

    // ---------------- step 2 ----------------
    CK(cudaMemsetAsync(tile_cnt, 0, (nb + 1) * 4, ctx().stream));
    GT_LAUNCH(k_g_symbolic, (long long)numblkC * TRc, (long long)numblkC * TRc, TRc, TCa, Wc, pair_ptr, pair_a, pair_b, A->tile_nnz, A->ptr, A->col,
              B->mask, meta.ptr, meta.mask, tile_cnt);
    rc = exclusive_scan<int>(tile_cnt, meta.tile_nnz, numblkC, d_tot);
    if (rc) return rc;
    GT_LAUNCH(k_g_ptr_scan, numblkC, numblkC, TRc, meta.ptr);
    long long nnzC = 0;
    rc = read_back_i64(d_tot, &nnzC);
    if (rc) return rc;
    if (nnzC >= (1ll << 31)) {
        set_error(TSG_ERR_OVERFLOW, "spgemm (general tiles): nnz(C) = %lld exceeds int32", nnzC);
        return last_error();
    }
    tm.mark(3);

    // ---------------- step 3 ----------------
    const size_t nzc = (size_t)(nnzC > 0 ? nnzC : 1);
    const size_t o_col = (nzc * 8 + 255) & ~(size_t)255;
    char *payload = (char *)dalloc(o_col + nzc * 2);
    if (!payload) return last_error();
    meta.val = (double *)payload; meta.col = (uint16_t *)(payload + o_col);
    meta.slab[1] = payload; meta.slab_bytes[1] = o_col + nzc * 2;
    meta.nnz = nnzC;
    bool dense32 = false;
#ifndef GT_EMULATE
    {   // 32 x 32 C tiles holding >= 128 entries on average: dense accumulator in registers (TSG_GT_NUMERIC=gather|dense overrides the rule;
        // measured: block-FEM, 192 per listed tile, 1.74 ms against 5.09 through the gather; 64^3 stencil, 78 per tile, 2.67 against 2.43)
        const char *e = getenv("TSG_GT_NUMERIC");
        const bool can = TRc == 32 && TCc == 32 && TCa <= 64 && numblkC > 0 && nnzC > 0;
        dense32 = can && (e && *e ? !strcmp(e, "dense") : nnzC >= 128ll * numblkC);
        if (dense32) {
            const size_t smem = (size_t)GD_WARPS * (TCa * GD_STRIDE + GD_MAXTRB / 2) * sizeof(double);
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_g_numeric_dense32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int grid = (numblkC + GD_WARPS - 1) / GD_WARPS;
            if (grid > ctx().num_sms * 16) grid = ctx().num_sms * 16;  // 4 resident CTAs per SM, four rounds: the warps stride over the tiles
            k_g_numeric_dense32<<<grid, GD_WARPS * 32, smem, ctx().stream>>>(
                numblkC, TCa, pair_ptr, pair_a, pair_b, A->tile_nnz, A->ptr, A->col, A->val, B->tile_nnz, B->ptr, B->col, B->val, meta.tile_nnz,
                meta.ptr, meta.mask, meta.col, meta.val);
            CK_LAUNCH();
        }
    }
#endif
    if (!dense32)
        GT_LAUNCH(k_g_numeric, nnzC, nnzC, numblkC, TRc, TCa, Wc, pair_ptr, pair_a, pair_b, A->tile_nnz, A->ptr, A->col, A->val, B->tile_nnz,
                  B->ptr, B->mask, B->val, meta.tile_nnz, meta.ptr, meta.mask, meta.col, meta.val);
    tm.mark(4);
    CK(cudaStreamSynchronize(ctx().stream));
    guard.g = nullptr;
    *C = meta;
    if (stats) {
        stats->numblkC = numblkC; stats->nnzC = nnzC; stats->pairs = npairs;
        stats->ms_step1 = tm.ms(0, 2); stats->ms_step2 = tm.ms(2, 3); stats->ms_step3 = tm.ms(3, 4); stats->ms_total = tm.ms(0, 4);
        stats->launches = (int)(ctx().launches - launches0);
        stats->tiles_dense = dense32 ? numblkC : 0;   // C tiles computed by k_g_numeric_dense32
        // algorithmic bytes, SURVEY.md 8(d), with this tile size's metadata: Ptr 2*TR, mask 2*TR*W, column index 4, tile nnz 4
        const long long mA = 2ll * A->tile_rows * (1 + A->tile_cols / 16) + 8, mB = 2ll * B->tile_rows * (1 + B->tile_cols / 16) + 8,
                        mC = 2ll * TRc * (1 + Wc) + 12;
        stats->algorithmic_bytes = A->nnz * 10 + A->numtile * mA + ((long long)A->tilem + 1) * 4 + B->nnz * 10 + B->numtile * (mB + 4) +
                                   ((long long)B->tilem + 1) * 4 + ((long long)B->tilen + 1) * 4 + nnzC * 10 + numblkC * mC +
                                   ((long long)A->tilem + 1) * 4;
    }
    return TSG_OK;
}

int gtile_tile2csr_device(const tsg_gtile *T, tsg_dcsr *out)
{
    memset(out, 0, sizeof(*out));
    if (T->col_major) { set_error(TSG_ERR_UNSUPPORTED, "tile2csr: tiles must be in row-major storage order"); return last_error(); }
    const int m = T->m;
    const long long nnz = T->nnz;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    const size_t o_ci = (((size_t)m + 1) * 4 + 255) & ~(size_t)255, o_v = o_ci + ((nz * 4 + 255) & ~(size_t)255);
    char *base = (char *)dalloc(o_v + nz * 8);
    if (!base) return last_error();
    out->m = m; out->n = T->n; out->nnz = nnz; out->owner = base;
    out->rowptr = (int *)base; out->colidx = (int *)(base + o_ci); out->val = (double *)(base + o_v);
    GtScratch S;
    int *rowcnt = S.take<int>((size_t)m + 1);
    long long *d_tot = S.take<long long>(1);
    if (!rowcnt || !d_tot) { dfree(base); memset(out, 0, sizeof(*out)); return last_error(); }
    GT_LAUNCH(k_g_row_counts, m, m, T->tile_rows, T->tile_ptr, T->tile_nnz, T->ptr, rowcnt);
    int rc = exclusive_scan<int>(rowcnt, out->rowptr, m, d_tot);
    long long total = 0;
    if (!rc) rc = read_back_i64(d_tot, &total);
    if (!rc && total != nnz) {
        set_error(TSG_ERR_INPUT, "tile2csr: the per-tile row offsets add up to %lld entries but the matrix holds %lld", total, nnz);
        rc = last_error();
    }
    if (rc) { dfree(base); memset(out, 0, sizeof(*out)); return rc; }
    GT_LAUNCH(k_g_row_fill, m, m, T->tile_rows, T->tile_cols, T->tile_ptr, T->tile_columnidx, T->tile_nnz, T->ptr, T->col, T->val, out->rowptr,
              out->colidx, out->val);
    CK(cudaStreamSynchronize(ctx().stream));
    return TSG_OK;
}

// Host SMatrix tile arrays (general tile size) -> device. tile_rows x tile_cols are the dimensions of ONE tile of this
// matrix (B of the reference: tile_size_n x tile_size_m).
int gtile_upload(const SMatrix *h, int col_major, int TR, int TC, tsg_gtile *out)
{
    memset(out, 0, sizeof(*out));
    if (gtile_check_size(TR, TC)) return last_error();
    if (h->numtile < 0 || h->nnz < 0) { set_error(TSG_ERR_INPUT, "tile upload: negative sizes"); return last_error(); }
    int rc = gtile_alloc(h->m, h->n, TR, TC, h->numtile, h->nnz, col_major, out);
    if (rc) return rc;
    if (h->tilem != out->tilem || h->tilen != out->tilen) {
        set_error(TSG_ERR_INPUT, "tile upload: the matrix says %d x %d tiles, %d x %d tiles of %d x %d cover it", h->tilem, h->tilen, out->tilem,
                  out->tilen, TR, TC);
        gtile_free(out);
        return last_error();
    }
    cudaStream_t s = ctx().stream;
    const size_t nt = (size_t)h->numtile, nz = (size_t)h->nnz, W = (size_t)TC / 16;
    CK(cudaMemcpyAsync(out->tile_ptr, h->tile_ptr, ((size_t)h->tilem + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(out->tile_nnz, h->tile_nnz, (nt + 1) * 4, cudaMemcpyHostToDevice, s));
    if (nt) {
        CK(cudaMemcpyAsync(out->tile_columnidx, h->tile_columnidx, nt * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(out->ptr, h->tile_csr_Ptr, nt * TR * 2, cudaMemcpyHostToDevice, s));
        if (h->mask) CK(cudaMemcpyAsync(out->mask, h->mask, nt * TR * W * 2, cudaMemcpyHostToDevice, s));
    }
    if (nz) {
        CK(cudaMemcpyAsync(out->val, h->tile_csr_Value, nz * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(out->col, h->tile_csr_Col, nz * 2, cudaMemcpyHostToDevice, s));
    }
    GT_LAUNCH(k_g_tile_rows, h->numtile, h->numtile, out->tilem, out->tile_ptr, out->tile_rowidx);
    if (nt && !h->mask)
        GT_LAUNCH(k_g_masks_from_tiles, (long long)h->numtile * TR, (long long)h->numtile * TR, TR, TC, col_major ? 0 : 1, out->tile_nnz, out->ptr,
                  out->col, out->mask);
    if (col_major) {
        GtScratch S;
        int *d_err = S.take<int>(1);
        if (!d_err) { gtile_free(out); return last_error(); }
        CK(cudaMemsetAsync(d_err, 0, 4, s));
        CK(cudaMemcpyAsync(out->csc_tile_ptr, h->csc_tile_ptr, ((size_t)h->tilen + 1) * 4, cudaMemcpyHostToDevice, s));
        if (nt) CK(cudaMemcpyAsync(out->csc_tile_rowidx, h->csc_tile_rowidx, nt * 4, cudaMemcpyHostToDevice, s));
        GT_LAUNCH(k_g_rm2csc, h->numtile, h->numtile, out->tile_columnidx, out->tile_rowidx, out->csc_tile_ptr, out->csc_tile_rowidx, out->rm2csc,
                  d_err);
        int eflag = 0;
        rc = read_back_i32(d_err, &eflag);
        if (!rc && eflag) {
            set_error(TSG_ERR_INPUT, "tile upload: csc_tile_ptr / csc_tile_rowidx do not list the tiles of tile_ptr / tile_columnidx");
            rc = last_error();
        }
        if (rc) { gtile_free(out); return rc; }
    }
    CK(cudaStreamSynchronize(s));
    return TSG_OK;
}

// device -> host SMatrix (arrays malloc()ed: the driver's free() / matrix_destroy() keep working)
int gtile_download(const tsg_gtile *t, SMatrix *h)
{
    if (t->nnz >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "tile download: nnz %lld does not fit SMatrix.nnz", t->nnz); return last_error(); }
    cudaStream_t s = ctx().stream;
    const size_t nt = (size_t)t->numtile, nz = (size_t)t->nnz, TR = (size_t)t->tile_rows, W = (size_t)t->tile_cols / 16;
    h->m = t->m; h->n = t->n; h->tilem = t->tilem; h->tilen = t->tilen; h->numtile = t->numtile; h->nnz = (int)t->nnz;
    h->tile_ptr = (int *)malloc(((size_t)t->tilem + 1) * 4);
    h->tile_columnidx = (int *)malloc((nt ? nt : 1) * 4);
    h->tile_rowidx = (int *)calloc(nt ? nt : 1, 4);
    h->tile_nnz = (int *)malloc((nt + 1) * 4);
    h->tile_csr_Value = (double *)malloc((nz ? nz : 1) * 8);
    h->tile_csr_Col = (uint16_t *)malloc((nz ? nz : 1) * 2);
    h->tile_csr_Ptr = (uint16_t *)malloc((nt ? nt : 1) * TR * 2);
    h->mask = (uint16_t *)malloc((nt ? nt : 1) * TR * W * 2);
    h->csc_tile_ptr = nullptr; h->csc_tile_rowidx = nullptr;
    if (t->col_major) {
        h->csc_tile_ptr = (int *)malloc(((size_t)t->tilen + 1) * 4);
        h->csc_tile_rowidx = (int *)malloc((nt ? nt : 1) * 4);
    }
    if (!h->tile_ptr || !h->tile_columnidx || !h->tile_rowidx || !h->tile_nnz || !h->tile_csr_Value || !h->tile_csr_Col || !h->tile_csr_Ptr ||
        !h->mask || (t->col_major && (!h->csc_tile_ptr || !h->csc_tile_rowidx))) {
        set_error(TSG_ERR_NOMEM, "tile download: host allocation failed");
        return last_error();
    }
    CK(cudaMemcpyAsync(h->tile_ptr, t->tile_ptr, ((size_t)t->tilem + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->tile_nnz, t->tile_nnz, (nt + 1) * 4, cudaMemcpyDeviceToHost, s));
    if (nt) {
        CK(cudaMemcpyAsync(h->tile_columnidx, t->tile_columnidx, nt * 4, cudaMemcpyDeviceToHost, s));
        // B: the reference allocates tile_rowidx and leaves it zero (src/csr2tile.h:336-337)
        if (!t->col_major) CK(cudaMemcpyAsync(h->tile_rowidx, t->tile_rowidx, nt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->tile_csr_Ptr, t->ptr, nt * TR * 2, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->mask, t->mask, nt * TR * W * 2, cudaMemcpyDeviceToHost, s));
    }
    if (nz) {
        CK(cudaMemcpyAsync(h->tile_csr_Value, t->val, nz * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->tile_csr_Col, t->col, nz * 2, cudaMemcpyDeviceToHost, s));
    }
    if (t->col_major) {
        CK(cudaMemcpyAsync(h->csc_tile_ptr, t->csc_tile_ptr, ((size_t)t->tilen + 1) * 4, cudaMemcpyDeviceToHost, s));
        if (nt) CK(cudaMemcpyAsync(h->csc_tile_rowidx, t->csc_tile_rowidx, nt * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return TSG_OK;
}

}  // namespace tsg
