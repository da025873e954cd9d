// csr2tile.cu -- CSR -> 16x16 tiled format on the device, plus the stable CSR transposition and
// the nnzCub count. Replaces the reference's CPU/OpenMP code:
//   csr2tile_row_major   src/csr2tile.h:205-277  (step1_kernel :6, step2_kernel :43, step3_kernel :112)
//   csr2tile_col_major   src/csr2tile.h:279-506
//   matrix_transposition src/utils.h:161-198
//   nnzCub loop          src/main.cu:155-160
//
// Design (differs from the reference, same results):
//   * Rows are sorted, so the tiles of a tile-row are the sorted union of (col/16) over its 16
//     rows. One HALF-WARP owns a tile-row, lane r walks row r; each iteration takes the minimum
//     pending tile column over the 16 lanes (shuffle min) and every lane consumes its entries of
//     that tile. This yields, in one pass and in ascending tile-column order, the tile list, the
//     per-row counts (-> Ptr by a 16-lane scan), the tile nnz and the row masks. No O(tilem*tilen)
//     scratch, no per-entry linear search (reference :25,:71-75,:155-167).
//   * Values/columns are then scattered by a warp per tile-row with coalesced CSR reads.
//   * col_major: the CSC-tile order is a stable radix sort of the row-major tile list by tile
//     column (radix_sort.cuh) instead of tiling B^T and transposing every tile twice (:300,:455).
#include "common.cuh"
#include "scan.cuh"
#include "radix_sort.cuh"
#include "kernels.h"

namespace tsg {

static int g_input_flags = 0;  // flags of the last csr2tile contract violation (the drop-in entry points retry on 2)

// ---------------------------------------------------------------------------------------------
// Tile discovery. FILL = false: count tiles per tile-row. FILL = true: emit, at row-major tile id,
// tile column / tile row / tile nnz count / per-row exclusive offsets (Ptr) / row masks.
// ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(128)
k_tiles_merge(int m, int tilem, const int *__restrict__ rowptr, const int *__restrict__ colidx,
              int *__restrict__ tile_cnt, const int *__restrict__ tile_ptr, int *__restrict__ tile_col,
              int *__restrict__ tile_row, int *__restrict__ tile_cntnnz, uint16_t *__restrict__ ptr,
              uint16_t *__restrict__ mask)
{
    const int hw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);  // tile-row
    const int l16 = threadIdx.x & 15;
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
    if (hw >= tilem) return;  // whole half-warp leaves together
    const int row = hw * TS + l16;
    int c = 0, e = 0;
    if (row < m) { c = rowptr[row]; e = rowptr[row + 1]; }
    int t = FILL ? tile_ptr[hw] : 0;
    int ntiles = 0;
    while (true) {
        int J = c < e ? (colidx[c] >> 4) : 0x7fffffff;
        int Jmin = J;
#pragma unroll
        for (int o = 8; o; o >>= 1) Jmin = min(Jmin, __shfl_xor_sync(hmask, Jmin, o));
        if (Jmin == 0x7fffffff) break;
        int cnt = 0;
        unsigned mk = 0;
        if (J == Jmin) {
            do {
                mk |= 0x8000u >> (colidx[c] & 15);  // column c <-> bit 15-c (src/csr2tile.h:195)
                c++; cnt++;
            } while (c < e && (colidx[c] >> 4) == Jmin);
        }
        if (FILL) {
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                int v = __shfl_up_sync(hmask, incl, o, 16);
                if (l16 >= o) incl += v;
            }
            int total = __shfl_sync(hmask, incl, 15, 16);
            ptr[(size_t)t * TS + l16] = (uint16_t)(incl - cnt);  // exclusive; rows past the edge repeat the total
            mask[(size_t)t * TS + l16] = (uint16_t)mk;
            if (l16 == 0) { tile_col[t] = Jmin; tile_row[t] = hw; tile_cntnnz[t] = total; }
            t++;
        }
        ntiles++;
    }
    if (!FILL && l16 == 0) tile_cnt[hw] = ntiles;
}

// ---------------------------------------------------------------------------------------------
// Scatter values / in-tile columns. One warp per tile-row; lanes stride over the (contiguous) CSR
// entries of its 16 rows. newid maps row-major tile index -> storage id (nullptr = identity).
// PACKED: Col = r*16+c (A, src/csr2tile.h:192); otherwise Col = c (B, :475).
// Also checks the input contract (sorted, duplicate-free, in range) and raises *err.
// ---------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(128)
k_tiles_scatter(int m, int n, int tilem, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                const double *__restrict__ val, const int *__restrict__ tile_ptr,
                const int *__restrict__ tile_col, const int *__restrict__ newid,
                const int *__restrict__ tile_nnz, const uint16_t *__restrict__ ptr,
                double *__restrict__ val_out, uint16_t *__restrict__ col_out, int *__restrict__ err)
{
    const int I = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (I >= tilem) return;
    const int r0 = I * TS;
    int rpv = rowptr[min(r0 + min(lane, TS), m)];  // lanes 0..16 hold the 17 row boundaries
    const int start = __shfl_sync(FULL_MASK, rpv, 0), end = __shfl_sync(FULL_MASK, rpv, 16);
    const int tbase = tile_ptr[I], tcount = tile_ptr[I + 1] - tbase;
    for (int p0 = start; p0 < end; p0 += 32) {
        int p = p0 + lane;
        bool live = p < end;
        int pc = live ? p : end - 1;
        // r = largest l in [0,15] with rp[l] <= pc
        int r = 0;
#pragma unroll
        for (int s = 8; s; s >>= 1) {
            int v = __shfl_sync(FULL_MASK, rpv, r + s);
            if (v <= pc) r += s;
        }
        int rstart = __shfl_sync(FULL_MASK, rpv, r);
        if (!live) continue;
        int col = colidx[p];
        if (col < 0 || col >= n) { atomicOr(err, 1); continue; }
        int J = col >> 4, cc = col & 15;
        if (p > rstart && colidx[p - 1] >= col) atomicOr(err, 2);  // unsorted or duplicate
        int lo = 0, hi = tcount;  // lower_bound of J in the tile-row's tile columns
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (tile_col[tbase + mid] < J) lo = mid + 1; else hi = mid;
        }
        if (lo >= tcount || tile_col[tbase + lo] != J) { atomicOr(err, 4); continue; }
        int q = p;
        while (q > rstart && (colidx[q - 1] >> 4) == J) q--;
        int t = newid ? newid[tbase + lo] : tbase + lo;
        size_t dst = (size_t)tile_nnz[t] + ptr[(size_t)t * TS + r] + (p - q);
        val_out[dst] = val[p];
        col_out[dst] = (uint16_t)(PACKED ? (r * TS + cc) : cc);
    }
}

// csc_tile_ptr / column pointer from keys sorted ascending: ptr[j] = first position with key >= j.
__global__ void k_boundaries(const uint32_t *__restrict__ sorted_keys, long long n, int nkeys_domain,
                             int *__restrict__ out_ptr)
{
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > n) return;
    long long prev = q == 0 ? -1 : (long long)sorted_keys[q - 1];
    long long cur = q == n ? (long long)nkeys_domain : (long long)sorted_keys[q];
    for (long long j = prev + 1; j <= cur; j++) out_ptr[j] = (int)q;
}

// Gather per-tile metadata into CSC-tile order: perm[q] = row-major index of the q-th stored tile.
__global__ void k_permute_tiles(int numtile, const uint32_t *__restrict__ perm, const int *__restrict__ tile_row_rm,
                                const int *__restrict__ cnt_rm, const uint16_t *__restrict__ ptr_rm,
                                const uint16_t *__restrict__ mask_rm, int *__restrict__ csc_rowidx,
                                int *__restrict__ rm2csc, int *__restrict__ cnt_out, uint16_t *__restrict__ ptr_out,
                                uint16_t *__restrict__ mask_out)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int q = (int)(g >> 4), r = (int)(g & 15);
    if (q >= numtile) return;
    int t = (int)perm[q];
    ptr_out[(size_t)q * TS + r] = ptr_rm[(size_t)t * TS + r];
    mask_out[(size_t)q * TS + r] = mask_rm[(size_t)t * TS + r];
    if (r == 0) { csc_rowidx[q] = tile_row_rm[t]; rm2csc[t] = q; cnt_out[q] = cnt_rm[t]; }
}

__global__ void k_copy_u32(const int *__restrict__ in, uint32_t *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}

// row index of every CSR entry (binary search over rowptr; used by the transposition only)
__global__ void k_entry_rows(int m, const int *__restrict__ rowptr, long long nnz, const uint32_t *__restrict__ perm,
                             const double *__restrict__ val, int *__restrict__ out_col, double *__restrict__ out_val)
{
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    int p = (int)perm[q];
    int lo = 0, hi = m;  // largest row with rowptr[row] <= p
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (rowptr[mid] <= p) lo = mid; else hi = mid;
    }
    out_col[q] = lo;
    out_val[q] = val[p];
}

__global__ void k_nnzcub(long long nnzA, const int *__restrict__ colA, const int *__restrict__ rowptrB,
                         unsigned long long *__restrict__ out)
{
    unsigned long long s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnzA; i += (long long)gridDim.x * blockDim.x) {
        int k = colA[i];
        s += (unsigned long long)(rowptrB[k + 1] - rowptrB[k]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

static int bits_for(int domain)
{
    int b = 1;
    while (b < 31 && (1ll << b) < (long long)domain) b++;
    return b;
}

int sort_pairs_device(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, long long n, int key_bits,
                      uint32_t **keys_res, uint32_t **vals_res)
{
    if (n >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "sort: %lld items do not fit int32 positions", n); return last_error(); }
    return radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, n, key_bits, keys_res, vals_res, vals_a);
}

// Layout of a tiled matrix inside ONE device slab; a pure function of the sizes so that a peer
// GPU can allocate the same layout and receive the slab with a single broadcast.
int tile_alloc_layout(int m, int n, int numtile, long long nnz, int col_major, tsg_dtile *out)
{
    memset(out, 0, sizeof(*out));
    out->m = m; out->n = n; out->tilem = (m + TS - 1) / TS; out->tilen = (n + TS - 1) / TS;
    out->numtile = numtile; out->nnz = nnz; out->col_major = col_major;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t nt = (size_t)(numtile > 0 ? numtile : 1), nz = (size_t)(nnz > 0 ? nnz : 1);
    size_t o_tile_ptr = take(((size_t)out->tilem + 1) * 4);
    size_t o_tile_col = take(nt * 4);
    size_t o_tile_row = take(nt * 4);
    size_t o_tile_nnz = take((nt + 1) * 4);
    size_t o_val = take(nz * 8);
    size_t o_col = take(nz * 2);
    size_t o_ptr = take(nt * TS * 2);
    size_t o_mask = take(nt * TS * 2);
    size_t o_pat = take(nt * 4);
    size_t o_cscp = 0, o_cscr = 0, o_rm2 = 0;
    if (col_major) {
        o_cscp = take(((size_t)out->tilen + 1) * 4);
        o_cscr = take(nt * 4);
        o_rm2 = take(nt * 4);
    }
    char *base = (char *)dalloc(off);
    if (!base) return last_error();
    out->slab[0] = base; out->slab_bytes[0] = off;
    out->tile_ptr = (int *)(base + o_tile_ptr);
    out->tile_columnidx = (int *)(base + o_tile_col);
    out->tile_rowidx = (int *)(base + o_tile_row);
    out->tile_nnz = (int *)(base + o_tile_nnz);
    out->val = (double *)(base + o_val);
    out->col = (uint16_t *)(base + o_col);
    out->ptr = (uint16_t *)(base + o_ptr);
    out->mask = (uint16_t *)(base + o_mask);
    out->pat = (int *)(base + o_pat);
    out->npat = -1;
    if (col_major) {
        out->csc_tile_ptr = (int *)(base + o_cscp);
        out->csc_tile_rowidx = (int *)(base + o_cscr);
        out->rm2csc = (int *)(base + o_rm2);
    }
    return TSG_OK;
}

int csr2tile_device(const tsg_dcsr *A, int col_major, tsg_dtile *out)
{
    Ctx &c = ctx();
    const int m = A->m, n = A->n;
    if (A->nnz >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "csr2tile: nnz %lld does not fit int32", A->nnz); return last_error(); }
    const int tilem = (m + TS - 1) / TS, tilen = (n + TS - 1) / TS;
    int *tile_ptr_tmp = dalloc_n<int>((size_t)tilem + 1);
    if (!tile_ptr_tmp) return last_error();
    const int hw_blocks = ceil_div((long long)tilem * 16, 128);
    if (tilem > 0) {
        k_tiles_merge<false><<<hw_blocks, 128, 0, c.stream>>>(m, tilem, A->rowptr, A->colidx, tile_ptr_tmp, nullptr,
                                                               nullptr, nullptr, nullptr, nullptr, nullptr);
        CK_LAUNCH();
    }
    int rc = exclusive_scan<int>(tile_ptr_tmp, tile_ptr_tmp, tilem);
    if (rc) return rc;
    int numtile = 0;
    rc = read_back_i32(tile_ptr_tmp + tilem, &numtile);
    if (rc) return rc;
    rc = tile_alloc_layout(m, n, numtile, A->nnz, col_major, out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out->tile_ptr, tile_ptr_tmp, ((size_t)tilem + 1) * 4, cudaMemcpyDeviceToDevice, c.stream));
    dfree(tile_ptr_tmp);
    CK(cudaMemsetAsync(c.d_scalars, 0, sizeof(long long), c.stream));
    int *err = (int *)c.d_scalars;
    const int warp_blocks = ceil_div((long long)tilem * 32, 128);

    if (!col_major) {
        if (numtile > 0) {
            k_tiles_merge<true><<<hw_blocks, 128, 0, c.stream>>>(m, tilem, A->rowptr, A->colidx, nullptr, out->tile_ptr,
                                                                  out->tile_columnidx, out->tile_rowidx, out->tile_nnz,
                                                                  out->ptr, out->mask);
            CK_LAUNCH();
        }
        rc = exclusive_scan<int>(out->tile_nnz, out->tile_nnz, numtile);
        if (rc) return rc;
        if (numtile > 0) {
            k_tiles_scatter<true><<<warp_blocks, 128, 0, c.stream>>>(m, n, tilem, A->rowptr, A->colidx, A->val, out->tile_ptr,
                                                                     out->tile_columnidx, nullptr, out->tile_nnz, out->ptr,
                                                                     out->val, out->col, err);
            CK_LAUNCH();
        }
    } else {
        size_t nt = (size_t)(numtile > 0 ? numtile : 1);
        int *cnt_rm = dalloc_n<int>(nt);
        uint16_t *ptr_rm = dalloc_n<uint16_t>(nt * TS), *mask_rm = dalloc_n<uint16_t>(nt * TS);
        uint32_t *ka = dalloc_n<uint32_t>(nt), *kb = dalloc_n<uint32_t>(nt), *va = dalloc_n<uint32_t>(nt), *vb = dalloc_n<uint32_t>(nt);
        if (!cnt_rm || !ptr_rm || !mask_rm || !ka || !kb || !va || !vb) return last_error();
        uint32_t *ks = ka, *perm = va;
        if (numtile > 0) {
            k_tiles_merge<true><<<hw_blocks, 128, 0, c.stream>>>(m, tilem, A->rowptr, A->colidx, nullptr, out->tile_ptr,
                                                                  out->tile_columnidx, out->tile_rowidx, cnt_rm, ptr_rm, mask_rm);
            CK_LAUNCH();
            k_copy_u32<<<ceil_div(numtile, 256), 256, 0, c.stream>>>(out->tile_columnidx, ka, numtile);
            CK_LAUNCH();
            rc = radix_sort_pairs(ka, nullptr, kb, vb, numtile, bits_for(tilen), &ks, &perm, va);
            if (rc) return rc;
            k_permute_tiles<<<ceil_div((long long)numtile * 16, 256), 256, 0, c.stream>>>(
                numtile, perm, out->tile_rowidx, cnt_rm, ptr_rm, mask_rm, out->csc_tile_rowidx, out->rm2csc, out->tile_nnz,
                out->ptr, out->mask);
            CK_LAUNCH();
        }
        k_boundaries<<<ceil_div((long long)numtile + 1, 256), 256, 0, c.stream>>>(ks, numtile, tilen, out->csc_tile_ptr);
        CK_LAUNCH();
        rc = exclusive_scan<int>(out->tile_nnz, out->tile_nnz, numtile);
        if (rc) return rc;
        if (numtile > 0) {
            k_tiles_scatter<false><<<warp_blocks, 128, 0, c.stream>>>(m, n, tilem, A->rowptr, A->colidx, A->val, out->tile_ptr,
                                                                      out->tile_columnidx, out->rm2csc, out->tile_nnz,
                                                                      out->ptr, out->val, out->col, err);
            CK_LAUNCH();
        }
        dfree(cnt_rm); dfree(ptr_rm); dfree(mask_rm); dfree(ka); dfree(kb); dfree(va); dfree(vb);
    }
    int flag = 0;
    rc = read_back_i32(err, &flag);
    if (rc) return rc;
    if (flag) {
        g_input_flags = flag;
        set_error(TSG_ERR_INPUT, "csr2tile: CSR input violates the contract (flags=%d: 1=column out of range, 2=row not sorted/duplicate, 4=internal)", flag);
        return last_error();
    }
    return tile_patterns_device(out);  // pattern ids of the tiles (plans.cu): format metadata, like the masks
}

int last_input_flags() { return g_input_flags; }

// ---------------------------------------------------------------------------------------------
// Range validation of a CSR on the device: rowptr[0] = 0, rowptr non-decreasing, rowptr[m] = nnz, 0 <= col < n.
// Everything downstream indexes arrays by these values (k_boundaries writes ptr[key], k_nnzcub reads rowptrB[col]),
// so it runs once per CSR that enters the library (tsg_csr_upload / tsg_csr_wrap), before any other kernel sees it.
// Sortedness inside a row is checked later, by the scatter kernel of csr2tile (it only compares, never indexes).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_csr_check(int m, int n, long long nnz, const int *__restrict__ rowptr, const int *__restrict__ colidx, int *__restrict__ err)
{
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (t0 == 0 && (rowptr[0] != 0 || (long long)rowptr[m] != nnz)) bad |= 8;
    for (long long i = t0; i < m; i += stride)
        if (rowptr[i] > rowptr[i + 1] || rowptr[i] < 0) bad |= 8;
    for (long long p = t0; p < nnz; p += stride) {
        const int c = colidx[p];
        if (c < 0 || c >= n) bad |= 1;
    }
    if (bad) atomicOr(err, bad);
}

int csr_check_device(const tsg_dcsr *A)
{
    Ctx &c = ctx();
    if (A->m < 0 || A->n < 0 || A->nnz < 0) { set_error(TSG_ERR_INPUT, "CSR with negative sizes"); return last_error(); }
    CK(cudaMemsetAsync(c.d_scalars, 0, sizeof(long long), c.stream));
    const long long work = A->nnz > A->m ? A->nnz : (long long)A->m + 1;
    int blocks = (int)min((long long)c.num_sms * 16, (work + 255) / 256);
    if (blocks < 1) blocks = 1;
    k_csr_check<<<blocks, 256, 0, c.stream>>>(A->m, A->n, A->nnz, A->rowptr, A->colidx, (int *)c.d_scalars);
    CK_LAUNCH();
    int flag = 0;
    int rc = read_back_i32((int *)c.d_scalars, &flag);
    if (rc) return rc;
    if (flag) {
        set_error(TSG_ERR_INPUT, "CSR input out of range (flags=%d: 1=column index outside [0,n), 8=row pointer not monotone / rowptr[m] != nnz)", flag);
        return last_error();
    }
    return TSG_OK;
}

// ---------------------------------------------------------------------------------------------
// Rows [r0, r1) of a device CSR as a CSR of its own: a rebased copy of the row pointer; colidx / val are BORROWED
// from the parent (which must outlive the slice). This is how each rank of the multi-GPU path takes its tile-rows
// out of the broadcast matrix without a host round trip.
// ---------------------------------------------------------------------------------------------
__global__ void k_rebase(const int *__restrict__ in, int *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int base = in[0];
    if (i < n) out[i] = in[i] - base;
}

int csr_row_slice_device(const tsg_dcsr *A, int r0, int r1, tsg_dcsr *out)
{
    Ctx &c = ctx();
    memset(out, 0, sizeof(*out));
    if (r0 < 0 || r1 < r0 || r1 > A->m) { set_error(TSG_ERR_INPUT, "csr_row_slice: rows [%d,%d) outside [0,%d]", r0, r1, A->m); return last_error(); }
    int *rp = dalloc_n<int>((size_t)(r1 - r0) + 1);
    if (!rp) return last_error();
    k_rebase<<<ceil_div((long long)(r1 - r0) + 1, 256), 256, 0, c.stream>>>(A->rowptr + r0, rp, r1 - r0 + 1);
    CK_LAUNCH();
    int ends[2] = {0, 0};
    int rc = publish_words(&c.h_scalars[14], A->rowptr + r0, 1);
    if (!rc) rc = read_back_i32(A->rowptr + r1, &ends[1]);
    if (rc) { dfree(rp); return rc; }
    ends[0] = *(const volatile int *)&c.h_scalars[14];
    out->m = r1 - r0; out->n = A->n; out->nnz = (long long)ends[1] - ends[0];
    out->rowptr = rp; out->colidx = A->colidx + ends[0]; out->val = A->val + ends[0]; out->owner = rp;
    return TSG_OK;
}

// ---------------------------------------------------------------------------------------------
// Canonical form of a CSR whose rows are not sorted and / or hold duplicate columns (the reference's MatrixMarket
// loader, src/mmio_highlevel.h:593-759, neither sorts nor merges): entries ordered by (row, column), duplicates
// merged -- dup_policy 0: values summed in their original order; 1: the first one kept. Two stable LSD radix sorts
// (by column, then by row) give the order; a flag scan compacts.
// ---------------------------------------------------------------------------------------------
__global__ void k_rows_of_perm(int m, const int *__restrict__ rowptr, long long nnz, const uint32_t *__restrict__ perm,
                               uint32_t *__restrict__ rows)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const int p = (int)perm[q];
    int lo = 0, hi = m;  // largest row with rowptr[row] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] <= p) lo = mid; else hi = mid;
    }
    rows[q] = (uint32_t)lo;
}

__global__ void k_canon_flags(long long nnz, const uint32_t *__restrict__ rows, const uint32_t *__restrict__ perm,
                              const int *__restrict__ colidx, int *__restrict__ keep)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    keep[q] = q == 0 || rows[q] != rows[q - 1] || colidx[perm[q]] != colidx[perm[q - 1]];
}

__global__ void k_canon_emit(long long nnz, int dup_policy, const uint32_t *__restrict__ rows, const uint32_t *__restrict__ perm,
                             const int *__restrict__ colidx, const double *__restrict__ val, const int *__restrict__ newpos,
                             int *__restrict__ out_col, double *__restrict__ out_val)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    if (newpos[q + 1] == newpos[q]) return;  // a duplicate: merged into the first of its run
    const int col = colidx[perm[q]];
    double v = val[perm[q]];
    if (dup_policy == 0)
        for (long long x = q + 1; x < nnz && newpos[x + 1] == newpos[x]; x++) v += val[perm[x]];
    out_col[newpos[q]] = col;
    out_val[newpos[q]] = v;
}

__global__ void k_canon_rowptr(int m, const int *__restrict__ rowptr, const int *__restrict__ newpos, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= m) out[i] = newpos[rowptr[i]];
}

int csr_canonicalize_device(const tsg_dcsr *A, int dup_policy, tsg_dcsr *out)
{
    Ctx &c = ctx();
    memset(out, 0, sizeof(*out));
    const long long nnz = A->nnz;
    if (nnz >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "canonicalize: nnz too large"); return last_error(); }
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    uint32_t *ka = dalloc_n<uint32_t>(nz), *kb = dalloc_n<uint32_t>(nz), *va = dalloc_n<uint32_t>(nz), *vb = dalloc_n<uint32_t>(nz);
    int *keep = dalloc_n<int>(nz + 1);
    int rc = (!ka || !kb || !va || !vb || !keep) ? last_error() : TSG_OK;
    uint32_t *ks = ka, *perm = va;
    long long kept = 0;
    if (!rc && nnz > 0) {
        k_copy_u32<<<ceil_div(nnz, 256), 256, 0, c.stream>>>(A->colidx, ka, nnz);
        rc = radix_sort_pairs(ka, nullptr, kb, vb, nnz, bits_for(A->n), &ks, &perm, va);
        if (!rc) {  // second key: the row of every entry, in the order the first sort left them
            uint32_t *rows = ks == ka ? kb : ka, *spare_v = perm == va ? vb : va;
            k_rows_of_perm<<<ceil_div(nnz, 256), 256, 0, c.stream>>>(A->m, A->rowptr, nnz, perm, rows);
            uint32_t *ks2 = nullptr, *perm2 = nullptr;
            rc = radix_sort_pairs(rows, perm, ks, spare_v, nnz, bits_for(A->m), &ks2, &perm2, nullptr);
            ks = ks2; perm = perm2;
        }
        if (!rc) {
            k_canon_flags<<<ceil_div(nnz, 256), 256, 0, c.stream>>>(nnz, ks, perm, A->colidx, keep);
            rc = exclusive_scan<int>(keep, keep, nnz);
        }
        if (!rc) { int k32 = 0; rc = read_back_i32(keep + nnz, &k32); kept = k32; }
    }
    if (!rc) {
        const size_t kz = (size_t)(kept > 0 ? kept : 1);
        const size_t o_ci = (((size_t)A->m + 1) * 4 + 255) & ~(size_t)255, o_v = o_ci + ((kz * 4 + 255) & ~(size_t)255);
        char *base = (char *)dalloc(o_v + kz * 8);
        if (!base) rc = last_error();
        else {
            out->m = A->m; out->n = A->n; out->nnz = kept; out->owner = base;
            out->rowptr = (int *)base; out->colidx = (int *)(base + o_ci); out->val = (double *)(base + o_v);
            if (nnz > 0) {
                k_canon_emit<<<ceil_div(nnz, 256), 256, 0, c.stream>>>(nnz, dup_policy, ks, perm, A->colidx, A->val, keep, out->colidx, out->val);
                k_canon_rowptr<<<ceil_div((long long)A->m + 1, 256), 256, 0, c.stream>>>(A->m, A->rowptr, keep, out->rowptr);
                if (!cuda_ok(cudaGetLastError(), "canonicalize kernels", __FILE__, __LINE__)) rc = last_error();
            } else if (!cuda_ok(cudaMemsetAsync(out->rowptr, 0, ((size_t)A->m + 1) * 4, c.stream), "memset", __FILE__, __LINE__)) rc = last_error();
        }
    }
    dfree(ka); dfree(kb); dfree(va); dfree(vb); dfree(keep);
    if (rc && out->owner) { dfree(out->owner); memset(out, 0, sizeof(*out)); }
    return rc;
}

int transpose_device(const tsg_dcsr *A, tsg_dcsr *AT)
{
    Ctx &c = ctx();
    memset(AT, 0, sizeof(*AT));
    const long long nnz = A->nnz;
    if (nnz >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "transpose: nnz too large"); return last_error(); }
    size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    size_t o_rp = 0, o_ci = (((size_t)A->n + 1) * 4 + 255) & ~(size_t)255, o_v = o_ci + ((nz * 4 + 255) & ~(size_t)255);
    char *base = (char *)dalloc(o_v + nz * 8);
    if (!base) return last_error();
    AT->m = A->n; AT->n = A->m; AT->nnz = nnz; AT->owner = base;
    AT->rowptr = (int *)(base + o_rp); AT->colidx = (int *)(base + o_ci); AT->val = (double *)(base + o_v);
    uint32_t *ka = dalloc_n<uint32_t>(nz), *kb = dalloc_n<uint32_t>(nz), *va = dalloc_n<uint32_t>(nz), *vb = dalloc_n<uint32_t>(nz);
    if (!ka || !kb || !va || !vb) return last_error();
    uint32_t *ks = ka, *perm = va;
    if (nnz > 0) {
        k_copy_u32<<<ceil_div(nnz, 256), 256, 0, c.stream>>>(A->colidx, ka, nnz);
        CK_LAUNCH();
        int rc = radix_sort_pairs(ka, nullptr, kb, vb, nnz, bits_for(A->n), &ks, &perm, va);
        if (rc) return rc;
        k_entry_rows<<<ceil_div(nnz, 256), 256, 0, c.stream>>>(A->m, A->rowptr, nnz, perm, A->val, AT->colidx, AT->val);
        CK_LAUNCH();
    }
    k_boundaries<<<ceil_div(nnz + 1, 256), 256, 0, c.stream>>>(ks, nnz, A->n, AT->rowptr);
    CK_LAUNCH();
    dfree(ka); dfree(kb); dfree(va); dfree(vb);
    return TSG_OK;
}

// Row masks of a tiled matrix from its Ptr / Col arrays, for tiles uploaded without a mask array (SMatrix.mask == NULL).
// Half-warp per tile, lane = row; Col & 15 is the column in both layouts (A stores row*16 + col, B / C the column).
__global__ void __launch_bounds__(128)
k_masks_from_tiles(int numtile, const int *__restrict__ tile_nnz, const uint16_t *__restrict__ ptr, const uint16_t *__restrict__ col,
                   uint16_t *__restrict__ mask)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = (int)(gid >> 4), r = threadIdx.x & 15;
    if (t >= numtile) return;
    const int base = tile_nnz[t];
    const int p0 = ptr[(size_t)t * TS + r];
    const int p1 = r < TS - 1 ? (int)ptr[(size_t)t * TS + r + 1] : tile_nnz[t + 1] - base;
    unsigned m = 0;
    for (int j = p0; j < p1; j++) m |= 0x8000u >> (col[base + j] & 15);
    mask[(size_t)t * TS + r] = (uint16_t)m;
}

int masks_from_tiles_device(tsg_dtile *T)
{
    Ctx &c = ctx();
    if (T->numtile > 0) {
        k_masks_from_tiles<<<ceil_div((long long)T->numtile * 16, 128), 128, 0, c.stream>>>(T->numtile, T->tile_nnz, T->ptr, T->col, T->mask);
        CK_LAUNCH();
    }
    return TSG_OK;
}

int nnzcub_device(const tsg_dcsr *A, const tsg_dcsr *B, unsigned long long *out)
{
    Ctx &c = ctx();
    CK(cudaMemsetAsync(c.d_scalars, 0, sizeof(long long), c.stream));
    if (A->nnz > 0) {
        int blocks = (int)min((long long)c.num_sms * 8, (A->nnz + 255) / 256);
        k_nnzcub<<<blocks, 256, 0, c.stream>>>(A->nnz, A->colidx, B->rowptr, (unsigned long long *)c.d_scalars);
        CK_LAUNCH();
    }
    long long v = 0;
    int rc = read_back_i64(c.d_scalars, &v);
    *out = (unsigned long long)v;
    return rc;
}

}  // namespace tsg
