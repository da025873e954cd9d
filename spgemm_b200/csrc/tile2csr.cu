// tile2csr.cu -- tiled format -> CSR on the device. Replaces the reference's serial CPU code
// (src/tile2csr.h:72-140; Tile_csr_to_csr_PTR :8-32, Tile_csr_to_csr :34-68).
//
// Same two passes as the reference, but in ONE kernel and parallel over rows: pass 1 sums the per-tile row counts
// (Ptr[r+1]-Ptr[r], the last row closing on the tile nnz, :22), pass 2 walks the tiles of the tile-row in ascending
// tile column and appends (tile_col*16 + Col, Val) at the row's cursor (:55-58). Explicit zeros are kept
// (:27-28,59-60). Works on A/C-style tiles whose Col holds the plain column (C, B) -- for A's packed row*16+col the
// column is Col & 15 and callers do not need it; tiles in row-major storage order only.
#include "common.cuh"
#include "kernels.h"

namespace tsg {

// One CTA per tile-row, warp r = matrix row 16*I + r, lanes = the tile-row's tiles. The number of entries in front of
// a tile-row is already known -- tile_nnz[tile_ptr[I]], the scanned tile offsets -- so no separate counting pass or
// global scan is needed: the 16 warps count their rows, a 16-entry prefix in shared memory gives every row pointer,
// and the warps then append (tile_col*16 + Col, Val) 32 tiles at a time: a warp scan of the per-tile counts, then lane k
// of the warp moves entry k of the flattened run (coalesced stores; round 2's first version let every lane copy its own
// tile's few entries one by one: 3x the memory instructions, strided). What the counting pass read of the first 128 tiles
// stays in registers for the copying pass. A tile-row with 10^5 tiles (R-MAT hubs) is walked by 512 threads, not 16.
constexpr int T2C_THREADS = TS * 32;

constexpr int T2C_KEEP = 4;  // per-lane (count, source, column base) of the first 4 x 32 tiles stay in registers between the two passes

__global__ void __launch_bounds__(T2C_THREADS)
k_tile2csr_rows(int m, int tilem, const int *__restrict__ tile_ptr, const int *__restrict__ tile_col,
                const int *__restrict__ tile_nnz, const uint16_t *__restrict__ ptr, const uint16_t *__restrict__ col,
                const double *__restrict__ val, int base, int *__restrict__ rowptr, int *__restrict__ out_col,
                double *__restrict__ out_val)
{
    __shared__ int s_cnt[TS];
    const int I = blockIdx.x, r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = tile_ptr[I], t1 = tile_ptr[I + 1];
    const int row = I * TS + r;
    // this lane's share of row r in tile t: n entries starting at src (position in the tiled payload), column base cb
    auto slice = [&](int t, int &n, int &src, int &cb) {
        n = 0; src = 0; cb = 0;
        if (t < t1) {
            const int b = tile_nnz[t], tn = tile_nnz[t + 1] - b;
            if (tn) {  // empty C tiles are legitimate (15/16 of them on hypersparse inputs): nothing to read
                const int p0 = ptr[(size_t)t * TS + r];
                n = (r < TS - 1 ? (int)ptr[(size_t)t * TS + r + 1] : tn) - p0;
                src = b + p0;
                if (n) cb = tile_col[t] * TS;
            }
        }
    };
    int kn[T2C_KEEP], ksrc[T2C_KEEP], kcb[T2C_KEEP];
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < T2C_KEEP; c++) {
        kn[c] = 0; ksrc[c] = 0; kcb[c] = 0;
        if (t0 + c * 32 < t1) slice(t0 + c * 32 + lane, kn[c], ksrc[c], kcb[c]);  // warp-uniform
        cnt += kn[c];
    }
    for (int t = t0 + T2C_KEEP * 32 + lane; t < t1; t += 32) {
        int n, src, cb;
        slice(t, n, src, cb);
        cnt += n;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
    if (lane == 0) s_cnt[r] = cnt;
    __syncthreads();
    int cursor = tile_nnz[t0];  // entries of the slab in front of this tile-row
    for (int q = 0; q < r; q++) cursor += s_cnt[q];
    if (lane == 0 && row < m) rowptr[row] = cursor + base;
    if (I == tilem - 1 && threadIdx.x == 0) rowptr[m] = tile_nnz[t1] + base;
    if (row >= m || cnt == 0) return;
    // the row's entries in 32 tiles, flattened: entry k belongs to the first lane whose inclusive count exceeds k (5 shuffles),
    // so consecutive lanes write consecutive CSR slots and read runs of one tile's row
    auto emit = [&](int n, int src0, int cb) {
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL_MASK, incl, 31);
        for (int k0 = 0; k0 < total; k0 += 32) {
            const int k = k0 + lane;
            int L = 0;
#pragma unroll
            for (int sft = 16; sft; sft >>= 1) {
                const int v = __shfl_sync(FULL_MASK, incl, L + sft - 1);
                if (v <= k) L += sft;
            }
            const int excl_L = __shfl_sync(FULL_MASK, incl - n, L), src_L = __shfl_sync(FULL_MASK, src0, L), cb_L = __shfl_sync(FULL_MASK, cb, L);
            if (k < total) {
                const int j = src_L + (k - excl_L);
                out_col[cursor + k] = cb_L + col[j];
                out_val[cursor + k] = val[j];
            }
        }
        cursor += total;
    };
#pragma unroll
    for (int c = 0; c < T2C_KEEP; c++)
        if (t0 + c * 32 < t1) emit(kn[c], ksrc[c], kcb[c]);  // warp-uniform
    for (int tb = t0 + T2C_KEEP * 32; tb < t1; tb += 32) {
        int n, src, cb;
        slice(tb + lane, n, src, cb);
        emit(n, src, cb);
    }
}

// Per-row sums of the stored values (C * ones): a size-independent checksum for slabs too large to
// bring back to the host. Half-warp per tile-row, lane r sums row r over the tiles of the tile-row.
__global__ void __launch_bounds__(128)
k_tile_rowsums(int m, int tilem, const int *__restrict__ tile_ptr, const int *__restrict__ tile_nnz,
               const uint16_t *__restrict__ ptr, const double *__restrict__ val, double *__restrict__ out,
               long long *__restrict__ out_cnt)
{
    const int I = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int r = threadIdx.x & 15;
    if (I >= tilem) return;
    const int row = I * TS + r;
    double s = 0.0;
    long long cnt = 0;
    const int t1 = tile_ptr[I + 1];
    for (int t = tile_ptr[I]; t < t1; t++) {
        const int base = tile_nnz[t], tnnz = tile_nnz[t + 1] - base;
        if (tnnz == 0) continue;
        const int p0 = ptr[(size_t)t * TS + r];
        const int p1 = r < TS - 1 ? (int)ptr[(size_t)t * TS + r + 1] : tnnz;
        for (int j = p0; j < p1; j++) s += val[base + j];
        cnt += p1 - p0;
    }
    if (row < m) { out[row] = s; out_cnt[row] = cnt; }
}

int tile_rowsums_device(const tsg_dtile *T, double *d_out, long long *d_cnt)
{
    Ctx &c = ctx();
    if (T->col_major) { set_error(TSG_ERR_UNSUPPORTED, "rowsums: tiles must be in row-major storage order"); return last_error(); }
    if (T->tilem > 0) {
        k_tile_rowsums<<<ceil_div((long long)T->tilem * 16, 128), 128, 0, c.stream>>>(T->m, T->tilem, T->tile_ptr, T->tile_nnz, T->ptr, T->val,
                                                                                      d_out, d_cnt);
        CK_LAUNCH();
    }
    return TSG_OK;
}

__global__ void k_add_base(int *__restrict__ p, int n, int base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += base;
}

// CSR of T into caller-provided device arrays (rowptr: T->m + 1 ints, colidx / val: T->nnz entries).
// `base` is added to every row pointer after the fill: a slab's CSR then continues the numbering of the slabs before it.
int tile2csr_into(const tsg_dtile *T, int *rowptr, int *colidx, double *val, int base)
{
    Ctx &c = ctx();
    if (T->col_major) { set_error(TSG_ERR_UNSUPPORTED, "tile2csr: tiles must be in row-major storage order"); return last_error(); }
    const int m = T->m;
    if (T->tilem > 0) {
        k_tile2csr_rows<<<T->tilem, T2C_THREADS, 0, c.stream>>>(m, T->tilem, T->tile_ptr, T->tile_columnidx, T->tile_nnz, T->ptr, T->col,
                                                                T->val, base, rowptr, colidx, val);
        CK_LAUNCH();
    } else {  // no rows: the row pointer is the single entry `base`
        CK(cudaMemsetAsync(rowptr, 0, 4, c.stream));
        if (base) {
            k_add_base<<<1, 32, 0, c.stream>>>(rowptr, 1, base);
            CK_LAUNCH();
        }
    }
    return TSG_OK;
}

int tile2csr_device(const tsg_dtile *T, tsg_dcsr *out)
{
    memset(out, 0, sizeof(*out));
    if (T->col_major) { set_error(TSG_ERR_UNSUPPORTED, "tile2csr: tiles must be in row-major storage order"); return last_error(); }
    const int m = T->m;
    const long long nnz = T->nnz;
    size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    size_t o_ci = (((size_t)m + 1) * 4 + 255) & ~(size_t)255, o_v = o_ci + ((nz * 4 + 255) & ~(size_t)255);
    char *base = (char *)dalloc(o_v + nz * 8);
    if (!base) return last_error();
    out->m = m; out->n = T->n; out->nnz = nnz; out->owner = base;
    out->rowptr = (int *)base; out->colidx = (int *)(base + o_ci); out->val = (double *)(base + o_v);
    return tile2csr_into(T, out->rowptr, out->colidx, out->val, 0);
}

}  // namespace tsg
