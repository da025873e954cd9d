// spgemm.cu -- TileSpGEMM steps 1 and 2 on B200 and the host orchestration of all three steps, for a slab
// [trow0, trow1) of C tile-rows. Replaces reference src/tilespgemm-cuda.h:2220-2844 (host) and its kernels:
//   step 1  tile_spgemm_step1_cuda_spa_kernel / _numeric_ (:279-392) and the nsparse hash path
//           (src/spgemm_nsparse_kernel.h:221-311,1171-1438)
//   step 2  tile_spgemm_step3_cuda_kernel_2level_halfwarp (:394-773)
//   step 3  tile_spgemm_step4_cuda_kernel_smem_v3[_halfwarp] (:1273-1952)  -> numeric.cu
//
// What is different from the reference (same results):
//   * Step 1 is a Gustavson expansion at tile level over a WINDOWED shared-memory bitmap [Jmin, Jmax] of the tile-row
//     (the reference's bitmap spans all tile columns and only exists for tilen <= 16384, else it falls back to a hash
//     path). Besides C's tile list it emits, per C tile, the list of matched (A tile, B tile) pairs, so the later
//     steps never intersect index lists (the reference re-intersects A's tile-row with B's tile-column by binary
//     search in both of its later steps and caches at most one pair, :538-547).
//   * Two kernels on the light path (tile-rows of <= S1_LIGHT_MAX pairs whose window fits a per-warp bitmap), a WARP
//     per tile-row and S1_WARPS tile-rows per CTA, each warp with its own slice of shared memory:
//       k_s1_count  weight (pairs), window and C tile count of every tile-row in one pass; everything else goes to
//                   the heavy list;
//       k_s1_fill   bitmap -> per-word prefix (rank of a tile column = two loads and a popcount) -> C's tile columns ->
//                   pair counts and cursors in shared memory -> pair lists, plus the slot of every pair in A-major order
//                   (what k_step3_sparse indexes by), plus the FUSED bitmask symbolic (step 2).
//   * Step 2 (C's row masks, Ptr, tile nnz) fused into k_s1_fill: while the warp enumerates the B tiles paired with
//     one A tile, A's row masks are warp-uniform, so the 16x16x16 boolean product of up to 32 pairs costs one
//     shared-memory load + OR per A entry. Heavy tile-rows (and matrices whose B tile-rows are too short to fill a
//     warp) use k_step2 (half-warp per C tile) or, for hypersparse tiles, k_step2_thread (one thread per C tile).
//   * Heavy tile-rows (R-MAT hubs): one 1024-thread CTA per row over a compacted list, bitmap in up to all of the
//     SM's shared memory, one global atomic per pair.
//   * Empty C tiles are kept with Ptr = mask = 0 and nnz 0 (the reference leaves them uninitialised).
//   * One scan per array (64-bit total next to 32-bit offsets), scratch in grow-only arenas, two host read-backs per
//     call on the light path (sizes of the two allocations), one more when heavy tile-rows exist.
// Superseded kernels measured on the way are described in profiles/README.md.
#include "common.cuh"
#include "scan.cuh"
#include "kernels.h"
#include "plans.cuh"
#include "pair_sort.h"

namespace tsg {

constexpr int S1_LIGHT_MAX = 2048;   // tile-rows with <= this many pairs run on one warp (deterministic pair order)
constexpr int S1_HEAVY_THREADS = 1024;  // one CTA per heavy tile-row: its window bitmap can take most of the SM's shared memory
constexpr int S1_SORT_MAX = 64;      // heavy path: pair lists up to this length are re-sorted by A tile with an insertion sort, longer ones with pair_sort.h

// ---------------------------------------------------------------------------------------------
// Step 1a: per tile-row weight w = #matched tile pairs, and the window [jlo, jhi] of tile columns
// the row can produce. (w is also the multi-GPU / slab balancing weight: nsparse set_intprod_num,
// src/spgemm_nsparse_kernel.h:135-151.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_step1_weights(int trow0, int ntr, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col,
                const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, int *__restrict__ w,
                int *__restrict__ jlo, int *__restrict__ jhi, int *__restrict__ scal /*[0]=max window words,[1]=err,[2]=max w*/)
{
    const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= ntr) return;
    const int I = trow0 + i;
    long long s = 0;
    int lo = 0x7fffffff, hi = -1;
    for (int ta = a_tile_ptr[I] + lane; ta < a_tile_ptr[I + 1]; ta += 32) {
        int K = a_tile_col[ta];
        int b0 = b_tile_ptr[K], b1 = b_tile_ptr[K + 1];
        if (b1 > b0) {
            s += b1 - b0;
            lo = min(lo, b_tile_col[b0]);
            hi = max(hi, b_tile_col[b1 - 1]);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(FULL_MASK, s, o);
        lo = min(lo, __shfl_xor_sync(FULL_MASK, lo, o));
        hi = max(hi, __shfl_xor_sync(FULL_MASK, hi, o));
    }
    if (lane == 0) {
        if (s > 0x7fffffffll) { atomicOr(&scal[1], 1); s = 0x7fffffff; }
        w[i] = (int)s;
        jlo[i] = lo;
        jhi[i] = hi;
        if (s > 0) {
            int nw = ((hi - (lo & ~31)) >> 5) + 1;
            atomicMax(&scal[0], nw);
            atomicMax(&scal[2], (int)s);
        }
    }
}

// rank of tile column J inside the window bitmap = slot of the C tile within its tile-row
__device__ __forceinline__ int s1_rank(const unsigned *bitmap, const int *pre8, int d)
{
    int wd = d >> 5;
    int r = pre8[wd >> 3];
    for (int k = wd & ~7; k < wd; k++) r += __popc(bitmap[k]);
    return r + __popc(bitmap[wd] & ((1u << (d & 31)) - 1));
}

// block-wide exclusive scan of one int per thread; returns exclusive value, *total = block sum
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (THREADS == 32) { *total = __shfl_sync(FULL_MASK, incl, 31); return incl - v; }
    __syncthreads();  // protect s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; k++) {
        int t = s_warp[k];
        if (k < warp) off += t;
        tot += t;
    }
    *total = tot;
    return off + incl - v;
}

// ---------------------------------------------------------------------------------------------
// Light path. S1_WARPS tile-rows per CTA, one warp each, each warp with its own slice of dynamic shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int S1_WARPS = 8;
enum { SC_NW_HEAVY = 0, SC_ERR = 1, SC_WMAX = 2, SC_MAXJ = 3, SC_NHEAVY = 4, SC_NW_LIGHT = 5, SC_NLIGHT = 6, SC_MAXNNZA = 7 };

// The B tile-rows a C tile-row expands: A tile ta pairs with B's tile-row K = a_tile_col[ta], tiles [b0, b1). The expansion
// loops below take them 32 A tiles at a time -- one per lane, handed out by shuffle -- and load the first 32 tile columns of
// the NEXT B tile-row while the current one is processed: a warp then waits for one load per A tile instead of a chain of
// three (a_tile_col -> b_tile_ptr -> b_tile_col). Measured on config 2: k_s1_fill is bound by exactly that latency.
struct BRange {
    int b0, b1;
    __device__ __forceinline__ BRange(const int *__restrict__ a_tile_col, const int *__restrict__ b_tile_ptr, int ta, int a1) : b0(0), b1(0)
    {
        if (ta < a1) { const int K = a_tile_col[ta]; b0 = b_tile_ptr[K]; b1 = b_tile_ptr[K + 1]; }
    }
};

// k_s1_count: per tile-row the weight w (matched tile pairs; also the multi-GPU / slab balancing weight, nsparse
// set_intprod_num, src/spgemm_nsparse_kernel.h:135-151), the window [jlo, jhi] of tile columns the row can produce, and
// -- for rows that fit the light path -- the number of distinct tile columns (C tiles) and, when bm_save is given and the
// window is at most bm_stride words, the window bitmap itself. The others join heavy_list.
__global__ void __launch_bounds__(S1_WARPS * 32)
k_s1_count(int trow0, int ntr, int bmw, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col,
           const int *__restrict__ a_tile_nnz, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, int *__restrict__ w, int *__restrict__ jlo,
           int *__restrict__ jhi, int *__restrict__ cnt, uint8_t *__restrict__ light, int *__restrict__ heavy_list,
           int *__restrict__ scal, unsigned *__restrict__ bm_save, int bm_stride, const int *__restrict__ row_list, int nlist)
{
    extern __shared__ unsigned s1c_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * S1_WARPS + warp;
    if (idx >= (row_list ? nlist : ntr)) return;
    const int i = row_list ? row_list[idx] : idx;  // tile-row templates (rowplans.cu): the representatives only
    unsigned *bitmap = s1c_smem + (size_t)warp * bmw;
    const int I = trow0 + i;
    const int a0 = a_tile_ptr[I], a1 = a_tile_ptr[I + 1];
    long long s = 0;
    int lo = 0x7fffffff, hi = -1;
    for (int ta = a0 + lane; ta < a1; ta += 32) {
        const int K = a_tile_col[ta];
        const int b0 = b_tile_ptr[K], b1 = b_tile_ptr[K + 1];
        if (b1 > b0) {
            s += b1 - b0;
            lo = min(lo, b_tile_col[b0]);
            hi = max(hi, b_tile_col[b1 - 1]);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(FULL_MASK, s, o);
        lo = min(lo, __shfl_xor_sync(FULL_MASK, lo, o));
        hi = max(hi, __shfl_xor_sync(FULL_MASK, hi, o));
    }
    if (s > 0x7fffffffll) {
        if (lane == 0) atomicOr(&scal[SC_ERR], 1);
        s = 0x7fffffff;
    }
    if (lane == 0) { w[i] = (int)s; jlo[i] = lo; jhi[i] = hi; }
    if (s == 0) {
        if (lane == 0) { cnt[i] = 0; light[i] = 0; }
        return;
    }
    const int lo32 = lo & ~31, nw = ((hi - lo32) >> 5) + 1;
    if (s > S1_LIGHT_MAX || nw > bmw) {  // heavy: more pairs than one warp should walk, or a window wider than its bitmap
        if (lane == 0) {
            cnt[i] = 0;
            light[i] = 0;
            heavy_list[atomicAdd(&scal[SC_NHEAVY], 1)] = i;
            atomicMax(&scal[SC_NW_HEAVY], nw);
            atomicMax(&scal[SC_WMAX], (int)s);
        }
        return;
    }
    for (int k = lane; k < nw; k += 32) bitmap[k] = 0;
    __syncwarp();
    for (int tc = a0; tc < a1; tc += 32) {
        const BRange br(a_tile_col, b_tile_ptr, tc + lane, a1);
        const int nt = min(32, a1 - tc);
        int nb0 = __shfl_sync(FULL_MASK, br.b0, 0), nb1 = __shfl_sync(FULL_MASK, br.b1, 0);
        int ncol = nb0 + lane < nb1 ? b_tile_col[nb0 + lane] : -1;
        for (int j = 0; j < nt; j++) {
            const int b0 = nb0, b1 = nb1, col = ncol;
            if (j + 1 < nt) {
                nb0 = __shfl_sync(FULL_MASK, br.b0, j + 1); nb1 = __shfl_sync(FULL_MASK, br.b1, j + 1);
                ncol = nb0 + lane < nb1 ? b_tile_col[nb0 + lane] : -1;
            }
            if (col >= 0) atomicOr(&bitmap[(col - lo32) >> 5], 1u << ((col - lo32) & 31));
            for (int tb = b0 + 32 + lane; tb < b1; tb += 32) {
                const int d = b_tile_col[tb] - lo32;
                atomicOr(&bitmap[d >> 5], 1u << (d & 31));
            }
        }
    }
    __syncwarp();
    int n = 0;
    const bool save = bm_save && nw <= bm_stride;
    for (int k = lane; k < nw; k += 32) {
        const unsigned word = bitmap[k];
        n += __popc(word);
        if (save) bm_save[(size_t)i * bm_stride + k] = word;  // k_s1_fill starts from it instead of expanding the row again
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(FULL_MASK, n, o);
    if (lane == 0) {
        cnt[i] = n;
        light[i] = 1;
        atomicMax(&scal[SC_MAXJ], n);
        atomicMax(&scal[SC_NW_LIGHT], nw);
        atomicMax(&scal[SC_WMAX], (int)s);
        atomicAdd(&scal[SC_NLIGHT], 1);
        atomicMax(&scal[SC_MAXNNZA], a_tile_nnz[a1] - a_tile_nnz[a0]);  // sizes the staged kernels' shared memory
    }
}

// k_s1_fill: everything else of steps 1 and 2 for a light tile-row (see the file header).
// Per-warp shared memory (32-bit words): bitmap[bmw] | pre[bmw] u16 | cur[nj] | (FUSE) bmT[16][32] u16 | cm[16][njpad] u16.
struct S1Fill {
    int trow0, ntr, bmw, nj, njpad, warp_words, hoff;
    const int *a_tile_ptr, *a_tile_col, *b_tile_ptr, *b_tile_col, *b_rm2csc;
    const int *jlo, *jhi, *wptr, *c_tile_ptr;
    const uint8_t *light;
    int *c_tile_col, *c_tile_row, *pair_ptr, *pair_end, *pair_a, *pair_b;
    uint16_t *pair_slot;
    const uint16_t *a_mask, *b_mask;
    uint16_t *c_ptr, *c_mask;
    int *c_cnt;
    // HASH: the recipe of every C tile (plans.cu) -- pattern ids of A's and B's tiles, the recipe table, the slot per C tile
    const int *pat_a, *pat_b;
    PlanTable table;
    int *rslot;
    unsigned *pair_pat;        // HASH: (A pattern << 16 | B pattern) of every pair, beside pair_a / pair_b (k_recipe_verify reads it)
    const unsigned *bm_saved;  // the window bitmaps k_s1_count saved (bm_stride words per tile-row), or null
    int bm_stride;
    const int *row_list;       // tile-row templates (rowplans.cu): run on these nlist representatives only ...
    int nlist;
    unsigned *pair_src;        // ... and record, beside every pair, where it came from: (A tile of the row << 16) | (tile of B's tile-row)
};

// FUSE: fused bitmask symbolic. HASH: instead, hash every C tile's (A pattern, B pattern) sequence into the recipe table
// (per-warp shared memory then holds hh[nj] 64-bit running hashes in place of bmT / cm).
template <bool FUSE, bool HASH>
__global__ void __launch_bounds__(S1_WARPS * 32, 5)
k_s1_fill(const __grid_constant__ S1Fill P)
{
    extern __shared__ __align__(16) unsigned s1f_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * S1_WARPS + warp;
    if (idx >= (P.row_list ? P.nlist : P.ntr)) return;
    const int i = P.row_list ? P.row_list[idx] : idx;
    if (!P.light[i]) return;
    unsigned *bitmap = s1f_smem + (size_t)warp * P.warp_words;
    uint16_t *pre = reinterpret_cast<uint16_t *>(bitmap + P.bmw);
    int *cur = reinterpret_cast<int *>(bitmap + P.bmw + P.bmw / 2);
    uint16_t *bmT = reinterpret_cast<uint16_t *>(cur + P.nj);  // [16][32]: row mask k of lane's B tile
    uint16_t *cm = bmT + TS * 32;                               // [16][njpad]: C's row masks of the tile-row
    unsigned long long *hh = reinterpret_cast<unsigned long long *>(bitmap + P.hoff);  // HASH: [nj], 8-byte aligned (warp_words, hoff even)
    const int I = P.trow0 + i;
    const int lo = P.jlo[i] & ~31, nw = ((P.jhi[i] - lo) >> 5) + 1;
    const int a0 = P.a_tile_ptr[I], a1 = P.a_tile_ptr[I + 1];
    const int cbase = P.c_tile_ptr[i], numJ = P.c_tile_ptr[i + 1] - cbase, wbase = P.wptr[i];

    // 1. window bitmap of the tile columns the row produces (as k_s1_count left it, when it was saved)
    if (P.bm_saved && nw <= P.bm_stride) {
        for (int k = lane; k < nw; k += 32) bitmap[k] = P.bm_saved[(size_t)i * P.bm_stride + k];
    } else {
        for (int k = lane; k < nw; k += 32) bitmap[k] = 0;
        __syncwarp();
        for (int tc = a0; tc < a1; tc += 32) {
            const BRange br(P.a_tile_col, P.b_tile_ptr, tc + lane, a1);
            const int nt = min(32, a1 - tc);
            int nb0 = __shfl_sync(FULL_MASK, br.b0, 0), nb1 = __shfl_sync(FULL_MASK, br.b1, 0);
            int ncol = nb0 + lane < nb1 ? P.b_tile_col[nb0 + lane] : -1;
            for (int j = 0; j < nt; j++) {
                const int b0 = nb0, b1 = nb1, col = ncol;
                if (j + 1 < nt) {
                    nb0 = __shfl_sync(FULL_MASK, br.b0, j + 1); nb1 = __shfl_sync(FULL_MASK, br.b1, j + 1);
                    ncol = nb0 + lane < nb1 ? P.b_tile_col[nb0 + lane] : -1;
                }
                if (col >= 0) atomicOr(&bitmap[(col - lo) >> 5], 1u << ((col - lo) & 31));
                for (int tb = b0 + 32 + lane; tb < b1; tb += 32) {
                    const int d = P.b_tile_col[tb] - lo;
                    atomicOr(&bitmap[d >> 5], 1u << (d & 31));
                }
            }
        }
    }
    __syncwarp();
    // 2. per-word exclusive prefix of the popcounts: rank(d) = pre[d / 32] + popc(bitmap[d / 32] below bit d % 32)
    {
        const int q = (nw + 31) >> 5, k0 = lane * q, k1 = min(k0 + q, nw);
        int sum = 0;
        for (int k = k0; k < k1; k++) sum += __popc(bitmap[k]);
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - sum;
        for (int k = k0; k < k1; k++) { pre[k] = (uint16_t)run; run += __popc(bitmap[k]); }
    }
    __syncwarp();
    // 3. C's tile columns (ascending) and rows; cursors and (fused) C masks start at zero
    for (int k = lane; k < nw; k += 32) {
        unsigned bits = bitmap[k];
        int r = cbase + pre[k];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            P.c_tile_col[r] = lo + k * 32 + b;
            P.c_tile_row[r] = I;
            r++;
        }
    }
    for (int k = lane; k < numJ; k += 32) cur[k] = 0;
    if (HASH)
        for (int k = lane; k < numJ; k += 32) hh[k] = 0x13198A2E03707344ull;
    if (FUSE)
        for (int k = lane; k < TS * P.njpad / 2; k += 32) reinterpret_cast<unsigned *>(cm)[k] = 0;
    __syncwarp();
    // 4. second expansion: pairs per C tile, the slot of every pair in A-major order, and the fused bitmask symbolic
    int aoff = 0;
    for (int tc = a0; tc < a1; tc += 32) {
        const BRange br(P.a_tile_col, P.b_tile_ptr, tc + lane, a1);
        const int nt = min(32, a1 - tc);
        int nb0 = __shfl_sync(FULL_MASK, br.b0, 0), nb1 = __shfl_sync(FULL_MASK, br.b1, 0);
        int ncol = nb0 + lane < nb1 ? P.b_tile_col[nb0 + lane] : -1;
        for (int j = 0; j < nt; j++) {
            const int ta = tc + j, b0 = nb0, b1 = nb1, col0 = ncol;
            if (j + 1 < nt) {
                nb0 = __shfl_sync(FULL_MASK, br.b0, j + 1); nb1 = __shfl_sync(FULL_MASK, br.b1, j + 1);
                ncol = nb0 + lane < nb1 ? P.b_tile_col[nb0 + lane] : -1;
            }
            unsigned amw[8];
            if (FUSE) {  // A's 16 row masks: one 32-byte line, the same for every lane
                const uint4 *ap = reinterpret_cast<const uint4 *>(P.a_mask + (size_t)ta * TS);
                const uint4 x = ap[0], y = ap[1];
                amw[0] = x.x; amw[1] = x.y; amw[2] = x.z; amw[3] = x.w; amw[4] = y.x; amw[5] = y.y; amw[6] = y.z; amw[7] = y.w;
            }
            for (int tb0 = b0; tb0 < b1; tb0 += 32) {
                const int tb = tb0 + lane;
                const bool valid = tb < b1;
                int slot = 0;
                if (valid) {
                    const int d = (tb0 == b0 ? col0 : P.b_tile_col[tb]) - lo, wd = d >> 5;
                    slot = pre[wd] + __popc(bitmap[wd] & ((1u << (d & 31)) - 1));
                    cur[slot]++;  // one A tile at a time: the lanes hold distinct slots
                    if (!HASH) P.pair_slot[wbase + aoff + (tb - b0)] = (uint16_t)slot;  // the plan kernels do not use it (k_pair_slots if they fail)
                }
                if (FUSE) {
                    if (valid) {
                        const uint4 *bp = reinterpret_cast<const uint4 *>(P.b_mask + (size_t)P.b_rm2csc[tb] * TS);
                        const uint4 x = bp[0], y = bp[1];
                        const unsigned bw[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
                        for (int q = 0; q < 8; q++) {  // little-endian u16 pairs
                            bmT[(2 * q) * 32 + lane] = (uint16_t)(bw[q] & 0xFFFFu);
                            bmT[(2 * q + 1) * 32 + lane] = (uint16_t)(bw[q] >> 16);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < TS; r++) {
                        unsigned m = (r & 1) ? (amw[r >> 1] >> 16) : (amw[r >> 1] & 0xFFFFu);
                        if (m) {  // warp-uniform
                            unsigned acc = 0;
                            do {
                                const int k = __clz(m) - 16;
                                acc |= bmT[k * 32 + lane];
                                m &= ~(0x8000u >> k);
                            } while (m);
                            if (valid) cm[r * P.njpad + slot] |= (uint16_t)acc;  // distinct slots per lane
                        }
                    }
                }
                __syncwarp();
            }
            aoff += b1 - b0;
        }
    }
    // 5. exclusive scan of the pair counts over the row's slots: pair_ptr; cur becomes the write cursor
    {
        int carry = wbase;
        for (int s0 = 0; s0 < numJ; s0 += 32) {
            const int sl = s0 + lane;
            const int v = sl < numJ ? cur[sl] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += t;
            }
            if (sl < numJ) { P.pair_ptr[cbase + sl] = carry + incl - v; cur[sl] = carry + incl - v; }
            carry += __shfl_sync(FULL_MASK, incl, 31);
        }
    }
    __syncwarp();
    // 6. third expansion: the pair lists, A tiles ascending inside every list (the serial SPA's summation order)
    for (int tc = a0; tc < a1; tc += 32) {
        const BRange br(P.a_tile_col, P.b_tile_ptr, tc + lane, a1);
        const unsigned mypa = HASH && tc + lane < a1 ? (unsigned)P.pat_a[tc + lane] : 0u;
        const int nt = min(32, a1 - tc);
        int nb0 = __shfl_sync(FULL_MASK, br.b0, 0), nb1 = __shfl_sync(FULL_MASK, br.b1, 0);
        int ncol = -1, nrm = 0;
        unsigned npb = 0;
        if (nb0 + lane < nb1) { ncol = P.b_tile_col[nb0 + lane]; nrm = P.b_rm2csc[nb0 + lane]; if (HASH) npb = (unsigned)P.pat_b[nb0 + lane]; }
        for (int j = 0; j < nt; j++) {
            const int ta = tc + j, b0 = nb0, b1 = nb1, col0 = ncol, rm0 = nrm;
            const unsigned pb0 = npb, pa = __shfl_sync(FULL_MASK, mypa, j);  // pattern ids (< 2^13); B's are in row-major tile order
            if (j + 1 < nt) {
                nb0 = __shfl_sync(FULL_MASK, br.b0, j + 1); nb1 = __shfl_sync(FULL_MASK, br.b1, j + 1);
                ncol = -1;
                if (nb0 + lane < nb1) { ncol = P.b_tile_col[nb0 + lane]; nrm = P.b_rm2csc[nb0 + lane]; if (HASH) npb = (unsigned)P.pat_b[nb0 + lane]; }
            }
            for (int tb = b0 + lane; tb < b1; tb += 32) {
                const bool first = tb < b0 + 32;
                const int d = (first ? col0 : P.b_tile_col[tb]) - lo, wd = d >> 5;
                const int slot = pre[wd] + __popc(bitmap[wd] & ((1u << (d & 31)) - 1));
                const int pos = cur[slot]++;
                const int b = first ? rm0 : P.b_rm2csc[tb];
                P.pair_a[pos] = ta;
                P.pair_b[pos] = b;
                if (HASH) {
                    const unsigned pb = first ? pb0 : (unsigned)P.pat_b[tb];
                    hh[slot] = plans::mix64(hh[slot], ((unsigned long long)pa << 32) | pb);
                    P.pair_pat[pos] = (pa << 16) | pb;
                    if (P.pair_src) P.pair_src[pos] = ((unsigned)(ta - a0) << 16) | (unsigned)(tb - b0);  // light rows: both < 2048
                }
            }
            __syncwarp();
        }
    }
    // 7. list ends; Ptr (exclusive row offsets), mask and nnz of each C tile
    for (int sl = lane; sl < numJ; sl += 32) {
        P.pair_end[cbase + sl] = cur[sl];
        if (HASH) {  // the recipe of C tile cbase + sl: its slot in the table; the smallest tile index owns the slot
            const int t = cbase + sl;
            const int slot = plans::table_insert(P.table.keys, plans::RCAP, hh[sl], P.table.count, plans::RMAX, P.table.fail);
            P.rslot[t] = slot;
            if (slot >= 0 && t < P.table.owner[slot]) atomicMin(&P.table.owner[slot], t);  // owner only decreases: a stale read costs one atomic
        }
        if (FUSE) {
            unsigned pw[8], mw[8];
            int run = 0;
#pragma unroll
            for (int r = 0; r < TS; r += 2) {
                const unsigned m0 = cm[r * P.njpad + sl], m1 = cm[(r + 1) * P.njpad + sl];
                const int p0 = run, p1 = run + __popc(m0);
                run = p1 + __popc(m1);
                pw[r >> 1] = (unsigned)p0 | ((unsigned)p1 << 16);
                mw[r >> 1] = m0 | (m1 << 16);
            }
            uint4 *dp = reinterpret_cast<uint4 *>(P.c_ptr + (size_t)(cbase + sl) * TS);
            uint4 *dm = reinterpret_cast<uint4 *>(P.c_mask + (size_t)(cbase + sl) * TS);
            dp[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]); dp[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
            dm[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]); dm[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
            P.c_cnt[cbase + sl] = run;
        }
    }
}

// The A-major slot array k_step3_sparse needs, for the one case k_s1_fill<.., HASH> did not write it: the recipe plans
// were attempted and failed (a 64-bit collision, too many recipes), so the generic numeric kernels run after all.
// One warp per light tile-row; the slot of (A tile, B tile) = position of the B tile's column in the row's C tile columns.
__global__ void __launch_bounds__(S1_WARPS * 32)
k_pair_slots(int trow0, int ntr, const uint8_t *__restrict__ light, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col,
             const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, const int *__restrict__ c_tile_ptr,
             const int *__restrict__ c_tile_col, const int *__restrict__ wptr, uint16_t *__restrict__ pair_slot)
{
    const int i = blockIdx.x * S1_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= ntr || !light[i]) return;
    const int I = trow0 + i, cbase = c_tile_ptr[i], numJ = c_tile_ptr[i + 1] - cbase;
    int out = wptr[i];
    for (int ta = a_tile_ptr[I]; ta < a_tile_ptr[I + 1]; ta++) {
        const int K = a_tile_col[ta], b0 = b_tile_ptr[K], b1 = b_tile_ptr[K + 1];
        for (int tb = b0 + lane; tb < b1; tb += 32) {
            const int col = b_tile_col[tb];
            int lo = 0, hi = numJ - 1;  // the column is there: step 1 listed it
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (c_tile_col[cbase + mid] < col) lo = mid + 1; else hi = mid;
            }
            pair_slot[out + (tb - b0)] = (uint16_t)lo;
        }
        out += b1 - b0;
    }
}

// ---------------------------------------------------------------------------------------------
// Heavy path: one 1024-thread CTA per tile-row of heavy_list (more than S1_LIGHT_MAX pairs, or a window wider than a
// warp's bitmap: R-MAT hubs). MODE 0 counts the distinct tile columns. MODE 1 emits the sorted tile-column list and,
// per C tile, the matched (A tile, B tile) pair list: ONE expansion and ONE global atomic per pair -- the atomicAdd
// that counts the pairs of a C tile also returns the pair's rank inside the tile's list; (slot, rank, A tile, B tile)
// is parked in a per-row scratch record, the counts are scanned, and the records are then placed.
// Dynamic smem: bitmap[nw_max] | pre8[nw_max/8 + 1] -- or, when a window does not fit shared memory (matrices wider than
// ~29 M columns whose hub tile-rows span all of them; the reference falls back to a global-memory hash there,
// src/spgemm_nsparse_kernel.h:1253-1283), the same two arrays in a per-CTA slice of GLOBAL memory (gscratch): the grid is then
// a few CTAs per SM that walk heavy_list in a loop. The symbolic of these rows is k_step2 / k_step2_thread.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(S1_HEAVY_THREADS)
k_s1_heavy(int trow0, int nw_max, int n_heavy, unsigned *__restrict__ gscratch, const int *__restrict__ heavy_list,
           const int *__restrict__ a_tile_ptr,
           const int *__restrict__ a_tile_col, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col,
           const int *__restrict__ b_rm2csc, const int *__restrict__ w, const int *__restrict__ jlo,
           const int *__restrict__ jhi, int *__restrict__ cnt /*MODE0 out*/, const int *__restrict__ c_tile_ptr,
           const int *__restrict__ wptr, int *__restrict__ c_tile_col, int *__restrict__ c_tile_row,
           int *__restrict__ pair_ptr, int *__restrict__ pair_end, int *__restrict__ pair_a, int *__restrict__ pair_b,
           int4 *__restrict__ pair_tmp)
{
    constexpr int THREADS = S1_HEAVY_THREADS, NWARPS = THREADS / 32;
    extern __shared__ unsigned s1_smem[];
    __shared__ int s_warp[NWARPS];
    __shared__ int s_carry;
    unsigned *bitmap = gscratch ? gscratch + (size_t)blockIdx.x * ((size_t)nw_max + nw_max / 8 + 2) : s1_smem;
    int *pre8 = (int *)(bitmap + nw_max);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
    __syncthreads();  // the previous row of this CTA is done with the bitmap
    const int i = heavy_list[h], I = trow0 + i;
    const int wi = w[i];
    const int lo = jlo[i] & ~31;
    const int nw = ((jhi[i] - lo) >> 5) + 1;
    const int a0 = a_tile_ptr[I], a1 = a_tile_ptr[I + 1];

    for (int k = tid; k < nw; k += THREADS) bitmap[k] = 0;
    __syncthreads();
    for (int ta = a0 + warp; ta < a1; ta += NWARPS) {
        int K = a_tile_col[ta];
        for (int tb = b_tile_ptr[K] + lane; tb < b_tile_ptr[K + 1]; tb += 32) {
            int d = b_tile_col[tb] - lo;
            atomicOr(&bitmap[d >> 5], 1u << (d & 31));
        }
    }
    __syncthreads();
    if (MODE == 0) {
        int s = 0, total;
        for (int k = tid; k < nw; k += THREADS) s += __popc(bitmap[k]);
        block_excl_scan<THREADS>(s, s_warp, &total);
        if (tid == 0) cnt[i] = total;
        continue;
    }
    // ---- MODE 1 ----
    const int ngroups = (nw + 7) >> 3;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int g0 = 0; g0 < ngroups; g0 += THREADS) {
        int g = g0 + tid, s = 0;
        if (g < ngroups)
            for (int k = g * 8; k < min(g * 8 + 8, nw); k++) s += __popc(bitmap[k]);
        int total;
        int ex = block_excl_scan<THREADS>(s, s_warp, &total);
        int carry = s_carry;
        if (g < ngroups) pre8[g] = carry + ex;
        __syncthreads();
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    const int numJ = s_carry;
    const int cbase = c_tile_ptr[i];
    for (int k = tid; k < nw; k += THREADS) {
        unsigned bits = bitmap[k];
        if (bits) {
            int r = pre8[k >> 3];
            for (int q = k & ~7; q < k; q++) r += __popc(bitmap[q]);
            while (bits) {
                int b = __ffs(bits) - 1;
                c_tile_col[cbase + r] = lo + k * 32 + b;
                c_tile_row[cbase + r] = I;
                r++;
                bits &= bits - 1;
            }
        }
    }
    __syncthreads();
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const unsigned lt = (1u << lane) - 1;
    int4 *tmp = pair_tmp + wptr[i];
    for (int ta = a0 + warp; ta < a1; ta += NWARPS) {
        const int K = a_tile_col[ta];
        const int b0 = b_tile_ptr[K], b1 = b_tile_ptr[K + 1];
        for (int tb0 = b0; tb0 < b1; tb0 += 32) {
            const int tb = tb0 + lane;
            const bool valid = tb < b1;
            int slot = 0, off = 0;
            if (valid) {
                slot = s1_rank(bitmap, pre8, b_tile_col[tb] - lo);
                off = atomicAdd(&pair_end[cbase + slot], 1);
            }
            const unsigned mask = __ballot_sync(FULL_MASK, valid);  // lane 0 is always valid
            const int leader = __ffs(mask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_carry, __popc(mask));
            base = __shfl_sync(FULL_MASK, base, leader);
            if (valid) tmp[base + __popc(mask & lt)] = make_int4(slot, off, ta, b_rm2csc[tb]);
        }
    }
    __syncthreads();
    if (tid == 0) s_carry = wptr[i];
    __syncthreads();
    for (int s0 = 0; s0 < numJ; s0 += THREADS) {  // counts -> [pair_ptr, pair_end)
        int sidx = s0 + tid;
        int v = sidx < numJ ? pair_end[cbase + sidx] : 0, total;
        int ex = block_excl_scan<THREADS>(v, s_warp, &total);
        int carry = s_carry;
        if (sidx < numJ) { pair_ptr[cbase + sidx] = carry + ex; pair_end[cbase + sidx] = carry + ex + v; }
        __syncthreads();
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    for (int e = tid; e < wi; e += THREADS) {
        const int4 rec = tmp[e];
        const int pos = pair_ptr[cbase + rec.x] + rec.y;
        pair_a[pos] = rec.z;
        pair_b[pos] = rec.w;
    }
    __syncthreads();
    // several warps appended concurrently: restore ascending-A-tile order so that the FP64 summation order of every C
    // entry is the serial SPA's and values are reproducible run to run (insertion sort for short lists, an in-place
    // heap sort -- pair_sort.h -- for the long ones of hub tile-rows; a thread owns a list)
    for (int sidx = tid; sidx < numJ; sidx += THREADS) {
        int b = pair_ptr[cbase + sidx], e = pair_end[cbase + sidx], len = e - b;
        if (len > 1 && len <= S1_SORT_MAX) {
            for (int x = b + 1; x < e; x++) {
                int ka = pair_a[x], kb = pair_b[x], y = x - 1;
                while (y >= b && pair_a[y] > ka) { pair_a[y + 1] = pair_a[y]; pair_b[y + 1] = pair_b[y]; y--; }
                pair_a[y + 1] = ka; pair_b[y + 1] = kb;
            }
        } else if (len > S1_SORT_MAX) {
            pair_heap_sort(pair_a + b, pair_b + b, len);
        }
    }
    }  // heavy_list loop
}

// ---------------------------------------------------------------------------------------------
// Step 2: bitmask symbolic. Half-warp per C tile, lane r owns C row r:
//   maskC[r] = OR over pairs, over set bits k of maskA[r], of maskB[k].
// Writes C's row masks, per-tile exclusive row offsets (Ptr) and the tile nnz count.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_step2(int numblkC, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end,
        const int *__restrict__ pair_a, const int *__restrict__ pair_b, const uint16_t *__restrict__ a_mask,
        const uint16_t *__restrict__ b_mask, uint16_t *__restrict__ c_ptr, uint16_t *__restrict__ c_mask,
        int *__restrict__ c_cnt, const int *__restrict__ c_tile_row, const uint8_t *__restrict__ light, int trow0)
{
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int l16 = threadIdx.x & 15;
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
    if (t >= numblkC) return;
    if (light && light[c_tile_row[t] - trow0]) return;  // masks already produced by the fused step-1 path
    unsigned cm = 0;
    const int pe = pair_end[t];
    for (int p = pair_ptr[t]; p < pe; p++) {
        const int a = pair_a[p], b = pair_b[p];
        unsigned am = a_mask[(size_t)a * TS + l16];
        const unsigned bm = b_mask[(size_t)b * TS + l16];
        while (__any_sync(hmask, am != 0)) {
            int k = __clz(am) - 16;  // smallest column with a set bit (bit 15-k); 16 when am == 0
            unsigned v = __shfl_sync(hmask, bm, k & 15, 16);
            if (am) { cm |= v; am &= ~(0x8000u >> k); }
        }
    }
    int n = __popc(cm), incl = n;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        int v = __shfl_up_sync(hmask, incl, o, 16);
        if (l16 >= o) incl += v;
    }
    c_ptr[(size_t)t * TS + l16] = (uint16_t)(incl - n);
    c_mask[(size_t)t * TS + l16] = (uint16_t)cm;
    if (l16 == 15) c_cnt[t] = incl;
}

// ---------------------------------------------------------------------------------------------
// Step 2 for HYPERSPARSE tiles (R-MAT: ~1 pair of ~1 entry per C tile): one THREAD per C tile. The
// thread walks the tile's pairs and, for each entry (r,k) of the A tile, ORs row mask k of the B tile
// into its private row mask r, kept in shared memory as cm[r][thread] (dynamic r without local
// memory, conflict-free). A half-warp per tile (k_step2) would idle 15 of 16 lanes here; on
// well-filled tiles the per-thread walks are uncoalesced and k_step2 wins (profiles/README.md).
// ---------------------------------------------------------------------------------------------
constexpr int S2T_THREADS = 128;

__global__ void __launch_bounds__(S2T_THREADS)
k_step2_thread(int numblkC, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end,
               const int *__restrict__ pair_a, const int *__restrict__ pair_b, const int *__restrict__ a_tile_nnz,
               const uint16_t *__restrict__ a_col, const uint16_t *__restrict__ b_mask, uint16_t *__restrict__ c_ptr,
               uint16_t *__restrict__ c_mask, int *__restrict__ c_cnt, const int *__restrict__ c_tile_row,
               const uint8_t *__restrict__ light, int trow0)
{
    __shared__ uint16_t cm[TS][S2T_THREADS];
    const int t = blockIdx.x * S2T_THREADS + threadIdx.x, tid = threadIdx.x;
    if (t >= numblkC) return;
    if (light && light[c_tile_row[t] - trow0]) return;  // masks already produced by the fused step-1 path
#pragma unroll
    for (int r = 0; r < TS; r++) cm[r][tid] = 0;
    const int p1 = pair_end[t];
    for (int p = pair_ptr[t]; p < p1; p++) {
        const int a = pair_a[p];
        const uint16_t *bm = b_mask + (size_t)pair_b[p] * TS;
        const int e1 = a_tile_nnz[a + 1];
        for (int e = a_tile_nnz[a]; e < e1; e++) {
            const unsigned col = a_col[e];  // A stores row*16+col
            cm[col >> 4][tid] |= bm[col & 15];
        }
    }
    unsigned pw[8], mw[8];
    int run = 0;
#pragma unroll
    for (int r = 0; r < TS; r += 2) {
        const unsigned m0 = cm[r][tid], m1 = cm[r + 1][tid];
        const int p0 = run, pn = run + __popc(m0);
        run = pn + __popc(m1);
        pw[r >> 1] = (unsigned)p0 | ((unsigned)pn << 16);
        mw[r >> 1] = m0 | (m1 << 16);
    }
    uint4 *dp = reinterpret_cast<uint4 *>(c_ptr + (size_t)t * TS);
    uint4 *dm = reinterpret_cast<uint4 *>(c_mask + (size_t)t * TS);
    dp[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]); dp[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
    dm[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]); dm[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
    c_cnt[t] = run;
}

// first/last tile and nonzero of A's tile-rows [trow0, trow1), stored straight into mapped host memory
__global__ void k_slab_extent(const int *__restrict__ tile_ptr, const int *__restrict__ tile_nnz, int trow0, int trow1, int *out)
{
    if (threadIdx.x == 0) {
        const int t0 = tile_ptr[trow0], t1 = tile_ptr[trow1];
        out[0] = t0; out[1] = t1; out[2] = tile_nnz[t0]; out[3] = tile_nnz[t1];
    }
}

// row-major tile index -> CSC storage id for a B uploaded from a host SMatrix (csr2tile_device
// fills rm2csc itself). One thread per stored tile: binary search its column in its tile-row.
__global__ void k_build_rm2csc(int tilen, const int *__restrict__ csc_tile_ptr, const int *__restrict__ csc_tile_rowidx,
                               const int *__restrict__ tile_ptr, const int *__restrict__ tile_col, int numtile,
                               int *__restrict__ rm2csc, int *__restrict__ err)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= numtile) return;
    int lo = 0, hi = tilen;  // tile column J of stored tile q: csc_tile_ptr[J] <= q < csc_tile_ptr[J+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (csc_tile_ptr[mid] <= q) lo = mid; else hi = mid;
    }
    const int J = lo, I = csc_tile_rowidx[q];
    int l = tile_ptr[I], h = tile_ptr[I + 1];
    while (l < h) {
        int mid = (l + h) >> 1;
        if (tile_col[mid] < J) l = mid + 1; else h = mid;
    }
    if (l < tile_ptr[I + 1] && tile_col[l] == J) rm2csc[l] = q;
    else atomicOr(err, 1);
}

int build_rm2csc_device(tsg_dtile *B)
{
    Ctx &c = ctx();
    if (B->numtile == 0) return TSG_OK;
    CK(cudaMemsetAsync(c.d_scalars, 0, sizeof(long long), c.stream));
    k_build_rm2csc<<<ceil_div(B->numtile, 256), 256, 0, c.stream>>>(B->tilen, B->csc_tile_ptr, B->csc_tile_rowidx, B->tile_ptr,
                                                                    B->tile_columnidx, B->numtile, B->rm2csc, (int *)c.d_scalars);
    CK_LAUNCH();
    int flag = 0;
    int rc = read_back_i32((int *)c.d_scalars, &flag);
    if (rc) return rc;
    if (flag) { set_error(TSG_ERR_INPUT, "B: csc_tile_* and tile_ptr/tile_columnidx are inconsistent"); return last_error(); }
    return TSG_OK;
}

int tilerow_weights_device(const tsg_dtile *A, const tsg_dtile *B, int **d_w, int **d_jlo, int **d_jhi)
{
    Ctx &c = ctx();
    const int ntr = A->tilem;
    int *w = dalloc_n<int>((size_t)ntr + 1), *jlo = dalloc_n<int>(ntr), *jhi = dalloc_n<int>(ntr);
    if (!w || !jlo || !jhi) return last_error();
    CK(cudaMemsetAsync(c.d_scalars, 0, 4 * sizeof(int), c.stream));
    if (ntr > 0) {
        k_step1_weights<<<ceil_div((long long)ntr * 32, 128), 128, 0, c.stream>>>(0, ntr, A->tile_ptr, A->tile_columnidx, B->tile_ptr,
                                                                                 B->tile_columnidx, w, jlo, jhi, (int *)c.d_scalars);
        CK_LAUNCH();
    }
    *d_w = w; *d_jlo = jlo; *d_jhi = jhi;
    return TSG_OK;
}

static int spgemm_device_impl(const tsg_dtile *A, const tsg_dtile *B, int trow0, int trow1, tsg_dtile *C, tsg_stats *stats, bool try_rowplans,
                              bool *retry);

int spgemm_device(const tsg_dtile *A, const tsg_dtile *B, int trow0, int trow1, tsg_dtile *C, tsg_stats *stats)
{
    bool retry = false;
    int rc = spgemm_device_impl(A, B, trow0, trow1, C, stats, rowplans_env_on(), &retry);
    if (retry) {  // the tile-row templates did not hold (a heavy representative, a 64-bit collision): once more without them
        tsg_tile_free(C);
        rc = spgemm_device_impl(A, B, trow0, trow1, C, stats, false, &retry);
        if (!rc && stats) stats->row_templates = -1;
    }
    return rc;
}

static int spgemm_device_impl(const tsg_dtile *A, const tsg_dtile *B, int trow0, int trow1, tsg_dtile *C, tsg_stats *stats, bool try_rowplans,
                              bool *retry)
{
    Ctx &c = ctx();
    *retry = false;
    memset(C, 0, sizeof(*C));
    if (A->n != B->m) { set_error(TSG_ERR_UNSUPPORTED, "spgemm: A is %dx%d but B is %dx%d", A->m, A->n, B->m, B->n); return last_error(); }
    if (A->col_major || !B->col_major || !B->rm2csc) {
        set_error(TSG_ERR_UNSUPPORTED, "spgemm: A must be row-major tiled and B col-major tiled (csr2tile_col_major)");
        return last_error();
    }
    if (trow1 < 0 || trow1 > A->tilem) trow1 = A->tilem;
    if (trow0 < 0) trow0 = 0;
    if (trow0 > trow1) trow0 = trow1;
    const int ntr = trow1 - trow0;
    const long long launches0 = c.launches;
    // timing events of this call; destroyed on every return path (errors included)
    struct Events {
        cudaEvent_t e[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } evs;
    for (int k = 0; k < 7; k++) CK(cudaEventCreate(&evs.e[k]));
    cudaEvent_t *ev = evs.e;
    const cudaEvent_t ev_s2 = evs.e[5], ev_s3 = evs.e[6];
    CK(cudaEventRecord(ev[0], c.stream));

    // ---------------- step 1a: weights, windows, C tile counts ----------------
    const size_t nr = (size_t)ntr + 1;
    // a warp's window bitmap in k_s1_count: up to 2048 words (65536 tile columns); wider windows take the heavy path
    int bmw1 = (B->tilen + 31) / 32 + 1;
    if (bmw1 > 2048) bmw1 = 2048;
    // the bitmaps of windows up to bm_stride words go from k_s1_count to k_s1_fill through global memory (<= 128 MB in all;
    // config 2: 66-word windows, 64 MB), so that k_s1_fill does not expand the tile-row a fourth time
    int bm_stride = bmw1 < 128 ? bmw1 : 128;
    if (ntr > 0 && (size_t)ntr * bm_stride * 4 > ((size_t)128 << 20)) bm_stride = (int)((((size_t)128 << 20) / 4) / (size_t)ntr);
    const size_t bm_bytes = (size_t)ntr * bm_stride * 4;
    const bool bm_keep = ntr > 0 && bm_stride >= 8 && !(getenv("TSG_S1_KEEP_BITMAPS") && *getenv("TSG_S1_KEEP_BITMAPS") == '0');
    if (!arena_reserve(0, 9 * arena_need(nr, 4) + arena_need(nr, 1) + arena_need((size_t)B->tilem + 1, 4) + (bm_keep ? arena_need(bm_bytes, 1) : 0)))
        return last_error();
    int *w = arena_take<int>(0, nr), *jlo = arena_take<int>(0, nr), *jhi = arena_take<int>(0, nr);
    int *wptr = arena_take<int>(0, nr), *cnt = arena_take<int>(0, nr), *c_tile_ptr = arena_take<int>(0, nr);
    int *heavy_list = arena_take<int>(0, nr);
    uint8_t *light = arena_take<uint8_t>(0, nr);
    if (!w || !jlo || !jhi || !wptr || !cnt || !c_tile_ptr || !heavy_list || !light) return last_error();
    int *scal = (int *)c.d_scalars;                 // SC_* counters (8 ints)
    long long *tot = c.d_scalars + 4;               // [0] pairs, [1] C tiles (64-bit scan totals)
    CK(cudaMemsetAsync(scal, 0, 6 * sizeof(long long), c.stream));
    unsigned *bm_save = bm_keep ? (unsigned *)arena_take<uint8_t>(0, bm_bytes) : nullptr;
    if (bm_keep && !bm_save) return last_error();
    // tile-row templates (rowplans.cu), together with the recipe plans only: hash every tile-row's signature; if the rows
    // repeat, k_s1_count and k_s1_fill run on the representatives and the other rows are instantiated from them
    int nsig = 0;
    const int *rep_list = nullptr;
    int *sig_slot = arena_take<int>(0, nr), *rep_of = arena_take<int>(0, nr), *bclass = arena_take<int>(0, (size_t)B->tilem + 1);
    if (!sig_slot || !rep_of || !bclass) return last_error();
    // below ~8 K tile-rows the dozen extra launches and the extra read-back cost more than the expansions they save
    const int rowplans_min_rows = getenv("TSG_ROWPLANS_MIN_ROWS") ? atoi(getenv("TSG_ROWPLANS_MIN_ROWS")) : 8192;
    if (try_rowplans && ntr >= rowplans_min_rows && ntr >= 64 && plans_wanted(A, B)) {
        int rc0 = rowplans_signatures(A, B, trow0, ntr, w, sig_slot, rep_of, bclass, scal + SC_ERR, &rep_list, &nsig);
        if (rc0) return rc0;
    }
    const bool rowplans = nsig > 0;
    if (ntr > 0) {
        const size_t smem = (size_t)S1_WARPS * bmw1 * 4;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_s1_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_s1_count<<<ceil_div(rowplans ? nsig : ntr, S1_WARPS), S1_WARPS * 32, smem, c.stream>>>(
            trow0, ntr, bmw1, A->tile_ptr, A->tile_columnidx, A->tile_nnz, B->tile_ptr, B->tile_columnidx, w, jlo, jhi, cnt, light, heavy_list, scal,
            bm_save, bm_stride, rep_list, nsig);
        CK_LAUNCH();
        if (rowplans) {
            int rc0 = rowplans_expand_counts(ntr, rep_of, cnt, light);
            if (rc0) return rc0;
        }
    }
    // one scan per array: 32-bit offsets for the kernels, the 64-bit total for the host (slab planning keeps it < 2^31)
    int rc = exclusive_scan<int>(w, wptr, ntr, tot);
    if (!rc) rc = exclusive_scan<int>(cnt, c_tile_ptr, ntr, tot + 1);
    if (!rc) rc = publish_words(c.h_scalars, c.d_scalars, 12);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    int hs[8];
    memcpy(hs, (const void *)c.h_scalars, sizeof(hs));
    const int n_heavy = hs[SC_NHEAVY], n_light = hs[SC_NLIGHT];
    if (rowplans && n_heavy > 0) { *retry = true; return TSG_OK; }  // a heavy representative: templates are for light rows
    const long long pairs = c.h_scalars[4];
    if (hs[SC_ERR] || pairs >= (1ll << 31)) {
        set_error(TSG_ERR_OVERFLOW, "spgemm: %lld tile pairs in tile-rows [%d,%d) exceed 32-bit indexing; use smaller slabs", pairs, trow0, trow1);
        return last_error();
    }
    long long numblkC = c.h_scalars[5];
    size_t heavy_smem = 0;
    int heavy_grid = 0;
    unsigned *heavy_gscratch = nullptr;
    struct Scratch { unsigned *&p; ~Scratch() { if (p) dfree(p); } } heavy_guard{heavy_gscratch};  // released on every return path
    if (n_heavy > 0) {  // count the C tiles of the heavy tile-rows, scan again (one more read-back; R-MAT only)
        const int nw_max = hs[SC_NW_HEAVY];
        heavy_smem = ((size_t)nw_max + (size_t)nw_max / 8 + 2) * 4;
        const char *kb = getenv("TSG_S1_SMEM_KB");  // tests: a small budget forces the global-memory bitmap
        if (heavy_smem > (kb && *kb ? (size_t)atoi(kb) * 1024 : c.smem_optin)) {
            // the window does not fit shared memory: bitmap + prefix in a per-CTA slice of global memory, 2 CTAs per SM walk the list
            heavy_grid = n_heavy < 2 * c.num_sms ? n_heavy : 2 * c.num_sms;
            heavy_gscratch = dalloc_n<unsigned>((size_t)heavy_grid * (heavy_smem / 4));
            if (!heavy_gscratch) return last_error();
            heavy_smem = 0;
        } else {
            heavy_grid = n_heavy;
        }
        if (heavy_smem > 48 * 1024) {
            CK(cudaFuncSetAttribute(k_s1_heavy<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_smem));
            CK(cudaFuncSetAttribute(k_s1_heavy<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_smem));
        }
        k_s1_heavy<0><<<heavy_grid, S1_HEAVY_THREADS, heavy_smem, c.stream>>>(trow0, nw_max, n_heavy, heavy_gscratch, heavy_list, A->tile_ptr, A->tile_columnidx, B->tile_ptr,
                                                                          B->tile_columnidx, B->rm2csc, w, jlo, jhi, cnt, nullptr, nullptr,
                                                                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        CK_LAUNCH();
        rc = exclusive_scan<int>(cnt, c_tile_ptr, ntr, tot + 1);
        if (!rc) rc = read_back_i64(tot + 1, &numblkC);
        if (rc) return rc;
    }
    if (numblkC >= (1ll << 30)) {  // numblkC*16 must index uint16 arrays with int offsets
        set_error(TSG_ERR_OVERFLOW, "spgemm: %lld C tiles in tile-rows [%d,%d); use smaller slabs", numblkC, trow0, trow1);
        return last_error();
    }
    CK(cudaEventRecord(ev[1], c.stream));

    // C metadata allocation (slab 0)
    const size_t nb = (size_t)(numblkC > 0 ? numblkC : 1), np = (size_t)(pairs > 0 ? pairs : 1);
    {
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        size_t o_tp = take(((size_t)ntr + 1) * 4), o_tc = take(nb * 4), o_tr = take(nb * 4), o_tn = take((nb + 1) * 4);
        size_t o_p = take(nb * TS * 2), o_m = take(nb * TS * 2);
        size_t cap = 0;
        char *base = (char *)cslab_take(0, off, &cap);
        if (!base) return last_error();
        C->slab[0] = base; C->slab_bytes[0] = cap;
        C->cached = 1;  // slab[0], slab[1] go back to the library's C slab cache in tsg_tile_free
        C->tile_ptr = (int *)(base + o_tp); C->tile_columnidx = (int *)(base + o_tc); C->tile_rowidx = (int *)(base + o_tr);
        C->tile_nnz = (int *)(base + o_tn); C->ptr = (uint16_t *)(base + o_p); C->mask = (uint16_t *)(base + o_m);
    }
    C->n = B->n; C->tilem = ntr; C->tilen = B->tilen; C->numtile = (int)numblkC; C->col_major = 0; C->trow0 = trow0;
    {
        long long r0 = (long long)trow0 * TS, r1 = (long long)trow1 * TS;
        if (r1 > A->m) r1 = A->m;
        C->m = (int)(r1 > r0 ? r1 - r0 : 0);
    }
    rc = copy_words(C->tile_ptr, c_tile_ptr, (size_t)ntr + 1);
    if (rc) return rc;
    const bool heavy_rows = n_heavy > 0;  // tile-rows on the multi-warp path park one 16-byte record per pair
    if (!arena_reserve(1, 4 * arena_need(nb + 1, 4) + 4 * arena_need(np, 4) + arena_need(np, 2) +
                              (heavy_rows ? arena_need(np, 16) : 0) + numeric_scratch_bytes(ntr, numblkC)))
        return last_error();
    int *pair_ptr = arena_take<int>(1, nb + 1), *pair_end = arena_take<int>(1, nb + 1), *pair_a = arena_take<int>(1, np), *pair_b = arena_take<int>(1, np);
    uint16_t *pair_slot = arena_take<uint16_t>(1, np);
    int4 *pair_tmp = heavy_rows ? arena_take<int4>(1, np) : nullptr;
    NumericBufs nbufs{arena_take<uint8_t>(1, (size_t)ntr + 1), arena_take<int>(1, nb), 0, 0};
    // recipe plans (plans.cu): attempted when both operands are made of few distinct tile patterns and no tile-row is heavy
    bool plans_on = plans_wanted(A, B) && !heavy_rows && numblkC > 0 && pairs > 0;
    int *rslot = plans_on ? arena_take<int>(1, nb) : nullptr, *recipe_id = plans_on ? arena_take<int>(1, nb) : nullptr;
    unsigned *pair_pat = plans_on ? arena_take<unsigned>(1, np) : nullptr;  // (A pattern, B pattern) per pair
    unsigned *pair_src = rowplans ? arena_take<unsigned>(1, np) : nullptr;  // tile-row templates: where every pair of a representative came from
    if (!pair_ptr || !pair_end || !pair_a || !pair_b || !pair_slot || (heavy_rows && !pair_tmp) || !nbufs.row_kind || !nbufs.dense_list ||
        (plans_on && (!rslot || !recipe_id || !pair_pat)) || (rowplans && !pair_src))
        return last_error();
    if (heavy_rows) CK(cudaMemsetAsync(pair_end, 0, nb * 4, c.stream));  // the heavy kernel counts with atomics
    if (rowplans && !plans_on) { *retry = true; return TSG_OK; }
    if (rowplans) CK(cudaMemsetAsync(rslot, 0xFF, nb * 4, c.stream));  // -1: "this tile's row is not a representative"
    CK(cudaEventRecord(ev[2], c.stream));  // [1..2] = allocation

    // ---------------- step 1b (+ fused step 2): tile columns, pair lists, C masks ----------------
    // the light path also produces C's masks / Ptr / tile nnz (fused step 2); TSG_FUSE=0 disables the fusion (A/B runs).
    // Worthwhile only when B's tile-rows are long enough to fill the lanes (>= 4 tiles per tile-row on average;
    // block-FEM has 3 and is faster through k_step2).
    static const int fuse_env = getenv("TSG_FUSE") ? atoi(getenv("TSG_FUSE")) : -1;
    bool fused = false;
    PlanTable ptab{nullptr, nullptr, nullptr, nullptr};
    if (plans_on) {
        rc = plans_begin(&ptab);
        if (rc) return rc;
    }
    if (numblkC > 0 && n_light > 0) {
        const int bmw = (hs[SC_NW_LIGHT] + 1) & ~1, nj = hs[SC_MAXJ], njpad = nj | 1;  // odd row stride: fewer bank conflicts
        const bool want = !plans_on && (fuse_env >= 0 ? fuse_env != 0 : (long long)B->numtile >= 4ll * B->tilem);
        int warp_words = bmw + bmw / 2 + nj + (TS * 32) / 2 + (TS * njpad + 1) / 2;
        fused = want && (size_t)S1_WARPS * warp_words * 4 <= c.smem_optin;
        const int hoff = (bmw + bmw / 2 + nj + 1) & ~1;
        if (!fused) warp_words = plans_on ? hoff + 2 * nj : bmw + bmw / 2 + nj;
        warp_words = (warp_words + 1) & ~1;
        S1Fill P{trow0, ntr, bmw, nj, njpad, warp_words, hoff, A->tile_ptr, A->tile_columnidx, B->tile_ptr, B->tile_columnidx, B->rm2csc,
                 jlo, jhi, wptr, C->tile_ptr, light, C->tile_columnidx, C->tile_rowidx, pair_ptr, pair_end, pair_a, pair_b, pair_slot,
                 A->mask, B->mask, C->ptr, C->mask, C->tile_nnz, A->pat, B->pat, ptab, rslot, pair_pat, bm_save, bm_stride, rep_list, nsig, pair_src};
        const size_t smem = (size_t)S1_WARPS * warp_words * 4;
        if (smem > c.smem_optin) { set_error(TSG_ERR_UNSUPPORTED, "step 1: %zu B of shared memory per CTA needed (> %zu)", smem, c.smem_optin); return last_error(); }
        const int blocks = ceil_div(rowplans ? nsig : ntr, S1_WARPS);
        if (fused) {
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_s1_fill<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_s1_fill<true, false><<<blocks, S1_WARPS * 32, smem, c.stream>>>(P);
        } else if (plans_on) {
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_s1_fill<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_s1_fill<false, true><<<blocks, S1_WARPS * 32, smem, c.stream>>>(P);
        } else {
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_s1_fill<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_s1_fill<false, false><<<blocks, S1_WARPS * 32, smem, c.stream>>>(P);
        }
        CK_LAUNCH();
    }
    if (numblkC > 0 && n_heavy > 0) {
        k_s1_heavy<1><<<heavy_grid, S1_HEAVY_THREADS, heavy_smem, c.stream>>>(trow0, hs[SC_NW_HEAVY], n_heavy, heavy_gscratch, heavy_list, A->tile_ptr, A->tile_columnidx,
                                                                          B->tile_ptr, B->tile_columnidx, B->rm2csc, w, jlo, jhi, nullptr,
                                                                          C->tile_ptr, wptr, C->tile_columnidx, C->tile_rowidx, pair_ptr,
                                                                          pair_end, pair_a, pair_b, pair_tmp);
        CK_LAUNCH();
    }

    // ---------------- step 2 ----------------
    CK(cudaEventRecord(ev_s2, c.stream));
    const PairLists plists{pair_ptr, pair_end, pair_a, pair_b, pair_slot};
    const int *d_plan_fail = nullptr;
    // pair-based symbolic for the C tiles the fused step-1 path did not cover: half-warp per tile, or thread per
    // tile when the tiles are hypersparse (<= 2 pairs per C tile and <= 2 entries per A tile on average)
    auto generic_symbolic = [&](bool skip_light) -> int {
        const uint8_t *lf = skip_light ? light : nullptr;
        const bool hypersparse = pairs <= 2 * numblkC && A->nnz <= 2ll * A->numtile;
        if (hypersparse)
            k_step2_thread<<<ceil_div(numblkC, S2T_THREADS), S2T_THREADS, 0, c.stream>>>((int)numblkC, pair_ptr, pair_end, pair_a, pair_b,
                                                                                         A->tile_nnz, A->col, B->mask, C->ptr, C->mask,
                                                                                         C->tile_nnz, C->tile_rowidx, lf, trow0);
        else
            k_step2<<<ceil_div(numblkC * 16, 128), 128, 0, c.stream>>>((int)numblkC, pair_ptr, pair_end, pair_a, pair_b, A->mask, B->mask,
                                                                       C->ptr, C->mask, C->tile_nnz, C->tile_rowidx, lf, trow0);
        CK_LAUNCH();
        return TSG_OK;
    };
    const RowTemplates rt{nsig, trow0, ntr, rep_list, rep_of, w, wptr, bclass, pair_ptr, pair_end, pair_a, pair_b, pair_src};
    if (plans_on) {  // C's masks / Ptr / tile nnz from the recipe plans (or nothing useful, if the fail flag comes up)
        rc = plans_symbolic_device(A, B, C, plists, pair_pat, rslot, recipe_id, rowplans ? &rt : nullptr, &d_plan_fail);
        if (rc) return rc;
    } else if (numblkC > 0 && (!fused || n_heavy > 0)) {
        rc = generic_symbolic(fused);
        if (rc) return rc;
    }
    rc = exclusive_scan<int>(C->tile_nnz, C->tile_nnz, numblkC, tot);
    if (rc) return rc;
    // pick the accumulator per C tile-row / tile (numeric.cu); its counters come back with nnz(C) in one read-back
    int *d_ns = (int *)(c.d_scalars + 16);
    long long nnzC = 0;
    int plan_recipes = 0;
    int *d_tau = nullptr;              // row plans (plans.cu): template index of every representative tile-row
    const int *d_rowplan_fail = nullptr;
    struct TauGuard { int *&p; ~TauGuard() { if (p) dfree(p); } } tau_guard{d_tau};
    if (plans_on) {
        if (rowplans) {  // the slots of every template's tile-row laid out warp by warp for the numeric kernel
            rc = plans_rowplan_build(C, rt, recipe_id, pair_ptr, &d_tau, &d_rowplan_fail);
            if (rc) return rc;
            if (d_rowplan_fail) rc = publish_words(&c.h_scalars[28], d_rowplan_fail, 1);
            if (rc) return rc;
        }
        rc = publish_words(&c.h_scalars[20], d_plan_fail, 1);
        if (!rc) rc = publish_words(&c.h_scalars[21], plans_recipe_count_ptr(), 1);
        if (!rc && rowplans) rc = publish_words(&c.h_scalars[25], rowplans_fail_ptr(), 1);
        if (!rc) rc = read_back_i64(tot, &nnzC);
        if (rc) return rc;
        if (rowplans && *(const volatile int *)&c.h_scalars[25]) { *retry = true; return TSG_OK; }  // a signature collision
        if (*(const volatile int *)&c.h_scalars[20]) {  // a collision or too many recipes: the generic kernels run instead
            if (getenv("TSG_DEBUG"))
                fprintf(stderr, "tsg: recipe plans failed with code %d (%d recipes, %d row templates): generic kernels\n",
                        *(const volatile int *)&c.h_scalars[20], *(const volatile int *)&c.h_scalars[21], nsig);
            plans_on = false;
            plan_recipes = -1;
            k_pair_slots<<<ceil_div(ntr, S1_WARPS), S1_WARPS * 32, 0, c.stream>>>(trow0, ntr, light, A->tile_ptr, A->tile_columnidx, B->tile_ptr,
                                                                                  B->tile_columnidx, C->tile_ptr, C->tile_columnidx, wptr, pair_slot);
            CK_LAUNCH();
            rc = generic_symbolic(false);
            if (!rc) rc = exclusive_scan<int>(C->tile_nnz, C->tile_nnz, numblkC, tot);
            if (rc) return rc;
        } else {
            plan_recipes = *(const volatile int *)&c.h_scalars[21];
        }
    }
    if (!plans_on) {
        rc = numeric_classify_device(A, C, trow0, ntr, wptr, light, &nbufs, d_ns);
        if (!rc) rc = publish_words(&c.h_scalars[16], d_ns, 8);
        if (!rc) rc = read_back_i64(tot, &nnzC);
        if (rc) return rc;
    }
    if (nnzC >= (1ll << 31)) {
        set_error(TSG_ERR_OVERFLOW, "spgemm: nnz(C) = %lld in tile-rows [%d,%d) exceeds int32; use smaller slabs", nnzC, trow0, trow1);
        return last_error();
    }
    int h_ns[8];
    memcpy(h_ns, (const void *)&c.h_scalars[16], sizeof(h_ns));
    CK(cudaEventRecord(ev[3], c.stream));

    // ---------------- step 3 ----------------
    {
        size_t nz = (size_t)(nnzC > 0 ? nnzC : 1);
        size_t o_c = (nz * 8 + 255) & ~(size_t)255;
        size_t cap = 0;
        char *base = (char *)cslab_take(1, o_c + nz * 2, &cap);
        if (!base) return last_error();
        C->slab[1] = base; C->slab_bytes[1] = cap;
        C->val = (double *)base; C->col = (uint16_t *)(base + o_c);
        C->nnz = nnzC;
    }
    CK(cudaEventRecord(ev_s3, c.stream));
    tsg_stats nst;
    memset(&nst, 0, sizeof(nst));
    if (plans_on) {
        // the largest tile-row: A values, C tiles and pairs as k_s1_count saw them (every tile-row is light on this path)
        const size_t need = plans_rows_need_bound(hs[SC_MAXNNZA], hs[SC_MAXJ], hs[SC_WMAX]);
        const bool rowplan_ok = d_tau && d_rowplan_fail && *(const volatile int *)&c.h_scalars[28] == 0;
        rc = plans_numeric_device(A, B, C, plists, recipe_id, trow0, ntr, wptr, need > (1u << 30) ? (1 << 30) : (int)need, &nst,
                                  rowplan_ok ? &rt : nullptr, rowplan_ok ? d_tau : nullptr, hs[SC_MAXNNZA], hs[SC_WMAX]);
    }
    else rc = numeric_device(A, B, C, trow0, ntr, wptr, plists, nbufs, h_ns, heavy_rows, &nst);
    if (rc) return rc;
    CK(cudaEventRecord(ev[4], c.stream));
    if (stats && ntr != A->tilem) {  // the slab's share of A (tiles, nonzeros) for the byte count below
        k_slab_extent<<<1, 32, 0, c.stream>>>(A->tile_ptr, A->tile_nnz, trow0, trow1,
                                              (int *)((char *)c.h_scalars_dev + 10 * sizeof(long long)));
        CK_LAUNCH();
    }
    CK(cudaStreamSynchronize(c.stream));

    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float f;
        stats->numblkC = numblkC; stats->nnzC = nnzC; stats->pairs = pairs;
        float s1a, al, s1b, s2, al2, s3, tot;
        CK(cudaEventElapsedTime(&s1a, ev[0], ev[1]));
        CK(cudaEventElapsedTime(&al, ev[1], ev[2]));
        CK(cudaEventElapsedTime(&s1b, ev[2], ev_s2));
        CK(cudaEventElapsedTime(&s2, ev_s2, ev[3]));
        CK(cudaEventElapsedTime(&al2, ev[3], ev_s3));
        CK(cudaEventElapsedTime(&s3, ev_s3, ev[4]));
        CK(cudaEventElapsedTime(&tot, ev[0], ev[4]));
        (void)f;
        stats->ms_step1 = s1a + s1b; stats->ms_step2 = s2; stats->ms_step3 = s3; stats->ms_alloc = al + al2; stats->ms_total = tot;
        stats->launches = (int)(c.launches - launches0);
        stats->rows_staged = nst.rows_staged; stats->rows_gather = nst.rows_gather; stats->tiles_dense = nst.tiles_dense;
        stats->rows_smem = nst.rows_smem; stats->tiles_nonempty = nst.tiles_nonempty;
        stats->plan_recipes = plan_recipes;
        stats->row_templates = rowplans ? nsig : 0;
        // algorithmic bytes, SURVEY.md 8(d). A's share is the slab's tiles; B is read whole.
        long long a_tiles = A->numtile, a_nnz = A->nnz;
        if (ntr != A->tilem) {
            const int *h = (const int *)&c.h_scalars[10];
            a_tiles = h[1] - h[0]; a_nnz = h[3] - h[2];
        }
        long long bytesA = a_nnz * 10 + a_tiles * 72 + ((long long)ntr + 1) * 4;
        long long bytesB = B->nnz * 10 + (long long)B->numtile * 72 + ((long long)B->tilem + 1) * 4 + (long long)B->numtile * 4 +
                           ((long long)B->tilen + 1) * 4;
        long long bytesC = nnzC * 10 + numblkC * 76 + ((long long)ntr + 1) * 4;
        stats->algorithmic_bytes = bytesA + bytesB + bytesC;
    }
    return TSG_OK;
}

}  // namespace tsg
