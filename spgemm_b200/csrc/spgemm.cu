// spgemm.cu -- TileSpGEMM steps 1-3 on B200, for a slab [trow0, trow1) of C tile-rows.
// Replaces reference src/tilespgemm-cuda.h:2220-2844 (host) and its kernels:
//   step 1  tile_spgemm_step1_cuda_spa_kernel / _numeric_ (:279-392) and the nsparse hash path
//           (src/spgemm_nsparse_kernel.h:221-311,1171-1438)
//   step 2  tile_spgemm_step3_cuda_kernel_2level_halfwarp (:394-773)
//   step 3  tile_spgemm_step4_cuda_kernel_smem_v3[_halfwarp] (:1273-1952)
//
// What is different from the reference (same results):
//   * Step 1 is a Gustavson expansion at tile level over a WINDOWED shared-memory bitmap
//     [Jmin, Jmax] of the tile-row (the reference's bitmap spans all tile columns and only exists
//     for tilen <= 16384, else it falls back to a hash path). Besides C's tile list it emits, per C
//     tile, the list of matched (A tile, B tile) pairs, so the later steps never intersect index
//     lists (the reference re-intersects A's tile-row with B's tile-column by binary search in both
//     of its later steps and caches at most one pair, :538-547).
//   * Step 2 (C's row masks, Ptr, tile nnz) is FUSED into step 1 on the one-warp path: while the warp
//     enumerates the B tiles paired with one A tile, A's row masks are warp-uniform, so the 16x16x16
//     boolean product of up to 32 pairs costs one shared-memory load + OR per A entry. Tile-rows on
//     the multi-warp path (and matrices whose B tile-rows are too short to fill a warp) use k_step2
//     (half-warp per C tile, lane r ORs B's row masks, fetched by shuffle, selected by A's row mask r) or,
//     for hypersparse tiles, k_step2_thread (one thread per C tile).
//   * Step 3 has two kernels, chosen per call from the average fill of A's tiles: a GATHER (one lane per C
//     nonzero, register accumulation in the serial SPA's summation order, no accumulator memory, coalesced
//     stores) for sparse tiles, and a DENSE ACCUMULATOR (warp per C tile, 8 register accumulators per lane,
//     the B tile expanded in shared memory) for well-filled tiles (block-FEM). Neither uses atomics; the
//     reference does one global atomicAdd per product plus a binary search (:1450,1558,1795,1900).
//     k_step3_dmma is the FP64 tensor-core (mma.sync m8n8k4) variant of the dense kernel, opt-in.
//   * Empty C tiles are kept with Ptr = mask = 0 and nnz 0 (the reference leaves them
//     uninitialised).
//   * Scratch lives in grow-only arenas; sizes are read back three times (pairs/window, numblkC, nnzC).
// Superseded kernels measured on the way (half-warp-per-tile numeric with a shared-memory accumulator,
// "rounds" numeric, flattened / half-warp-per-pair symbolic, tile-row-per-warp numeric, prefetching variants)
// are described in profiles/README.md; some are kept as text under scratch/.
#include "common.cuh"
#include "scan.cuh"
#include "kernels.h"

namespace tsg {

constexpr int S1_LIGHT_MAX = 2048;   // tile-rows with <= this many pairs run on one warp (deterministic pair order)
constexpr int S1_HEAVY_THREADS = 1024;  // one CTA per heavy tile-row: its window bitmap can take most of the SM's shared memory
constexpr int S1_SORT_MAX = 64;      // heavy path: pair lists up to this length are re-sorted by A tile

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
// Step 1b/1c: one CTA per C tile-row. MODE 0 counts the distinct tile columns. MODE 1 emits the
// sorted tile-column list, and per C tile the matched (A tile, B tile) pair list.
// THREADS = 32 handles rows with w in (0, S1_LIGHT_MAX]; THREADS = S1_HEAVY_THREADS the heavier ones.
// Dynamic smem: bitmap[nw_max] | pre8[nw_max/8 + 1] | (fused symbolic) bmT[8][32] u32 | cm[16][numJ_pad] u16.
//
// Fused bitmask symbolic (step 2) on the one-warp path: while the warp enumerates the pairs of A tile
// (I,K) with the B tiles of tile-row K (one B tile per lane), A's 16 row masks are WARP-UNIFORM, so the
// 16x16x16 boolean product of the pair costs one shared-memory load + OR per A entry for all <= 32
// pairs at once, instead of a per-pair pass over a half-warp (k_step2). The C row masks of the whole
// tile-row live in shared memory as cm[row][slot]; Ptr / mask / tile nnz are written at the end.
// ---------------------------------------------------------------------------------------------
struct S1Fuse {
    const uint16_t *a_mask, *b_mask;
    uint16_t *c_ptr, *c_mask;
    int *c_cnt;
    int numJ_pad;  // 0: not fused (k_step2 computes the masks from the pair lists)
    int4 *tmp;     // multi-warp path: one (slot, rank in slot, A tile, B tile) record per pair of the slab, or nullptr
};

template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS)
k_step1(int trow0, int nw_max, int wmin, int wmax, const int *__restrict__ a_tile_ptr,
        const int *__restrict__ a_tile_col, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col,
        const int *__restrict__ b_rm2csc, const int *__restrict__ w, const int *__restrict__ jlo,
        const int *__restrict__ jhi, int *__restrict__ cnt /*MODE0 out*/, const int *__restrict__ c_tile_ptr,
        const int *__restrict__ wptr, int *__restrict__ c_tile_col, int *__restrict__ c_tile_row,
        int *__restrict__ pair_ptr, int *__restrict__ pair_end, int *__restrict__ pair_a, int *__restrict__ pair_b,
        int *__restrict__ maxJ /*MODE0: max tile count over one-warp rows*/, S1Fuse fz)
{
    extern __shared__ unsigned s1_smem[];
    __shared__ int s_warp[THREADS / 32];
    __shared__ int s_carry;
    unsigned *bitmap = s1_smem;
    int *pre8 = (int *)(s1_smem + nw_max);
    const int i = blockIdx.x, I = trow0 + i;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = THREADS / 32;
    const int wi = w[i];
    if (wi <= wmin || wi > wmax) return;  // other launch's row (or nothing to do: MODE 0 output is pre-zeroed)
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
        if (tid == 0) {
            cnt[i] = total;
            if (THREADS == 32) atomicMax(maxJ, total);
        }
        return;
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
    const bool fuse = THREADS == 32 && fz.numJ_pad > 0;
    unsigned *bmT = s1_smem + nw_max + nw_max / 8 + 2;                    // [8][32]: B row masks 2j, 2j+1 of lane's tile
    uint16_t *cm = reinterpret_cast<uint16_t *>(bmT + 8 * 32);            // [16][numJ_pad]
    if (fuse) {
        for (int k = tid; k < (TS * fz.numJ_pad + 1) / 2; k += THREADS) reinterpret_cast<unsigned *>(cm)[k] = 0;
    }
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
    if (THREADS > 32 && fz.tmp) {
        // Multi-warp path, ONE expansion and ONE global atomic per pair: the atomicAdd that counts the pairs of a C tile
        // also returns the pair's rank inside the tile's list; (slot, rank, A tile, B tile) is parked in a per-row
        // scratch record, the counts are scanned, and the records are then placed -- without walking B's structure,
        // recomputing slots or touching the counters a second time (global atomics bound this path on R-MAT).
        if (tid == 0) s_carry = 0;
        __syncthreads();
        const unsigned lt = (1u << lane) - 1;
        int4 *tmp = fz.tmp + wptr[i];
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
        for (int sidx = tid; sidx < numJ; sidx += THREADS) {  // arrival order -> ascending A tile for short lists
            int b = pair_ptr[cbase + sidx], e = pair_end[cbase + sidx], len = e - b;
            if (len > 1 && len <= S1_SORT_MAX) {
                for (int x = b + 1; x < e; x++) {
                    int ka = pair_a[x], kb = pair_b[x], y = x - 1;
                    while (y >= b && pair_a[y] > ka) { pair_a[y + 1] = pair_a[y]; pair_b[y + 1] = pair_b[y]; y--; }
                    pair_a[y + 1] = ka; pair_b[y + 1] = kb;
                }
            }
        }
        return;
    }
    // pair counts per C tile (pair_end is zero on entry)
    for (int ta = a0 + warp; ta < a1; ta += NWARPS) {
        int K = a_tile_col[ta];
        for (int tb = b_tile_ptr[K] + lane; tb < b_tile_ptr[K + 1]; tb += 32) {
            int slot = s1_rank(bitmap, pre8, b_tile_col[tb] - lo);
            if (THREADS == 32) pair_end[cbase + slot]++;  // one A tile at a time, distinct slots per lane
            else atomicAdd(&pair_end[cbase + slot], 1);
        }
        if (THREADS == 32) __syncwarp();
    }
    __syncthreads();
    // exclusive scan of the counts over the tile-row's slots -> pair_ptr; pair_end becomes the cursor
    if (tid == 0) s_carry = wptr[i];
    __syncthreads();
    for (int s0 = 0; s0 < numJ; s0 += THREADS) {
        int s = s0 + tid;
        int v = s < numJ ? pair_end[cbase + s] : 0, total;
        int ex = block_excl_scan<THREADS>(v, s_warp, &total);
        int carry = s_carry;
        if (s < numJ) { pair_ptr[cbase + s] = carry + ex; pair_end[cbase + s] = carry + ex; }
        __syncthreads();
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    // write the pairs (and, fused, OR the pair's boolean product into the tile-row's C masks)
    for (int ta = a0 + warp; ta < a1; ta += NWARPS) {
        int K = a_tile_col[ta];
        unsigned amw[8];
        if (fuse) {  // A's 16 row masks: one 32-byte line, the same for every lane
            const uint4 *ap = reinterpret_cast<const uint4 *>(fz.a_mask + (size_t)ta * TS);
            const uint4 x = ap[0], y = ap[1];
            amw[0] = x.x; amw[1] = x.y; amw[2] = x.z; amw[3] = x.w; amw[4] = y.x; amw[5] = y.y; amw[6] = y.z; amw[7] = y.w;
        }
        for (int tb = b_tile_ptr[K] + lane; tb < b_tile_ptr[K + 1]; tb += 32) {
            int slot = s1_rank(bitmap, pre8, b_tile_col[tb] - lo);
            int pos;
            if (THREADS == 32) pos = pair_end[cbase + slot]++;
            else pos = atomicAdd(&pair_end[cbase + slot], 1);
            const int b = b_rm2csc[tb];
            pair_a[pos] = ta;
            pair_b[pos] = b;
            if (fuse) {
                const uint4 *bp = reinterpret_cast<const uint4 *>(fz.b_mask + (size_t)b * TS);
                const uint4 x = bp[0], y = bp[1];
                bmT[0 * 32 + lane] = x.x; bmT[1 * 32 + lane] = x.y; bmT[2 * 32 + lane] = x.z; bmT[3 * 32 + lane] = x.w;
                bmT[4 * 32 + lane] = y.x; bmT[5 * 32 + lane] = y.y; bmT[6 * 32 + lane] = y.z; bmT[7 * 32 + lane] = y.w;
#pragma unroll
                for (int r = 0; r < TS; r++) {
                    unsigned m = (r & 1) ? (amw[r >> 1] >> 16) : (amw[r >> 1] & 0xFFFFu);  // little-endian u16 pairs
                    if (m) {                                                            // warp-uniform
                        unsigned acc = 0;
                        do {
                            const int k = __clz(m) - 16;
                            const unsigned wd = bmT[(k >> 1) * 32 + lane];
                            acc |= (k & 1) ? (wd >> 16) : (wd & 0xFFFFu);
                            m &= ~(0x8000u >> k);
                        } while (m);
                        cm[r * fz.numJ_pad + slot] |= (uint16_t)acc;  // distinct slots per lane
                    }
                }
            }
        }
        if (THREADS == 32) __syncwarp();
    }
    if (fuse) {
        __syncwarp();
        for (int sl = lane; sl < numJ; sl += 32) {  // Ptr (exclusive row offsets), mask and nnz of each C tile
            unsigned pw[8], mw[8];
            int run = 0;
#pragma unroll
            for (int r = 0; r < TS; r += 2) {
                const unsigned m0 = cm[r * fz.numJ_pad + sl], m1 = cm[(r + 1) * fz.numJ_pad + sl];
                const int p0 = run, p1 = run + __popc(m0);
                run = p1 + __popc(m1);
                pw[r >> 1] = (unsigned)p0 | ((unsigned)p1 << 16);
                mw[r >> 1] = m0 | (m1 << 16);
            }
            uint4 *dp = reinterpret_cast<uint4 *>(fz.c_ptr + (size_t)(cbase + sl) * TS);
            uint4 *dm = reinterpret_cast<uint4 *>(fz.c_mask + (size_t)(cbase + sl) * TS);
            dp[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]); dp[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
            dm[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]); dm[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
            fz.c_cnt[cbase + sl] = run;
        }
    }
    if (THREADS > 32) {
        // several warps appended concurrently: restore ascending-A-tile order for short lists so the
        // FP64 summation order is reproducible (lists longer than S1_SORT_MAX keep arrival order).
        __syncthreads();
        for (int s = tid; s < numJ; s += THREADS) {
            int b = pair_ptr[cbase + s], e = pair_end[cbase + s], len = e - b;
            if (len > 1 && len <= S1_SORT_MAX) {
                for (int x = b + 1; x < e; x++) {
                    int ka = pair_a[x], kb = pair_b[x], y = x - 1;
                    while (y >= b && pair_a[y] > ka) { pair_a[y + 1] = pair_a[y]; pair_b[y + 1] = pair_b[y]; y--; }
                    pair_a[y + 1] = ka; pair_b[y + 1] = kb;
                }
            }
        }
    }
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
        int *__restrict__ c_cnt, const int *__restrict__ c_tile_row, const int *__restrict__ w, int trow0, int light_max)
{
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int l16 = threadIdx.x & 15;
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
    if (t >= numblkC) return;
    if (w && w[c_tile_row[t] - trow0] <= light_max) return;  // masks already produced by the fused step-1 path
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
               const int *__restrict__ w, int trow0, int light_max)
{
    __shared__ uint16_t cm[TS][S2T_THREADS];
    const int t = blockIdx.x * S2T_THREADS + threadIdx.x, tid = threadIdx.x;
    if (t >= numblkC) return;
    if (w && w[c_tile_row[t] - trow0] <= light_max) return;  // masks already produced by the fused step-1 path
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

template <int MODE>
static int launch_step1(int ntr, int trow0, int nw_max, int wmax_seen, const tsg_dtile *A, const tsg_dtile *B, const int *w,
                        const int *jlo, const int *jhi, int *cnt, const int *c_tile_ptr, const int *wptr, int *c_tile_col,
                        int *c_tile_row, int *pair_ptr, int *pair_end, int *pair_a, int *pair_b, int *maxJ, S1Fuse fz)
{
    Ctx &c = ctx();
    size_t smem = ((size_t)nw_max + (size_t)nw_max / 8 + 2) * 4;
    if (fz.numJ_pad > 0) smem += 8 * 32 * 4 + (size_t)fz.numJ_pad * TS * 2 + 4;
    if (smem > c.smem_optin) {
        set_error(TSG_ERR_UNSUPPORTED, "step 1: tile-column window of %d words needs %zu B of shared memory (> %zu)", nw_max, smem, c.smem_optin);
        return last_error();
    }
    if (smem > 48 * 1024) {
        CK(cudaFuncSetAttribute(k_step1<32, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_step1<S1_HEAVY_THREADS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    k_step1<32, MODE><<<ntr, 32, smem, c.stream>>>(trow0, nw_max, 0, S1_LIGHT_MAX, A->tile_ptr, A->tile_columnidx, B->tile_ptr,
                                                   B->tile_columnidx, B->rm2csc, w, jlo, jhi, cnt, c_tile_ptr, wptr, c_tile_col,
                                                   c_tile_row, pair_ptr, pair_end, pair_a, pair_b, maxJ, fz);
    CK_LAUNCH();
    if (wmax_seen > S1_LIGHT_MAX) {
        k_step1<S1_HEAVY_THREADS, MODE><<<ntr, S1_HEAVY_THREADS, smem, c.stream>>>(
            trow0, nw_max, S1_LIGHT_MAX, 0x7fffffff, A->tile_ptr, A->tile_columnidx, B->tile_ptr, B->tile_columnidx, B->rm2csc, w,
            jlo, jhi, cnt, c_tile_ptr, wptr, c_tile_col, c_tile_row, pair_ptr, pair_end, pair_a, pair_b, maxJ, fz);
        CK_LAUNCH();
    }
    return TSG_OK;
}

int spgemm_device(const tsg_dtile *A, const tsg_dtile *B, int trow0, int trow1, tsg_dtile *C, tsg_stats *stats)
{
    Ctx &c = ctx();
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

    // ---------------- step 1 ----------------
    if (!arena_reserve(0, 5 * arena_need((size_t)ntr + 1, 4) + 2 * arena_need((size_t)ntr + 1, 8))) return last_error();
    int *w = arena_take<int>(0, (size_t)ntr + 1), *jlo = arena_take<int>(0, (size_t)ntr + 1), *jhi = arena_take<int>(0, (size_t)ntr + 1);
    int *wptr = arena_take<int>(0, (size_t)ntr + 1);
    int *c_tile_ptr = arena_take<int>(0, (size_t)ntr + 1);
    long long *wptr64 = arena_take<long long>(0, (size_t)ntr + 1), *numblk64 = arena_take<long long>(0, (size_t)ntr + 1);
    if (!w || !jlo || !jhi || !wptr || !c_tile_ptr || !wptr64 || !numblk64) return last_error();
    int *scal = (int *)c.d_scalars;
    CK(cudaMemsetAsync(scal, 0, 4 * sizeof(int), c.stream));
    CK(cudaMemsetAsync(c_tile_ptr, 0, ((size_t)ntr + 1) * sizeof(int), c.stream));
    if (ntr > 0) {
        k_step1_weights<<<ceil_div((long long)ntr * 32, 128), 128, 0, c.stream>>>(trow0, ntr, A->tile_ptr, A->tile_columnidx, B->tile_ptr,
                                                                                 B->tile_columnidx, w, jlo, jhi, scal);
        CK_LAUNCH();
    }
    long long *wtot = c.d_scalars + 4;
    // 64-bit scan output for the total, 32-bit offsets for the kernels (slab planning keeps it < 2^31)
    int rc = exclusive_scan<long long>(w, wptr64, ntr);
    if (rc) return rc;
    rc = exclusive_scan<int>(w, wptr, ntr);
    if (rc) return rc;
    rc = copy_words(wtot, wptr64 + ntr, 2);
    if (rc) return rc;
    rc = publish_words(c.h_scalars, c.d_scalars, 12);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    const int *hs = (const int *)c.h_scalars;
    const int nw_max = hs[0] > 0 ? hs[0] : 1, werr = hs[1], wmax_seen = hs[2];
    const long long pairs = c.h_scalars[4];
    if (werr || pairs >= (1ll << 31)) {
        set_error(TSG_ERR_OVERFLOW, "spgemm: %lld tile pairs in tile-rows [%d,%d) exceed 32-bit indexing; use smaller slabs", pairs, trow0, trow1);
        return last_error();
    }
    if (ntr > 0 && pairs > 0) {
        rc = launch_step1<0>(ntr, trow0, nw_max, wmax_seen, A, B, w, jlo, jhi, c_tile_ptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, scal + 3, S1Fuse{nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr});
        if (rc) return rc;
    }
    rc = exclusive_scan<long long>(c_tile_ptr, numblk64, ntr);
    if (rc) return rc;
    rc = exclusive_scan<int>(c_tile_ptr, c_tile_ptr, ntr);
    if (rc) return rc;
    long long numblkC = 0;
    rc = publish_words(&c.h_scalars[8], numblk64 + ntr, 2);
    if (!rc) rc = publish_words(&c.h_scalars[9], scal + 3, 1);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    numblkC = c.h_scalars[8];
    const int maxJ_light = *(const int *)&c.h_scalars[9];
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
    const bool heavy_rows = wmax_seen > S1_LIGHT_MAX;  // tile-rows on the multi-warp path park one 16-byte record per pair
    if (!arena_reserve(1, 2 * arena_need(nb + 1, 4) + 2 * arena_need(np, 4) + arena_need(nb + 1, 8) + (heavy_rows ? arena_need(np, 16) : 0) +
                              numeric_scratch_bytes(ntr, numblkC)))
        return last_error();
    int *pair_ptr = arena_take<int>(1, nb + 1), *pair_end = arena_take<int>(1, nb + 1), *pair_a = arena_take<int>(1, np), *pair_b = arena_take<int>(1, np);
    long long *nnz64 = arena_take<long long>(1, nb + 1);
    int4 *pair_tmp = heavy_rows ? arena_take<int4>(1, np) : nullptr;
    NumericBufs nbufs{arena_take<uint8_t>(1, (size_t)ntr + 1), arena_take<int>(1, nb), 0, 0};
    if (!pair_ptr || !pair_end || !pair_a || !pair_b || !nnz64 || (heavy_rows && !pair_tmp) || !nbufs.row_kind || !nbufs.dense_list)
        return last_error();
    CK(cudaMemsetAsync(pair_end, 0, nb * 4, c.stream));
    CK(cudaEventRecord(ev[2], c.stream));  // [1..2] = allocation
    // the one-warp step-1 path also produces C's masks / Ptr / tile nnz (fused step 2) when the tile-row's
    // masks fit shared memory; TSG_FUSE=0 disables the fusion (A/B measurements)
    static const int fuse_env = getenv("TSG_FUSE") ? atoi(getenv("TSG_FUSE")) : -1;
    S1Fuse fz{A->mask, B->mask, C->ptr, C->mask, C->tile_nnz, 0, pair_tmp};
    {
        size_t need = ((size_t)nw_max + (size_t)nw_max / 8 + 2) * 4 + 8 * 32 * 4 + (size_t)(maxJ_light + 1) * TS * 2 + 4;
        // worthwhile only when B's tile-rows are long enough to fill the lanes (>= 4 tiles per tile-row on average;
        // block-FEM has 3 and is faster through k_step2)
        const bool want = fuse_env >= 0 ? fuse_env != 0 : (long long)B->numtile >= 4ll * B->tilem;
        if (want && maxJ_light > 0 && need <= c.smem_optin) fz.numJ_pad = maxJ_light | 1;  // odd row stride: fewer bank conflicts
    }
    const bool fused = fz.numJ_pad > 0;
    if (numblkC > 0) {
        rc = launch_step1<1>(ntr, trow0, nw_max, wmax_seen, A, B, w, jlo, jhi, nullptr, C->tile_ptr, wptr, C->tile_columnidx,
                             C->tile_rowidx, pair_ptr, pair_end, pair_a, pair_b, nullptr, fz);
        if (rc) return rc;
    }

    // ---------------- step 2 ----------------
    CK(cudaEventRecord(ev_s2, c.stream));
    // pair-based symbolic for the C tiles the fused step-1 path did not cover: half-warp per tile, or thread per
    // tile when the tiles are hypersparse (<= 2 pairs per C tile and <= 2 entries per A tile on average)
    if (numblkC > 0 && (!fused || wmax_seen > S1_LIGHT_MAX)) {
        const int *wf = fused ? w : nullptr;
        const bool hypersparse = pairs <= 2 * numblkC && A->nnz <= 2ll * A->numtile;
        if (hypersparse)
            k_step2_thread<<<ceil_div(numblkC, S2T_THREADS), S2T_THREADS, 0, c.stream>>>((int)numblkC, pair_ptr, pair_end, pair_a, pair_b,
                                                                                         A->tile_nnz, A->col, B->mask, C->ptr, C->mask,
                                                                                         C->tile_nnz, C->tile_rowidx, wf, trow0, S1_LIGHT_MAX);
        else
            k_step2<<<ceil_div(numblkC * 16, 128), 128, 0, c.stream>>>((int)numblkC, pair_ptr, pair_end, pair_a, pair_b, A->mask, B->mask,
                                                                       C->ptr, C->mask, C->tile_nnz, C->tile_rowidx, wf, trow0, S1_LIGHT_MAX);
        CK_LAUNCH();
    }
    rc = exclusive_scan<long long>(C->tile_nnz, nnz64, numblkC);
    if (rc) return rc;
    rc = exclusive_scan<int>(C->tile_nnz, C->tile_nnz, numblkC);
    if (rc) return rc;
    // pick the accumulator per C tile-row / tile (numeric.cu); its counters come back with nnz(C) in one read-back
    int *d_ns = (int *)(c.d_scalars + 16);
    rc = numeric_classify_device(A, C, trow0, ntr, wptr, &nbufs, d_ns);
    if (rc) return rc;
    long long nnzC = 0;
    rc = publish_words(&c.h_scalars[16], d_ns, 8);
    if (!rc) rc = read_back_i64(nnz64 + numblkC, &nnzC);
    if (rc) return rc;
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
    rc = numeric_device(A, B, C, trow0, ntr, wptr, PairLists{pair_ptr, pair_end, pair_a, pair_b}, nbufs, h_ns, heavy_rows, &nst);
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
        stats->rows_smem = nst.rows_smem;
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
