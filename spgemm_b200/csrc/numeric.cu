// numeric.cu -- TileSpGEMM step 3 (numeric) on B200. Replaces the reference's
// tile_spgemm_step4_cuda_kernel_smem_v3[_halfwarp] / _dns_noatomic_halfwarp (src/tilespgemm-cuda.h:1273-2218)
// and the binning that selects between them (:711-735, launches :2650-2728).
//
// The accumulator is chosen PER C TILE from its nonzero count (BASELINE north_star), and per C tile-row from what
// fits shared memory:
//   * k_step3_rows   -- sparse accumulator in SHARED MEMORY. One CTA per C tile-row. The tile-row's share of A
//                       (row masks, row pointers, values -- contiguous in A's row-major tile order), C's symbolic
//                       result (masks, row pointers) and the tile-row's pair lists are staged into shared memory
//                       with cp.async.bulk (TMA bulk copies, one mbarrier) and plain coalesced loads; one thread
//                       owns one non-empty ROW of one C tile and walks the tile's pairs once: A's row mask gives
//                       the k's, B's row k (mask + Ptr + values, read through L1) gives the products, each added
//                       to the row's compact accumulator in shared memory in the serial SPA's order (ascending A
//                       tile, then k). The finished tile-row leaves with coalesced stores. Every product is
//                       computed once and nothing is searched: ~20 thread-instructions per product where the
//                       lane-per-nonzero gather needs ~115.
//   * k_step3_dense  -- dense accumulator in REGISTERS for well-filled C tiles (nnz >= dense_th), warp per tile,
//                       over a compacted list of those tiles.  k_step3_dmma: its FP64 tensor-core variant, opt-in.
//   * k_step3_gather -- lane per C nonzero, register accumulation: C tile-rows that do not fit shared memory
//                       (R-MAT hub rows) or whose tiles are hypersparse (R-MAT: ~1 nonzero per tile, 15/16 of
//                       the listed tiles empty), where walking (tile, row) slots would be mostly wasted.
// All three add the contributions of a C entry in the same order, so values do not depend on the selection.
#include "common.cuh"
#include "kernels.h"

namespace tsg {

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy helpers (PTX; SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy; dst, src 16-byte aligned, bytes a positive multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Classification. Row kinds: what computes the (non-dense) tiles of a C tile-row.
// ---------------------------------------------------------------------------------------------
enum : uint8_t { ROW_NONE = 0, ROW_STAGED = 1, ROW_GATHER = 2 };
enum { NS_MAXNEED = 0, NS_ROWS_STAGED = 1, NS_ROWS_GATHER = 2, NS_DENSE = 3, NS_GLO = 4, NS_GHI = 5, NS_NONEMPTY = 6 };

__host__ __device__ __forceinline__ size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory bytes k_step3_rows needs for one tile-row (must match the carve-up in the kernel)
__host__ __device__ __forceinline__ size_t s3r_need(int nA, int nnzA, int numJ, int nnzC, int W)
{
    return 2 * al16((size_t)nA * 32) + al16(((size_t)nnzA + 2) * 8) + 2 * al16((size_t)numJ * 32) + al16((size_t)nnzC * 8) +
           al16(((size_t)nA + 1) * 4) + al16(((size_t)numJ + 1) * 4) + 2 * al16((size_t)numJ * 4) + 3 * al16((size_t)W * 4) +
           al16((size_t)nnzC * 2) + al16((size_t)numJ * 2) + al16((size_t)(nnzC < numJ * 15 ? nnzC : numJ * 15) * 2);
}

__global__ void k_ns_init(int *scal)
{
    if (threadIdx.x < 8) scal[threadIdx.x] = threadIdx.x == NS_GLO ? 0x7fffffff : 0;
}

// Pass 1, one thread per C tile: the compacted list of the tiles that take the dense accumulator (order irrelevant:
// tiles are independent), and row_kind[row] = ROW_STAGED for every tile-row that holds at least one non-empty tile
// left to the sparse accumulators (pass 2 decides between staged and gather for exactly those rows).
__global__ void __launch_bounds__(256)
k_s3_classify_tiles(int numblkC, int trow0, const int *__restrict__ c_tile_nnz, const int *__restrict__ c_tile_row, int dense_th,
                    int *__restrict__ dense_list, uint8_t *__restrict__ row_kind, int *__restrict__ scal)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int cnt = t < numblkC ? c_tile_nnz[t + 1] - c_tile_nnz[t] : 0;
    const bool dense = cnt >= dense_th;
    if (cnt > 0 && !dense) row_kind[c_tile_row[t] - trow0] = ROW_STAGED;  // benign race: every writer stores the same value
    const unsigned m = __ballot_sync(FULL_MASK, dense), ne = __ballot_sync(FULL_MASK, cnt > 0);
    if (lane == 0 && ne) atomicAdd(&scal[NS_NONEMPTY], __popc(ne));
    if (!m) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(&scal[NS_DENSE], __popc(m));
    base = __shfl_sync(FULL_MASK, base, 0);
    if (dense) dense_list[base + __popc(m & ((1u << lane) - 1))] = t;
}

__global__ void __launch_bounds__(256)
k_s3_classify_rows(int ntr, int trow0, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_nnz,
                   const int *__restrict__ c_tile_ptr, const int *__restrict__ c_tile_nnz, const int *__restrict__ wptr,
                   const uint8_t *__restrict__ light, int smem_cap, int min_fill, int force_kind, uint8_t *__restrict__ row_kind,
                   int *__restrict__ scal)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint8_t kind = ROW_NONE;
    int need = 0, n0 = 0x7fffffff, n1 = 0;
    if (i < ntr) {
        const int c0 = c_tile_ptr[i], c1 = c_tile_ptr[i + 1], numJ = c1 - c0;
        if (numJ > 0 && row_kind[i] != ROW_NONE) {  // pass 1 found a tile for the sparse accumulators
            n0 = c_tile_nnz[c0]; n1 = c_tile_nnz[c1];
            const int nnzC = n1 - n0;
            if (nnzC > 0) {
                const int a0 = a_tile_ptr[trow0 + i], a1 = a_tile_ptr[trow0 + i + 1];
                const size_t nb = s3r_need(a1 - a0, a_tile_nnz[a1] - a_tile_nnz[a0], numJ, nnzC, wptr[i + 1] - wptr[i]);
                // staged rows need the A-major pair slots, which only the light step-1 path emits
                const bool fits = nb <= (size_t)smem_cap && light[i];
                kind = force_kind ? (uint8_t)force_kind : (fits && (long long)nnzC >= (long long)min_fill * numJ ? ROW_STAGED : ROW_GATHER);
                if (kind == ROW_STAGED && !fits) kind = ROW_GATHER;  // a forced choice still has to fit
                need = kind == ROW_STAGED ? (int)nb : 0;
            }
        }
        row_kind[i] = kind;
    }
    const unsigned ms = __ballot_sync(FULL_MASK, kind == ROW_STAGED), mg = __ballot_sync(FULL_MASK, kind == ROW_GATHER);
    if (kind != ROW_GATHER) { n0 = 0x7fffffff; n1 = 0; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        need = max(need, __shfl_xor_sync(FULL_MASK, need, o));
        n0 = min(n0, __shfl_xor_sync(FULL_MASK, n0, o));
        n1 = max(n1, __shfl_xor_sync(FULL_MASK, n1, o));
    }
    if (lane == 0) {
        if (ms) { atomicAdd(&scal[NS_ROWS_STAGED], __popc(ms)); atomicMax(&scal[NS_MAXNEED], need); }
        if (mg) { atomicAdd(&scal[NS_ROWS_GATHER], __popc(mg)); atomicMin(&scal[NS_GLO], n0); atomicMax(&scal[NS_GHI], n1); }
    }
}

// ---------------------------------------------------------------------------------------------
// k_step3_rows: sparse accumulator in shared memory, CTA per C tile-row (see the file header).
//
// What ncu showed on the way here (profiles/README.md, r2a/r2b): with one thread per (tile, row) and every pair walked
// by every row, a warp averaged 10 active lanes of 32 -- most tiles of a stencil's C hold a handful of entries, and two
// thirds of a tile's pairs involve a one-entry "corner" tile that touches a single row. So the work of a tile-row is
// split three ways, each part laid out so that the lanes of a warp do the same thing:
//   R  well-filled C tiles (>= S3R_ROWS_MIN entries): a HALF-WARP owns the tile, lane = row. It walks the pairs whose
//      A and B tiles both hold more than S3R_SPARSE_MAX entries: A's row mask gives the k's, B's rows k the products,
//      each added to the row's compact accumulator. Tiles are handed out in descending order of their pair count from
//      a shared counter, so the two half-warps of a warp mostly run tiles of the same shape in lockstep.
//   G  the other non-empty C tiles that have such a pair at all: lane = C NONZERO, over a compacted list, register
//      accumulation (the gather formulation, but with A, C's structure and the pair lists in shared memory).
//   S  pairs with a sparse tile on either side are left to k_step3_sparse, a second kernel that adds their contributions
//      for all C tiles of a tile-row in Gustavson order (lane = one B tile of B's tile-row K; the lanes update different
//      C tiles, so nothing conflicts, and the pair lists are not even read). Inside this kernel the same loop kept two
//      of eight warps busy behind a CTA barrier (42 % of the stall samples), and walking the pair lists of a stencil's
//      small C tiles -- fed by sparse pairs only -- cost more than all the products of the well-filled ones.
// The order in which a C entry's contributions are added is fixed (R/G in pair order, then S in ascending A tile), so
// results are reproducible run to run; for entries fed by both it is not the serial SPA's order (values agree to
// rounding, and exactly for integer-valued inputs).
// ---------------------------------------------------------------------------------------------
constexpr int S3R_SPARSE_MAX = 2;   // tiles with at most this many entries are "sparse" (phase S / short path)
constexpr int S3R_ROWS_MIN = 16;    // C tiles with at least this many entries run lane-per-row (phase R)
constexpr unsigned PF_SPARSE = 0x80000000u;
constexpr int S3R_BUCKETS = 64;

struct S3Rows {
    int trow0, dense_th;
    const int *a_tile_ptr, *a_tile_col, *a_tile_nnz;
    const uint16_t *a_ptr, *a_mask, *a_col;
    const double *a_val;
    const int *b_tile_ptr, *b_tile_col, *b_rm2csc, *b_tile_nnz;
    const uint16_t *b_ptr, *b_mask, *b_col;
    const double *b_val;
    const int *c_tile_ptr, *c_tile_col, *c_tile_nnz;
    const uint16_t *c_ptr, *c_mask;
    uint16_t *c_col;
    double *c_val;
    const int *wptr, *pair_ptr, *pair_end, *pair_a, *pair_b;
    const uint16_t *pair_slot;
    const uint8_t *row_kind;
};

// row of entry j of a tile whose 16 exclusive row offsets are q0,q1: the last row whose offset is <= j
__device__ __forceinline__ int s3_row_of(const uint4 q0, const uint4 q1, int j)
{
    const unsigned key = (unsigned)j * 0x10001u;
    const int n = __popc(__vcmpleu2(q0.x, key)) + __popc(__vcmpleu2(q0.y, key)) + __popc(__vcmpleu2(q0.z, key)) + __popc(__vcmpleu2(q0.w, key)) +
                  __popc(__vcmpleu2(q1.x, key)) + __popc(__vcmpleu2(q1.y, key)) + __popc(__vcmpleu2(q1.z, key)) + __popc(__vcmpleu2(q1.w, key));
    return (n >> 4) - 1;  // each u16 that compares <= contributes 16 set bits
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 4 * 256 / THREADS)
k_step3_rows(const __grid_constant__ S3Rows P)
{
    extern __shared__ __align__(128) unsigned char s3r_smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_hist[S3R_BUCKETS];
    __shared__ int s_nord, s_ngather, s_next;
    const int i = blockIdx.x, tid = threadIdx.x;
    if (P.row_kind[i] != ROW_STAGED) return;
    const int I = P.trow0 + i;
    const int a0 = P.a_tile_ptr[I], nA = P.a_tile_ptr[I + 1] - a0;
    const int av0 = P.a_tile_nnz[a0], av1 = P.a_tile_nnz[a0 + nA];
    const int c0 = P.c_tile_ptr[i], numJ = P.c_tile_ptr[i + 1] - c0;
    const int n0 = P.c_tile_nnz[c0], nnzC = P.c_tile_nnz[c0 + numJ] - n0;
    const int w0 = P.wptr[i], W = P.wptr[i + 1] - w0;
    const int av0a = av0 & ~1;                       // bulk copies need 16-byte aligned sources: start at an even element
    const int nav = (av1 - av0a + 1) & ~1;           // and copy an even number of doubles (the slab is padded)

    size_t off = 0;
    auto carve = [&](size_t bytes) { unsigned char *p = s3r_smem + off; off += al16(bytes); return p; };
    uint16_t *s_am = (uint16_t *)carve((size_t)nA * 32);
    uint16_t *s_ap = (uint16_t *)carve((size_t)nA * 32);
    double *s_aval = (double *)carve(((size_t)(av1 - av0) + 2) * 8);
    uint16_t *s_cm = (uint16_t *)carve((size_t)numJ * 32);
    uint16_t *s_cp = (uint16_t *)carve((size_t)numJ * 32);
    double *s_out = (double *)carve((size_t)nnzC * 8);
    int *s_annz = (int *)carve(((size_t)nA + 1) * 4);
    int *s_cnnz = (int *)carve(((size_t)numJ + 1) * 4);
    int *s_pp = (int *)carve((size_t)numJ * 4);
    int *s_pe = (int *)carve((size_t)numJ * 4);
    int *s_pa = (int *)carve((size_t)W * 4);
    int *s_pb = (int *)carve((size_t)W * 4);
    int *s_pbn = (int *)carve((size_t)W * 4);
    uint16_t *s_ocol = (uint16_t *)carve((size_t)nnzC * 2);
    uint16_t *s_order = (uint16_t *)carve((size_t)numJ * 2);
    uint16_t *s_gitem = (uint16_t *)carve((size_t)min(nnzC, numJ * (S3R_ROWS_MIN - 1)) * 2);

    if (tid == 0) { mbar_init(&s_bar, 1); s_nord = 0; s_ngather = 0; s_next = 0; }
    if (tid < S3R_BUCKETS) s_hist[tid] = 0;
    __syncthreads();
    if (tid == 0) {  // one elected thread arms the barrier and issues the five bulk copies
        mbar_expect_tx(&s_bar, (uint32_t)(nA * 64 + nav * 8 + numJ * 64));
        bulk_g2s(s_am, P.a_mask + (size_t)a0 * TS, (uint32_t)nA * 32, &s_bar);
        bulk_g2s(s_ap, P.a_ptr + (size_t)a0 * TS, (uint32_t)nA * 32, &s_bar);
        if (nav > 0) bulk_g2s(s_aval, P.a_val + av0a, (uint32_t)nav * 8, &s_bar);
        bulk_g2s(s_cm, P.c_mask + (size_t)c0 * TS, (uint32_t)numJ * 32, &s_bar);
        bulk_g2s(s_cp, P.c_ptr + (size_t)c0 * TS, (uint32_t)numJ * 32, &s_bar);
    }
    // everything whose source is only 4-byte aligned: plain coalesced loads, overlapping the bulk copies
    for (int k = tid; k <= nA; k += THREADS) s_annz[k] = P.a_tile_nnz[a0 + k] - av0a;
    int dense_here = 0;
    for (int k = tid; k <= numJ; k += THREADS) {
        const int v = P.c_tile_nnz[c0 + k] - n0;
        s_cnnz[k] = v;
        if (k < numJ) {
            const int pp = P.pair_ptr[c0 + k] - w0, pe = P.pair_end[c0 + k] - w0;
            s_pp[k] = pp;
            s_pe[k] = pe;
            const int cnt = P.c_tile_nnz[c0 + k + 1] - n0 - v;
            if (cnt >= P.dense_th) dense_here = 1;
            else if (cnt >= S3R_ROWS_MIN) atomicAdd(&s_hist[S3R_BUCKETS - 1 - min(pe - pp, S3R_BUCKETS - 1)], 1);  // most pairs first
        }
    }
    for (int k = tid; k < W; k += THREADS) {  // pairs with a sparse tile on either side belong to k_step3_sparse
        const int ta = P.pair_a[w0 + k], b = P.pair_b[w0 + k];
        const int bn0 = P.b_tile_nnz[b];
        unsigned f = (unsigned)(ta - a0);
        if (P.a_tile_nnz[ta + 1] - P.a_tile_nnz[ta] <= S3R_SPARSE_MAX || P.b_tile_nnz[b + 1] - bn0 <= S3R_SPARSE_MAX) f |= PF_SPARSE;
        s_pa[k] = (int)f;
        s_pb[k] = b;
        s_pbn[k] = bn0;
    }
    for (int k = tid; k < nnzC; k += THREADS) s_out[k] = 0.0;
    const int has_dense = __syncthreads_or(dense_here);
    if (tid < 32) {  // exclusive scan of the 64 bucket counts by one warp
        const int v0 = s_hist[2 * tid], v1 = s_hist[2 * tid + 1];
        int incl = v0 + v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, o);
            if (tid >= o) incl += t;
        }
        s_hist[2 * tid] = incl - v0 - v1;
        s_hist[2 * tid + 1] = incl - v1;
        if (tid == 31) s_nord = incl;
    }
    mbar_wait(&s_bar, 0);
    __syncthreads();
    for (int k = tid; k < numJ; k += THREADS) {
        const int tb = s_cnnz[k], cnt = s_cnnz[k + 1] - tb;
        if (cnt == 0 || cnt >= P.dense_th) continue;
        if (cnt >= S3R_ROWS_MIN) {  // counting sort of the phase-R tiles by pair count, descending
            s_order[atomicAdd(&s_hist[S3R_BUCKETS - 1 - min(s_pe[k] - s_pp[k], S3R_BUCKETS - 1)], 1)] = (uint16_t)k;
            continue;
        }
        // a sparsely filled tile: its columns are written here; its nonzeros join phase G's list if any of its pairs
        // is left to this kernel (a stencil's small C tiles are fed by sparse pairs only)
        int o = tb;
        for (int r = 0; r < TS && o < tb + cnt; r++) {
            unsigned m = s_cm[k * TS + r];
            while (m) { const int c = __clz(m) - 16; m ^= 0x8000u >> c; s_ocol[o++] = (uint16_t)c; }
        }
        bool work = false;
        for (int p = s_pp[k]; p < s_pe[k] && !work; p++) work = !((unsigned)s_pa[p] & PF_SPARSE);
        if (work) {
            const int base = atomicAdd(&s_ngather, cnt);
            for (int j = 0; j < cnt; j++) s_gitem[base + j] = (uint16_t)(tb + j);
        }
    }
    __syncthreads();

    const int lane = tid & 31, l16 = tid & 15;
    const unsigned hm = 0xFFFFu << (lane & 16);
    const double *__restrict__ bvals = P.b_val;

    // ---------------- phase R: half-warp per well-filled tile, lane = row ----------------
    const int nord = s_nord;
    while (true) {
        int q = 0;
        if (l16 == 0) q = atomicAdd(&s_next, 1);
        q = __shfl_sync(hm, q, lane & 16);
        if (q >= nord) break;
        const int s = s_order[q];
        const int r = l16;
        const unsigned cm = s_cm[s * TS + r];
        if (!cm) continue;
        const int rowbase = s_cnnz[s] + s_cp[s * TS + r];
        {
            unsigned m = cm;
            int o = rowbase;
            do { const int c = __clz(m) - 16; m ^= 0x8000u >> c; s_ocol[o++] = (uint16_t)c; } while (m);
        }
        const unsigned cmr = __brev(cm) >> 16;  // bit c = column c
        const int pe = s_pe[s];
        for (int p = s_pp[s]; p < pe; p++) {
            const unsigned fa = (unsigned)s_pa[p];
            if (fa & PF_SPARSE) continue;  // k_step3_sparse
            const int a = (int)fa;
            unsigned am = s_am[a * TS + r];
            if (!am) continue;  // the pair does not touch this row
            const int bbase = s_pbn[p];
            int ia = s_annz[a] + s_ap[a * TS + r];
            const int b = s_pb[p];
            const uint16_t *bmk = P.b_mask + (size_t)b * TS, *bpt = P.b_ptr + (size_t)b * TS;
            do {  // the bits of A's row mask are the k's of the row, ascending
                const int k = __clz(am) - 16;
                am ^= 0x8000u >> k;
                const double av = s_aval[ia++];
                unsigned bm = __brev((unsigned)bmk[k]) >> 16;
                int ib = bbase + bpt[k];
                while (bm) {  // B's row k: every entry is a product into C's row r
                    const unsigned low = bm & (0u - bm);
                    const int o = rowbase + __popc(cmr & (low - 1));  // rank of the column in C's row
                    bm ^= low;
                    s_out[o] = fma(av, bvals[ib++], s_out[o]);
                }
            } while (am);
        }
    }

    // ---------------- phase G: lane = one nonzero of the sparsely filled tiles ----------------
    const int ngather = s_ngather;
    for (int g = tid; g < ngather; g += THREADS) {
        const int o = s_gitem[g];
        int lo = 0, hi = numJ - 1;  // the tile holding nonzero o: largest s with s_cnnz[s] <= o
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_cnnz[mid] <= o) lo = mid; else hi = mid - 1;
        }
        const int s = lo;
        const uint4 *pq = reinterpret_cast<const uint4 *>(s_cp + s * TS);
        const int r = s3_row_of(pq[0], pq[1], o - s_cnnz[s]);
        unsigned cmr = __brev((unsigned)s_cm[s * TS + r]) >> 16;
        for (int n = o - s_cnnz[s] - (int)s_cp[s * TS + r]; n > 0; n--) cmr &= cmr - 1;  // drop the n smaller columns
        const int c = __ffs(cmr) - 1;
        const unsigned cbit = 0x8000u >> c;
        double acc = 0.0;
        const int pe = s_pe[s];
        for (int p = s_pp[s]; p < pe; p++) {
            const unsigned fa = (unsigned)s_pa[p];
            if (fa & PF_SPARSE) continue;  // k_step3_sparse
            const int a = (int)fa;
            unsigned am = s_am[a * TS + r];
            if (!am) continue;
            const int b = s_pb[p], bbase = s_pbn[p];
            int ia = s_annz[a] + s_ap[a * TS + r];
            do {
                const int k = __clz(am) - 16;
                am ^= 0x8000u >> k;
                const unsigned bm = P.b_mask[(size_t)b * TS + k];
                if (bm & cbit) acc = fma(s_aval[ia], bvals[bbase + (int)P.b_ptr[(size_t)b * TS + k] + __popc(bm >> (16 - c))], acc);
                ia++;
            } while (am);
        }
        s_out[o] = acc;
    }
    __syncthreads();

    if (!has_dense) {
        for (int k = tid; k < nnzC; k += THREADS) { P.c_val[n0 + k] = s_out[k]; P.c_col[n0 + k] = s_ocol[k]; }
    } else {  // leave the ranges of the dense tiles alone
        for (int k = tid; k < nnzC; k += THREADS) {
            int lo = 0, hi = numJ - 1;  // largest s with s_cnnz[s] <= k
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_cnnz[mid] <= k) lo = mid; else hi = mid - 1;
            }
            if (s_cnnz[lo + 1] - s_cnnz[lo] >= P.dense_th) continue;
            P.c_val[n0 + k] = s_out[k];
            P.c_col[n0 + k] = s_ocol[k];
        }
    }
}

// Phase S of the staged tile-rows (see k_step3_rows): every pair with a sparse tile (<= S3R_SPARSE_MAX entries) on either
// side, in Gustavson order. One WARP per C tile-row walks the A tiles (I,K) of the tile-row in ascending order; lane = one
// B tile (K,J) of B's tile-row K; the slot of C tile (I,J) in the tile-row comes from step 1 (pair_slot, A-major order):
//   * A tile sparse: for each of its entries (r, k, v), v * B(k, :) is added into C's row r;
//   * A tile not sparse, B tile sparse: for each entry (k, c, v) of the B tile, A(r, k) * v is added into C(r, c) for the
//     rows r of A that hold column k.
// The lanes of a warp update different C tiles, so nothing conflicts; C's values (written by k_step3_rows) are updated
// in place, in a fixed order. Tiles of the dense accumulator and gathered tile-rows are skipped: those kernels walk all
// pairs themselves.
__global__ void __launch_bounds__(256)
k_step3_sparse(int ntr, const __grid_constant__ S3Rows P)
{
    const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= ntr || P.row_kind[i] != ROW_STAGED) return;
    const int I = P.trow0 + i;
    const int c0 = P.c_tile_ptr[i];
    const int a1 = P.a_tile_ptr[I + 1];
    int widx = P.wptr[i];  // A-major index of the first pair of the current A tile (the order step 1 enumerated them in)
    for (int a = P.a_tile_ptr[I]; a < a1; a++) {
        const int e0 = P.a_tile_nnz[a], e1 = P.a_tile_nnz[a + 1];
        const bool a_sparse = e1 - e0 <= S3R_SPARSE_MAX;
        const int K = P.a_tile_col[a];
        const int t0 = P.b_tile_ptr[K], t1 = P.b_tile_ptr[K + 1];
        const int wa = widx;
        widx += t1 - t0;
        for (int tb0 = t0; tb0 < t1; tb0 += 32) {
            const int tb = tb0 + lane;
            // this lane's pair: B tile and C tile (skipped if the C tile is empty or belongs to the dense accumulator)
            int b = 0, bn0 = 0, bn = 0, t = 0, cb = 0;
            bool live = tb < t1;
            if (live) {
                b = P.b_rm2csc[tb];
                bn0 = P.b_tile_nnz[b];
                bn = P.b_tile_nnz[b + 1] - bn0;
                live = a_sparse || bn <= S3R_SPARSE_MAX;
            }
            if (live) {
                t = c0 + P.pair_slot[wa + (tb - t0)];  // C tile (I, J) of this pair, as step 1 ranked it
                cb = P.c_tile_nnz[t];
                const int cnt = P.c_tile_nnz[t + 1] - cb;
                live = cnt > 0 && cnt < P.dense_th;
            }
            if (a_sparse) {
                for (int e = e0; e < e1; e++) {  // warp-uniform: A's entries (r, k, v)
                    const unsigned col = P.a_col[e];  // A stores row*16+col
                    const int r = col >> 4, k = col & 15;
                    const double av = P.a_val[e];
                    if (live) {
                        const unsigned cm = P.c_mask[(size_t)t * TS + r];
                        unsigned bm = cm ? __brev((unsigned)P.b_mask[(size_t)b * TS + k]) >> 16 : 0u;
                        if (bm) {
                            int ib = bn0 + P.b_ptr[(size_t)b * TS + k];
                            const int rowbase = cb + P.c_ptr[(size_t)t * TS + r];
                            const unsigned cmr = __brev(cm) >> 16;
                            do {
                                const unsigned low = bm & (0u - bm);
                                const int o = rowbase + __popc(cmr & (low - 1));
                                bm ^= low;
                                P.c_val[o] = fma(av, P.b_val[ib++], P.c_val[o]);
                            } while (bm);
                        }
                    }
                }
            } else if (live) {  // B's one or two entries (k, c, v) against the rows of A that hold column k
                const uint4 *pq = reinterpret_cast<const uint4 *>(P.b_ptr + (size_t)b * TS);
                const uint4 q0 = pq[0], q1 = pq[1];
                const int abase = e0;
                for (int j = 0; j < bn; j++) {
                    const int k = s3_row_of(q0, q1, j), c = P.b_col[bn0 + j];
                    const double bv = P.b_val[bn0 + j];
                    const unsigned kbit = 0x8000u >> k;
                    for (int r = 0; r < TS; r++) {
                        const unsigned am = P.a_mask[(size_t)a * TS + r];
                        if (!(am & kbit)) continue;
                        const unsigned cmr = __brev((unsigned)P.c_mask[(size_t)t * TS + r]) >> 16;
                        const int o = cb + P.c_ptr[(size_t)t * TS + r] + __popc(cmr & ((1u << c) - 1));
                        const double av = P.a_val[abase + P.a_ptr[(size_t)a * TS + r] + __popc(am >> (16 - k))];
                        P.c_val[o] = fma(av, bv, P.c_val[o]);
                    }
                }
            }
            __syncwarp();  // the next pass may update the same C entries from other lanes
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Gather formulation, one LANE per C nonzero (g = position in C's Val/Col).
// blk2tile[g/32] gives the tile holding nonzero 32*(g/32); the lane finds its tile, row r and
// column c from tile_nnz / Ptr / mask, then for every pair (A tile a, B tile b) of the tile reads
// A's row mask r (zero: the pair does not touch this row and nothing else of the pair is loaded);
// its bits are the k's of A's row r in ascending order. For an entry (r,k) with value av, B has
// (k,c) iff bit (15-c) of B's row mask k is set, at Ptr_b[k] + popc(mask bits of columns < c).
// The sum stays in a register: no accumulator memory, no atomics, coalesced stores, and the
// summation order (ascending A tile, then ascending k) is the serial SPA's.
// ---------------------------------------------------------------------------------------------
__global__ void k_blk2tile(int numblkC, const int *__restrict__ c_tile_nnz, int *__restrict__ blk2tile)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numblkC) return;
    const int s = c_tile_nnz[t], e = c_tile_nnz[t + 1];
    for (int blk = (s + 31) >> 5; (blk << 5) < e; blk++) blk2tile[blk] = t;
}

struct S3Gather {
    int numblkC, nnzC, trow0, dense_th;
    const int *blk2tile, *pair_ptr, *pair_end, *pair_a, *pair_b;
    const int *a_tile_nnz;
    const uint16_t *a_ptr, *a_mask;
    const double *a_val;
    const int *b_tile_nnz;
    const uint16_t *b_ptr, *b_mask;
    const double *b_val;
    const int *c_tile_nnz, *c_tile_row;
    const uint16_t *c_ptr, *c_mask;
    uint16_t *c_col;
    double *c_val;
    const uint8_t *row_kind;  // nullptr: every tile is gathered
};

template <int UNROLL>
__device__ __forceinline__ void s3_gather_one(int g, const S3Gather &P)
{
    const int blk = g >> 5, nblk = (P.nnzC + 31) >> 5;
    int lo = P.blk2tile[blk], hi = blk + 1 < nblk ? P.blk2tile[blk + 1] : P.numblkC - 1;
    while (lo < hi) {  // largest tile t in [lo,hi] with tile_nnz[t] <= g (it is non-empty and holds g)
        int mid = (lo + hi + 1) >> 1;
        if (P.c_tile_nnz[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const int t = lo;
    const int tbase = P.c_tile_nnz[t];
    if (P.row_kind) {  // tiles computed by the other numeric kernels
        if (P.row_kind[P.c_tile_row[t] - P.trow0] != ROW_GATHER) return;
        if (P.c_tile_nnz[t + 1] - tbase >= P.dense_th) return;
    }
    const int off = g - tbase;
    // row: largest r with Ptr[r] <= off; the 16 u16 offsets are one aligned 32-byte line
    const uint4 *pp = reinterpret_cast<const uint4 *>(P.c_ptr + (size_t)t * TS);
    const uint4 q0 = pp[0], q1 = pp[1];
    const unsigned key = (unsigned)off * 0x10001u;
    int r = -1;
    r += __popc(__vcmpleu2(q0.x, key)) + __popc(__vcmpleu2(q0.y, key)) + __popc(__vcmpleu2(q0.z, key)) + __popc(__vcmpleu2(q0.w, key)) +
         __popc(__vcmpleu2(q1.x, key)) + __popc(__vcmpleu2(q1.y, key)) + __popc(__vcmpleu2(q1.z, key)) + __popc(__vcmpleu2(q1.w, key));
    r = ((r + 1) >> 4) - 1;  // each u16 that compares <= contributes 16 set bits
    unsigned cm = __brev(P.c_mask[(size_t)t * TS + r]) >> 16;  // bit c = column c present
    for (int n = off - (int)P.c_ptr[(size_t)t * TS + r]; n > 0; n--) cm &= cm - 1;  // drop the n smaller columns
    const int c = __ffs(cm) - 1;
    const unsigned cbit = 0x8000u >> c;

    double acc = 0.0;
    // one pair: `am` = A's row mask r (non-zero); its bits are the k's in ascending order, so A's Col array is never read
    auto pair_contrib = [&](int a, int b, unsigned am) {
        int ia = P.a_tile_nnz[a] + P.a_ptr[(size_t)a * TS + r];
        const int bbase = P.b_tile_nnz[b];
        do {
            const int k = __clz(am) - 16;
            am ^= 0x8000u >> k;
            const unsigned bm = P.b_mask[(size_t)b * TS + k];
            if (bm & cbit) {
                const int pos = (int)P.b_ptr[(size_t)b * TS + k] + __popc(bm >> (16 - c));
                acc = fma(P.a_val[ia], P.b_val[bbase + pos], acc);
            }
            ia++;
        } while (am);
    };
    const int p1 = P.pair_end[t];
    int p = P.pair_ptr[t];
    if (UNROLL > 1) {  // the row masks of UNROLL pairs in flight; contributions are still added in pair order
        for (; p + UNROLL <= p1; p += UNROLL) {
            int a[UNROLL];
            unsigned am[UNROLL];
#pragma unroll
            for (int j = 0; j < UNROLL; j++) a[j] = P.pair_a[p + j];
#pragma unroll
            for (int j = 0; j < UNROLL; j++) am[j] = P.a_mask[(size_t)a[j] * TS + r];
#pragma unroll
            for (int j = 0; j < UNROLL; j++)
                if (am[j]) pair_contrib(a[j], P.pair_b[p + j], am[j]);
        }
    }
    for (; p < p1; p++) {
        const int a = P.pair_a[p];
        const unsigned am = P.a_mask[(size_t)a * TS + r];  // zero: the pair does not touch row r, nothing else of it is loaded
        if (am) pair_contrib(a, P.pair_b[p], am);
    }
    P.c_val[g] = acc;
    P.c_col[g] = (uint16_t)c;
}

// CHUNKED = false: one nonzero per thread, blocks balanced by the hardware scheduler.
// CHUNKED = true: a CTA walks `chunk` consecutive nonzeros (a few C tile-rows), so that the A tiles they share stay in
// its SM's L1; used when the work per nonzero is even (no heavy tile-rows) and the grid stays large.
template <bool CHUNKED, int UNROLL>
__global__ void __launch_bounds__(256)
k_step3_gather(int chunk, int g0, int g1, const __grid_constant__ S3Gather P)
{
    if (!CHUNKED) {
        const long long g = (long long)g0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (g < g1) s3_gather_one<UNROLL>((int)g, P);
        return;
    }
    const long long cend = min((long long)g1, (long long)g0 + ((long long)blockIdx.x + 1) * chunk);
    for (long long g = (long long)g0 + (long long)blockIdx.x * chunk + threadIdx.x; g < cend; g += blockDim.x)
        s3_gather_one<UNROLL>((int)g, P);
}

// ---------------------------------------------------------------------------------------------
// Dense accumulator for WELL-FILLED C tiles (block-FEM: C tiles hold 160 nonzeros on average). One warp per listed
// C tile; lane = (row r = lane/2, column half h = lane%2) owns the 8 entries C[r][8h..8h+7] in REGISTERS. Per pair
// the B tile is expanded to a dense 16x16 tile in shared memory (rows padded to 18 doubles: 16-byte aligned, 2-way
// conflicts at worst); the two lanes of row r walk A's row r and, per entry (r,k,av), do 8 FMAs with B's dense row k
// (4 x LDS.128). No masks, popcounts or atomics in the inner loop; the row is compacted through C's mask at the end.
// Summation order per C entry: ascending A tile, then ascending k -- the serial SPA's order.
// ---------------------------------------------------------------------------------------------
constexpr int S3D_WARPS = 4;
constexpr int S3D_LD = 18;

struct S3Dense {
    int ntiles;
    const int *list;  // C tiles to compute (nullptr: tiles 0..ntiles-1, empty ones skipped)
    const int *pair_ptr, *pair_end, *pair_a, *pair_b;
    const int *a_tile_nnz;
    const uint16_t *a_ptr, *a_col;
    const double *a_val;
    const int *b_tile_nnz;
    const uint16_t *b_ptr, *b_col;
    const double *b_val;
    const int *c_tile_nnz;
    const uint16_t *c_ptr, *c_mask;
    uint16_t *c_col;
    double *c_val;
};

__global__ void __launch_bounds__(S3D_WARPS * 32)
k_step3_dense(const __grid_constant__ S3Dense P)
{
    __shared__ __align__(16) double Bd_s[S3D_WARPS][TS * S3D_LD];
    const int wi = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, r = lane >> 1, h = lane & 1;
    if (wi >= P.ntiles) return;
    const int t = P.list ? P.list[wi] : wi;
    const int cbase = P.c_tile_nnz[t];
    if (P.c_tile_nnz[t + 1] == cbase) return;  // empty tile (warp-uniform)
    double *Bd = Bd_s[w];
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.0;
    // Software pipeline over the pairs (round 2, after ncu showed the 32 x 32 sibling of this kernel, gentile.cu, waiting on
    // dependent global loads 51 % of the time): the tile ids of pair p+1 are requested at the top of pair p, what depends
    // on them (offsets, row ranges) after B's expansion, the next A entry under the current entry's FMAs; B's entries are
    // expanded lane per ENTRY (coalesced, two per lane in flight; the row of entry x = last k with Ptr[k] <= x, found by
    // four shuffles over the offsets held by lanes 0..15) instead of lane per row with one dependent load chain per entry.
    const int p0 = P.pair_ptr[t], p1 = P.pair_end[t];
    if (p0 < p1) {
        int a = P.pair_a[p0], b = P.pair_b[p0];
        int abase = P.a_tile_nnz[a], bbase = P.b_tile_nnz[b], nB = P.b_tile_nnz[b + 1] - bbase;
        int ia = P.a_ptr[(size_t)a * TS + r];
        int ia1 = r < TS - 1 ? (int)P.a_ptr[(size_t)a * TS + r + 1] : P.a_tile_nnz[a + 1] - abase;
        int pb = lane < TS ? (int)P.b_ptr[(size_t)b * TS + lane] : 0x7fffffff;
        for (int p = p0; p < p1; p++) {
            const bool more = p + 1 < p1;
            int an = 0, bn = 0;
            if (more) { an = P.pair_a[p + 1]; bn = P.pair_b[p + 1]; }
            int kc = 0;
            double av = 0.0;
            if (ia < ia1) { kc = P.a_col[abase + ia] & 15; av = P.a_val[abase + ia]; }  // A stores row*16+col
            double2 *z = reinterpret_cast<double2 *>(Bd + r * S3D_LD + 8 * h);
            z[0] = z[1] = z[2] = z[3] = make_double2(0.0, 0.0);
            __syncwarp();
            for (int x0 = 0; x0 < nB; x0 += 64) {  // warp-uniform trip count
                const int xa = x0 + lane, xb = x0 + 32 + lane;
                int ca = 0, cb = 0;
                double va = 0.0, vb = 0.0;
                if (xa < nB) { ca = P.b_col[bbase + xa]; va = P.b_val[bbase + xa]; }
                if (xb < nB) { cb = P.b_col[bbase + xb]; vb = P.b_val[bbase + xb]; }
                int ka = 0, kb = 0;
#pragma unroll
                for (int st = 8; st; st >>= 1) {
                    const int va_ = __shfl_sync(FULL_MASK, pb, ka + st), vb_ = __shfl_sync(FULL_MASK, pb, kb + st);
                    if (va_ <= xa) ka += st;
                    if (vb_ <= xb) kb += st;
                }
                if (xa < nB) Bd[ka * S3D_LD + ca] = va;
                if (xb < nB) Bd[kb * S3D_LD + cb] = vb;
            }
            __syncwarp();
            int abase_n = 0, bbase_n = 0, nB_n = 0, ia_n = 0, ia1_n = 0, pb_n = 0x7fffffff;
            if (more) {
                abase_n = P.a_tile_nnz[an]; bbase_n = P.b_tile_nnz[bn]; nB_n = P.b_tile_nnz[bn + 1] - bbase_n;
                ia_n = P.a_ptr[(size_t)an * TS + r];
                ia1_n = r < TS - 1 ? (int)P.a_ptr[(size_t)an * TS + r + 1] : P.a_tile_nnz[an + 1] - abase_n;
                if (lane < TS) pb_n = (int)P.b_ptr[(size_t)bn * TS + lane];
            }
            for (; ia < ia1; ia++) {
                int kn = 0;
                double avn = 0.0;
                if (ia + 1 < ia1) { kn = P.a_col[abase + ia + 1] & 15; avn = P.a_val[abase + ia + 1]; }
                const double2 *br = reinterpret_cast<const double2 *>(Bd + kc * S3D_LD + 8 * h);
                const double2 b0 = br[0], b1 = br[1], b2 = br[2], b3 = br[3];
                acc[0] = fma(av, b0.x, acc[0]); acc[1] = fma(av, b0.y, acc[1]);
                acc[2] = fma(av, b1.x, acc[2]); acc[3] = fma(av, b1.y, acc[3]);
                acc[4] = fma(av, b2.x, acc[4]); acc[5] = fma(av, b2.y, acc[5]);
                acc[6] = fma(av, b3.x, acc[6]); acc[7] = fma(av, b3.y, acc[7]);
                kc = kn; av = avn;
            }
            __syncwarp();  // Bd is rewritten for the next pair
            abase = abase_n; bbase = bbase_n; nB = nB_n; ia = ia_n; ia1 = ia1_n; pb = pb_n;
        }
    }
    const unsigned cm = P.c_mask[(size_t)t * TS + r];
    const int rowbase = cbase + P.c_ptr[(size_t)t * TS + r];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c = 8 * h + j;
        if (cm & (0x8000u >> c)) {
            const int pos = rowbase + __popc(cm >> (16 - c));
            P.c_val[pos] = acc[j];
            P.c_col[pos] = (uint16_t)c;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core experiment (BASELINE north_star: "DMMA only for near-dense tile pairs, and only if ncu shows
// they beat the CUDA-core path"). Same work split as k_step3_dense, but BOTH tiles of a pair are expanded to dense
// 16x16 tiles in shared memory and the 16x16x16 product is issued as 2x2 output blocks x 4 k-steps = 16
// mma.sync.m8n8k4.f64 (SASS: DMMA); the 8 accumulators per lane are the C fragments. (tcgen05 has no FP64 kind, so
// this is the only tensor-core path FP64 has.) Selected with TSG_STEP3=dmma; outcome in profiles/README.md.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(S3D_WARPS * 32)
k_step3_dmma(const __grid_constant__ S3Dense P)
{
    __shared__ __align__(16) double AB_s[S3D_WARPS][2][TS * S3D_LD];
    const int wi = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
    if (wi >= P.ntiles) return;
    const int t = P.list ? P.list[wi] : wi;
    const int cbase = P.c_tile_nnz[t];
    if (P.c_tile_nnz[t + 1] == cbase) return;  // empty tile (warp-uniform)
    double *Ad = AB_s[w][0], *Bd = AB_s[w][1];
    double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c10[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
    const int p1 = P.pair_end[t];
    for (int p = P.pair_ptr[t]; p < p1; p++) {
        const int a = P.pair_a[p], b = P.pair_b[p];
        const int abase = P.a_tile_nnz[a], aend = P.a_tile_nnz[a + 1], bbase = P.b_tile_nnz[b];
        double2 *z = reinterpret_cast<double2 *>(AB_s[w][0]);
        for (int k = lane; k < 2 * TS * S3D_LD / 2; k += 32) z[k] = make_double2(0.0, 0.0);
        __syncwarp();
        for (int e = abase + lane; e < aend; e += 32) {
            const unsigned col = P.a_col[e];  // row*16+col
            Ad[(col >> 4) * S3D_LD + (col & 15)] = P.a_val[e];
        }
        if (lane < TS) {
            int ib = P.b_ptr[(size_t)b * TS + lane];
            const int ib1 = lane < TS - 1 ? (int)P.b_ptr[(size_t)b * TS + lane + 1] : P.b_tile_nnz[b + 1] - bbase;
            for (; ib < ib1; ib++) Bd[lane * S3D_LD + P.b_col[bbase + ib]] = P.b_val[bbase + ib];
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const double a0 = Ad[g * S3D_LD + 4 * ks + tg], a1 = Ad[(8 + g) * S3D_LD + 4 * ks + tg];
            const double b0 = Bd[(4 * ks + tg) * S3D_LD + g], b1 = Bd[(4 * ks + tg) * S3D_LD + 8 + g];
            dmma_m8n8k4(c00[0], c00[1], a0, b0);
            dmma_m8n8k4(c01[0], c01[1], a0, b1);
            dmma_m8n8k4(c10[0], c10[1], a1, b0);
            dmma_m8n8k4(c11[0], c11[1], a1, b1);
        }
        __syncwarp();
    }
    // fragment (i,j) holds C[8i+g][8j+2tg+{0,1}]
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int R = 8 * i + g;
        const unsigned cm = P.c_mask[(size_t)t * TS + R];
        const int rowbase = cbase + P.c_ptr[(size_t)t * TS + R];
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int c = 8 * j + 2 * tg + q;
                if (cm & (0x8000u >> c)) {
                    const int pos = rowbase + __popc(cm >> (16 - c));
                    P.c_val[pos] = i == 0 ? (j == 0 ? c00[q] : c01[q]) : (j == 0 ? c10[q] : c11[q]);
                    P.c_col[pos] = (uint16_t)c;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

// TSG_STEP3 = auto (default) | rows | gather | dense | dmma : force one accumulator for every tile (A/B measurements)
static const char *s3_mode()
{
    const char *m = getenv("TSG_STEP3");  // read on every call: the tests switch it between calls
    return m && *m ? m : "auto";
}

size_t numeric_scratch_bytes(int ntr, long long numblkC)
{
    return arena_need((size_t)ntr + 1, 1) + arena_need((size_t)(numblkC > 0 ? numblkC : 1), 4);
}

// Enqueue the classification of the slab's tile-rows and tiles (after C's tile_nnz has been scanned). Results: row_kind,
// dense_list and the NS_* counters in d_ns (device ints), which the caller reads back together with nnz(C).
int numeric_classify_device(const tsg_dtile *A, const tsg_dtile *C, int trow0, int ntr, const int *wptr, const uint8_t *light, NumericBufs *nb,
                            int *d_ns)
{
    Ctx &c = ctx();
    const char *mode = s3_mode();
    const long long numblkC = C->numtile;
    k_ns_init<<<1, 32, 0, c.stream>>>(d_ns);
    CK_LAUNCH();
    int force_kind = 0, dense_th = env_int("TSG_DENSE_TH", 96);
    if (!strcmp(mode, "gather")) { force_kind = ROW_GATHER; dense_th = 1 << 20; }
    else if (!strcmp(mode, "rows")) { force_kind = ROW_STAGED; dense_th = 1 << 20; }
    else if (!strcmp(mode, "dense") || !strcmp(mode, "dmma")) dense_th = 1;
    nb->dense_th = dense_th;
    size_t cap = (size_t)env_int("TSG_ROWS_SMEM_KB", 72) * 1024;
    if (cap > c.smem_optin) cap = c.smem_optin;
    nb->smem_cap = (int)cap;
    if (ntr > 0 && numblkC > 0) {
        CK(cudaMemsetAsync(nb->row_kind, 0, (size_t)ntr, c.stream));
        k_s3_classify_tiles<<<ceil_div(numblkC, 256), 256, 0, c.stream>>>((int)numblkC, trow0, C->tile_nnz, C->tile_rowidx, dense_th,
                                                                         nb->dense_list, nb->row_kind, d_ns);
        CK_LAUNCH();
        k_s3_classify_rows<<<ceil_div(ntr, 256), 256, 0, c.stream>>>(ntr, trow0, A->tile_ptr, A->tile_nnz, C->tile_ptr, C->tile_nnz, wptr,
                                                                      light, (int)cap, env_int("TSG_ROWS_MIN_FILL", 8), force_kind,
                                                                      nb->row_kind, d_ns);
        CK_LAUNCH();
    }
    return TSG_OK;
}

int numeric_device(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, int trow0, int ntr, const int *wptr, const PairLists &pl,
                   const NumericBufs &nb, const int *h_ns, bool heavy_rows, tsg_stats *stats)
{
    Ctx &c = ctx();
    const long long numblkC = C->numtile, nnzC = C->nnz;
    if (nnzC <= 0 || numblkC <= 0) return TSG_OK;
    const char *mode = s3_mode();
    const int n_staged = h_ns[NS_ROWS_STAGED], n_gather = h_ns[NS_ROWS_GATHER], n_dense = h_ns[NS_DENSE];
    if (stats) {
        stats->rows_staged = n_staged; stats->rows_gather = n_gather; stats->tiles_dense = n_dense; stats->rows_smem = h_ns[NS_MAXNEED];
        stats->tiles_nonempty = h_ns[NS_NONEMPTY];
    }

    if (n_dense > 0) {
        S3Dense P{n_dense, nb.dense_list, pl.ptr, pl.end, pl.a, pl.b, A->tile_nnz, A->ptr, A->col, A->val, B->tile_nnz, B->ptr, B->col, B->val,
                  C->tile_nnz, C->ptr, C->mask, C->col, C->val};
        const int blocks = ceil_div((long long)n_dense * 32, S3D_WARPS * 32);
        if (!strcmp(mode, "dmma")) k_step3_dmma<<<blocks, S3D_WARPS * 32, 0, c.stream>>>(P);
        else k_step3_dense<<<blocks, S3D_WARPS * 32, 0, c.stream>>>(P);
        CK_LAUNCH();
    }
    if (n_staged > 0) {
        S3Rows P{trow0, nb.dense_th, A->tile_ptr, A->tile_columnidx, A->tile_nnz, A->ptr, A->mask, A->col, A->val,
                 B->tile_ptr, B->tile_columnidx, B->rm2csc, B->tile_nnz, B->ptr, B->mask, B->col, B->val,
                 C->tile_ptr, C->tile_columnidx, C->tile_nnz, C->ptr, C->mask, C->col, C->val, wptr, pl.ptr, pl.end, pl.a, pl.b, pl.slot,
                 nb.row_kind};
        const size_t smem = ((size_t)h_ns[NS_MAXNEED] + 1023) & ~(size_t)1023;
        // 128-thread CTAs when a tile-row has few (tile, row) slots (2D meshes): fewer idle threads, more CTAs per SM
        const bool narrow = numblkC * TS < (long long)ntr * 192;
        if (smem > 48 * 1024) {
            if (narrow) CK(cudaFuncSetAttribute(k_step3_rows<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else CK(cudaFuncSetAttribute(k_step3_rows<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        if (narrow) k_step3_rows<128><<<ntr, 128, smem, c.stream>>>(P);
        else k_step3_rows<256><<<ntr, 256, smem, c.stream>>>(P);
        CK_LAUNCH();
        k_step3_sparse<<<ceil_div((long long)ntr * 32, 256), 256, 0, c.stream>>>(ntr, P);
        CK_LAUNCH();
    }
    if (n_gather > 0) {
        const int g0 = h_ns[NS_GLO], g1 = h_ns[NS_GHI];  // nonzero range spanned by the gathered tile-rows
        if (!arena_reserve(2, arena_need((size_t)((nnzC + 31) >> 5) + 1, 4))) return last_error();
        int *blk2tile = arena_take<int>(2, (size_t)((nnzC + 31) >> 5) + 1);
        if (!blk2tile) return last_error();
        k_blk2tile<<<ceil_div(numblkC, 256), 256, 0, c.stream>>>((int)numblkC, C->tile_nnz, blk2tile);
        CK_LAUNCH();
        const bool all = n_staged == 0 && n_dense == 0;
        S3Gather P{(int)numblkC, (int)nnzC, trow0, nb.dense_th, blk2tile, pl.ptr, pl.end, pl.a, pl.b, A->tile_nnz, A->ptr, A->mask, A->val,
                   B->tile_nnz, B->ptr, B->mask, B->val, C->tile_nnz, C->tile_rowidx, C->ptr, C->mask, C->col, C->val,
                   all ? nullptr : nb.row_kind};
        // consecutive nonzeros per CTA: 8192 when the work per nonzero is even (no heavy tile-rows) and there are
        // enough of them to keep >= 16 CTAs per SM busy; otherwise one 256-thread pass, balanced by the block scheduler
        const long long span = (long long)g1 - g0;
        int chunk = 256;
        if (!heavy_rows && span >= (long long)c.num_sms * 16 * 8192) chunk = 8192;
        const int chunk_env = env_int("TSG_GATHER_CHUNK", 0);
        if (chunk_env >= 256) chunk = chunk_env;
        if (chunk > 256) k_step3_gather<true, 2><<<ceil_div(span, chunk), 256, 0, c.stream>>>(chunk, g0, g1, P);
        else k_step3_gather<false, 2><<<ceil_div(span, 256), 256, 0, c.stream>>>(256, g0, g1, P);
        CK_LAUNCH();
    }
    return TSG_OK;
}

}  // namespace tsg
