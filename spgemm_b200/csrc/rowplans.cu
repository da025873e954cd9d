// rowplans.cu -- tile-ROW templates: step 1 for matrices whose tile-rows repeat themselves.
//
// The recipe plans (plans.cu) use the fact that a structured matrix is made of few distinct tiles. One level up the same
// holds for its tile-rows: on the 27-point stencil all interior tile-rows of C = A*B are translates of each other -- the
// same tile columns relative to the diagonal, the same B tile-rows behind every A tile, the same tile patterns. For such
// a tile-row everything step 1 computes (C's tile columns, the pair lists, and with the patterns also the recipe of every
// C tile) is the representative's, shifted. So:
//   * k_brow_sig / k_brow_verify   B's tile-rows are put into CLASSES first: two tile-rows of B are in one class when their
//                   tile columns relative to the diagonal and their tile patterns agree (64-bit hash into a table like
//                   the recipes' -- atomicCAS, owner = smallest index -- then every tile-row is compared with its class
//                   owner element by element). One pass over B's tiles (3.5 M on config 2), not over the pairs (81 M);
//   * k_row_sig     per tile-row of the slab: the pair count w (step 1 needs it anyway) and a hash of the row's SIGNATURE
//                   -- per A tile (K - I, pattern, class of B's tile-row K) -- inserted into a second table;
//   * k_s1_count / k_s1_fill (spgemm.cu) then run on the REPRESENTATIVE tile-rows only; fill also records, beside every
//     pair, which A tile of the row and which tile of B's tile-row it came from (pair_src);
//   * k_rows_expand copies the C tile counts to the other rows (before the scans and the allocation);
//   * k_rows_instantiate compares every other tile-row's signature with its representative's, element by element (with
//     the classes verified, that is exact; a hash collision raises the fail flag and the whole call is redone without
//     templates), and writes the row's C tile columns, pair lists and recipe ids from the representative's.
// No bitmap, no shared-memory atomics, no sorting and no recipe hashing for the other rows, and B's tile structure is read
// once per B tile instead of four times per pair. Attempted only together with the recipe plans (both operands made of few
// patterns, no heavy tile-row) and when the rows repeat at least 4 times on average. TSG_ROWPLANS=0 switches it off.
// Replaces, on such inputs, what reference src/tilespgemm-cuda.h:10-392 (step 1) computes per tile-row.
#include "common.cuh"
#include "scan.cuh"
#include "kernels.h"
#include "plans.cuh"

namespace tsg {

using namespace plans;

namespace {

constexpr int RPCAP = 1 << 15;   // signature table slots
constexpr int RPMAX = RPCAP / 4; // more distinct tile-row signatures than this => the regular path
constexpr int BCCAP = 1 << 15;   // class table slots (B's tile-rows)
constexpr int BCMAX = BCCAP / 4; // more classes than this => the regular path

struct RowPlanCtx {
    unsigned long long *keys = nullptr;
    int *owner = nullptr, *flags = nullptr, *dense = nullptr, *rep_row = nullptr;
    unsigned long long *bkeys = nullptr;
    int *bowner = nullptr;
    int *ctl = nullptr;  // [0] signatures inserted, [1] fail, [2] classes inserted, [3] class fail
    int device = -1;
};
RowPlanCtx g_rp;

int rp_init()
{
    Ctx &c = ctx();
    if (g_rp.device == c.device && g_rp.keys) return TSG_OK;
    g_rp = RowPlanCtx();
    g_rp.keys = dalloc_n<unsigned long long>(RPCAP);
    g_rp.owner = dalloc_n<int>(RPCAP);
    g_rp.flags = dalloc_n<int>(RPCAP + 1);
    g_rp.dense = dalloc_n<int>(RPCAP + 1);
    g_rp.rep_row = dalloc_n<int>(RPMAX);
    g_rp.bkeys = dalloc_n<unsigned long long>(BCCAP);
    g_rp.bowner = dalloc_n<int>(BCCAP);
    g_rp.ctl = dalloc_n<int>(4);
    if (!g_rp.keys || !g_rp.owner || !g_rp.flags || !g_rp.dense || !g_rp.rep_row || !g_rp.bkeys || !g_rp.bowner || !g_rp.ctl) {
        g_rp = RowPlanCtx();
        return last_error();
    }
    g_rp.device = c.device;
    return TSG_OK;
}

// position-dependent element hash; the signature hash is the SUM of these, so the lanes can add in any order
__device__ __forceinline__ unsigned long long sig_elem(unsigned pos, unsigned long long v)
{
    unsigned long long x = (v + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull ^ ((unsigned long long)(pos + 1u) * 0x94D049BB133111EBull);
    x ^= x >> 29;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

}  // namespace

void rowplans_shutdown() { g_rp = RowPlanCtx(); }

// One warp per tile-row K of B: its class (bclass[K]; -1 for an empty tile-row).
__global__ void __launch_bounds__(256)
k_brow_sig(int tilem, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, const int *__restrict__ pat_b,
           int *__restrict__ bclass, unsigned long long *keys, int *owner, int *ctl)
{
    const int K = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (K >= tilem) return;
    const int b0 = b_tile_ptr[K], b1 = b_tile_ptr[K + 1];
    unsigned long long acc = 0;
    for (int tb = b0 + lane; tb < b1; tb += 32)
        acc += sig_elem((unsigned)(tb - b0), ((unsigned long long)(unsigned)(b_tile_col[tb] - K) << 16) ^ (unsigned)pat_b[tb]);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
    if (lane) return;
    if (b1 == b0) { bclass[K] = -1; return; }
    const int slot = table_insert(keys, BCCAP, mix64(acc, (unsigned long long)(unsigned)(b1 - b0)), ctl + 2, BCMAX, ctl + 3);
    bclass[K] = slot;
    if (slot >= 0 && K < owner[slot]) atomicMin(&owner[slot], K);  // owner only decreases: a stale read costs one atomic
}

// every tile-row of B against the owner of its class, element by element: after this, equal class = equal tile-row
__global__ void __launch_bounds__(256)
k_brow_verify(int tilem, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, const int *__restrict__ pat_b,
              const int *__restrict__ bclass, const int *__restrict__ owner, int *ctl)
{
    const int K = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (K >= tilem) return;
    const int slot = bclass[K];
    if (slot < 0) { if (slot != -1 || b_tile_ptr[K + 1] != b_tile_ptr[K]) { if (lane == 0) ctl[3] = 1; } return; }
    const int Kr = owner[slot];
    if (Kr == K) return;
    const int b0 = b_tile_ptr[K], n = b_tile_ptr[K + 1] - b0, r0 = b_tile_ptr[Kr];
    bool bad = Kr < 0 || Kr >= tilem || b_tile_ptr[Kr + 1] - r0 != n;
    if (!bad)
        for (int o = lane; o < n; o += 32) bad |= b_tile_col[b0 + o] - K != b_tile_col[r0 + o] - Kr || pat_b[b0 + o] != pat_b[r0 + o];
    if (__any_sync(FULL_MASK, bad) && lane == 0) ctl[3] = 1;
}

// One warp per tile-row of the slab: w[i] and the row's signature slot (sig_slot[i]; -1 for a tile-row without pairs).
__global__ void __launch_bounds__(256)
k_row_sig(int trow0, int ntr, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col, const int *__restrict__ pat_a,
          const int *__restrict__ b_tile_ptr, const int *__restrict__ bclass, int *__restrict__ w, int *__restrict__ sig_slot,
          unsigned long long *keys, int *owner, int *ctl, int *__restrict__ sc_err)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= ntr) return;
    const int I = trow0 + i, a0 = a_tile_ptr[I], a1 = a_tile_ptr[I + 1];
    unsigned long long acc = 0;
    long long s = 0;
    for (int ta = a0 + lane; ta < a1; ta += 32) {
        const int K = a_tile_col[ta];
        acc += sig_elem((unsigned)(ta - a0), (((unsigned long long)(unsigned)(K - I) << 32) | (unsigned)bclass[K]) ^
                                                 ((unsigned long long)(unsigned)pat_a[ta] * 0x9E3779B97F4A7C15ull));
        s += b_tile_ptr[K + 1] - b_tile_ptr[K];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        acc += __shfl_xor_sync(FULL_MASK, acc, o);
        s += __shfl_xor_sync(FULL_MASK, s, o);
    }
    if (lane) return;
    if (s > 0x7fffffffll) { atomicOr(sc_err, 1); s = 0x7fffffff; }
    w[i] = (int)s;
    if (s == 0) { sig_slot[i] = -1; return; }
    if (a1 - a0 > 0xFFFF) ctl[1] = 1;  // pair_src keeps the A tile's position in the row in 16 bits: no templates for this slab
    const unsigned long long h = mix64(acc, ((unsigned long long)(unsigned)(a1 - a0) << 32) | (unsigned)s);
    const int slot = table_insert(keys, RPCAP, h, ctl, RPMAX, ctl + 1);
    sig_slot[i] = slot;
    if (slot >= 0 && i < owner[slot]) atomicMin(&owner[slot], i);  // owner only decreases: a stale read costs one atomic
}

__global__ void __launch_bounds__(256)
k_tab_flags(int cap, const int *__restrict__ owner, int *__restrict__ flags)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) flags[s] = owner[s] != NO_OWNER;
}

__global__ void __launch_bounds__(256)
k_tab_reps(int cap, int rmax, const int *__restrict__ owner, const int *__restrict__ dense, int *__restrict__ rep)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && owner[s] != NO_OWNER && dense[s] < rmax) rep[dense[s]] = owner[s];
}

// rep_of[i] = the representative tile-row of row i (itself for a representative, -1 for a row without pairs).
__global__ void __launch_bounds__(256)
k_rows_rep_of(int ntr, const int *__restrict__ sig_slot, const int *__restrict__ owner, int *__restrict__ rep_of)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntr) return;
    const int slot = sig_slot[i];
    rep_of[i] = slot < 0 ? -1 : owner[slot];
}

// After k_s1_count ran on the representatives: every other row takes its representative's C tile count and light flag.
__global__ void __launch_bounds__(256)
k_rows_expand(int ntr, const int *__restrict__ rep_of, int *__restrict__ cnt, uint8_t *__restrict__ light)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntr) return;
    const int r = rep_of[i];
    if (r < 0) { cnt[i] = 0; light[i] = 0; }
    else if (r != i) { cnt[i] = cnt[r]; light[i] = light[r]; }
}

// One warp per tile-row that is not its own representative: verify its signature against the representative's,
// element by element, then write its share of C's tile lists and of the pair lists from the representative's -- in the
// order of the pair lists, so the stores are coalesced: pair q of the
// representative came from (A tile ta of the row, tile o of B's tile-row), so pair q of this row is
// (a0 + ta, rm2csc[first tile of B's tile-row behind a0 + ta, + o]).
struct RowInst {
    int trow0, ntr;
    const int *rep_of, *w, *wptr, *c_tile_ptr;
    const int *a_tile_ptr, *a_tile_col, *pat_a, *b_tile_ptr, *bclass, *b_rm2csc;
    const unsigned *pair_src;
    int *c_tile_col, *c_tile_row, *pair_ptr, *pair_end, *pair_a, *pair_b, *recipe_id;
    int *fail;
};

constexpr int RI_STAGE = 64;  // first tiles of the B tile-rows behind up to this many A tiles are kept in shared memory

__global__ void __launch_bounds__(256)
k_rows_instantiate(const __grid_constant__ RowInst P)
{
    __shared__ int s_b0[8][RI_STAGE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * 8 + warp;
    if (i >= P.ntr) return;
    const int r = P.rep_of[i];
    if (r < 0 || r == i) return;
    const int I = P.trow0 + i, Ir = P.trow0 + r, shift = I - Ir;
    const int a0 = P.a_tile_ptr[I], nA = P.a_tile_ptr[I + 1] - a0, ar0 = P.a_tile_ptr[Ir];
    const int wbase = P.wptr[i], wbr = P.wptr[r], w = P.w[i];
    if (nA != P.a_tile_ptr[Ir + 1] - ar0 || w != P.w[r]) {  // not the same row after all: nothing of it can be trusted
        if (lane == 0) *P.fail = 1;
        return;
    }
    const int cbase = P.c_tile_ptr[i], numJ = P.c_tile_ptr[i + 1] - cbase, cbr = P.c_tile_ptr[r];
    for (int s = lane; s < numJ; s += 32) {
        P.c_tile_col[cbase + s] = P.c_tile_col[cbr + s] + shift;
        P.c_tile_row[cbase + s] = I;
        P.pair_ptr[cbase + s] = P.pair_ptr[cbr + s] - wbr + wbase;
        P.pair_end[cbase + s] = P.pair_end[cbr + s] - wbr + wbase;
        P.recipe_id[cbase + s] = P.recipe_id[cbr + s];
    }
    bool bad = false;
    for (int t = lane; t < nA; t += 32) {  // the two signatures, element by element (equal class = equal B tile-row: k_brow_verify)
        const int K = P.a_tile_col[a0 + t], Kr = P.a_tile_col[ar0 + t];
        bad |= K - I != Kr - Ir || P.pat_a[a0 + t] != P.pat_a[ar0 + t] || P.bclass[K] != P.bclass[Kr];
        if (t < RI_STAGE) s_b0[warp][t] = P.b_tile_ptr[K];
    }
    if (__any_sync(FULL_MASK, bad)) {  // a 64-bit collision: the call is redone without templates; write nothing that could run off
        if (lane == 0) *P.fail = 1;
        return;
    }
    __syncwarp();
    for (int q = lane; q < w; q += 32) {
        const unsigned src = P.pair_src[wbr + q];
        const int ta = (int)(src >> 16), o = (int)(src & 0xFFFFu);
        const int b0 = ta < RI_STAGE ? s_b0[warp][ta] : P.b_tile_ptr[P.a_tile_col[a0 + ta]];
        P.pair_a[wbase + q] = a0 + ta;
        P.pair_b[wbase + q] = P.b_rm2csc[b0 + o];
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
bool rowplans_env_on()
{
    const char *e = getenv("TSG_ROWPLANS");
    return !(e && *e == '0');
}

// Step 1a with templates: w and the signature of every tile-row, the representatives (dense numbering), rep_of.
// One read-back (number of signatures, fail). Returns *nsig = 0 when the regular path should run instead.
int rowplans_signatures(const tsg_dtile *A, const tsg_dtile *B, int trow0, int ntr, int *w, int *sig_slot, int *rep_of, int *bclass, int *sc_err,
                        const int **rep_list, int *nsig)
{
    Ctx &c = ctx();
    *nsig = 0;
    int rc = rp_init();
    if (rc) return rc;
    RowPlanCtx &p = g_rp;
    CK(cudaMemsetAsync(p.keys, 0, (size_t)RPCAP * 8, c.stream));
    CK(cudaMemsetAsync(p.owner, 0x7f, (size_t)RPCAP * 4, c.stream));
    CK(cudaMemsetAsync(p.bkeys, 0, (size_t)BCCAP * 8, c.stream));
    CK(cudaMemsetAsync(p.bowner, 0x7f, (size_t)BCCAP * 4, c.stream));
    CK(cudaMemsetAsync(p.ctl, 0, 4 * sizeof(int), c.stream));
    k_brow_sig<<<ceil_div(B->tilem, 8), 256, 0, c.stream>>>(B->tilem, B->tile_ptr, B->tile_columnidx, B->pat, bclass, p.bkeys, p.bowner, p.ctl);
    CK_LAUNCH();
    k_brow_verify<<<ceil_div(B->tilem, 8), 256, 0, c.stream>>>(B->tilem, B->tile_ptr, B->tile_columnidx, B->pat, bclass, p.bowner, p.ctl);
    CK_LAUNCH();
    k_row_sig<<<ceil_div(ntr, 8), 256, 0, c.stream>>>(trow0, ntr, A->tile_ptr, A->tile_columnidx, A->pat, B->tile_ptr, bclass, w, sig_slot, p.keys,
                                                      p.owner, p.ctl, sc_err);
    CK_LAUNCH();
    k_tab_flags<<<ceil_div(RPCAP, 256), 256, 0, c.stream>>>(RPCAP, p.owner, p.flags);
    CK_LAUNCH();
    rc = exclusive_scan<int>(p.flags, p.dense, RPCAP);
    if (rc) return rc;
    k_tab_reps<<<ceil_div(RPCAP, 256), 256, 0, c.stream>>>(RPCAP, RPMAX, p.owner, p.dense, p.rep_row);
    k_rows_rep_of<<<ceil_div(ntr, 256), 256, 0, c.stream>>>(ntr, sig_slot, p.owner, rep_of);
    CK_LAUNCH();
    rc = publish_words(&c.h_scalars[24], p.dense + RPCAP, 1);
    if (!rc) rc = publish_words(&c.h_scalars[25], p.ctl, 4);  // two 8-byte slots: [25] = count, fail; [26] = class count, class fail
    if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    const int n = *(const volatile int *)&c.h_scalars[24];
    const volatile int *ctl = (const volatile int *)&c.h_scalars[25];
    if (ctl[1] || ctl[3] || n <= 0 || n > RPMAX || (long long)n * 4 > ntr) return TSG_OK;  // rows do not repeat enough: the regular path
    *nsig = n;
    *rep_list = p.rep_row;
    return TSG_OK;
}

int rowplans_expand_counts(int ntr, const int *rep_of, int *cnt, uint8_t *light)
{
    Ctx &c = ctx();
    k_rows_expand<<<ceil_div(ntr, 256), 256, 0, c.stream>>>(ntr, rep_of, cnt, light);
    CK_LAUNCH();
    return TSG_OK;
}

int *rowplans_fail_ptr() { return g_rp.ctl ? g_rp.ctl + 1 : nullptr; }

int rowplans_instantiate(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, const RowTemplates &rt, int *recipe_id)
{
    Ctx &c = ctx();
    RowInst P{rt.trow0, rt.ntr, rt.rep_of, rt.w, rt.wptr, C->tile_ptr, A->tile_ptr, A->tile_columnidx, A->pat, B->tile_ptr, rt.bclass,
              B->rm2csc, rt.pair_src, C->tile_columnidx, C->tile_rowidx, rt.pair_ptr, rt.pair_end, rt.pair_a, rt.pair_b, recipe_id, g_rp.ctl + 1};
    k_rows_instantiate<<<ceil_div(rt.ntr, 8), 256, 0, c.stream>>>(P);
    CK_LAUNCH();
    if (getenv("TSG_ROWPLANS_FORCE_FAIL")) CK(cudaMemsetAsync(g_rp.ctl + 1, 1, 1, c.stream));  // tests: the redo-without-templates path
    return TSG_OK;
}

}  // namespace tsg
