// plans.cu -- recipe plans: a bit-exact fast path of steps 2 and 3 for matrices whose tiles repeat themselves.
//
// Structured inputs are made of few distinct tiles: the 27-point stencil on a 128^3 grid has 3 distinct tile patterns (the
// 16 row masks of a tile) and its 13.7 M listed C tiles follow 42 distinct "recipes" (the sequence of (A pattern, B
// pattern) over a C tile's pairs); lap2d 256^2: 4 / 20; block-FEM 2M: 12 / 22; a 100^3 stencil: 22 / 408. A recipe fixes
// everything the symbolic step computes (mask, Ptr, nnz of the C tile) and, per C nonzero, the list of sources
// (pair, position in A's tile, position in B's tile) in the serial SPA's order. So:
//   * csr2tile gives every tile a pattern id (k_pattern_insert / _verify: device hash table, full-key verification);
//   * k_s1_fill hashes each C tile's pair sequence while it writes the pair lists (spgemm.cu) and inserts it into the
//     recipe table; k_recipe_verify compares every tile's sequence with its recipe's representative (with tile-row
//     templates, rowplans.cu, both see the representative tile-rows only and the other rows copy their recipe ids);
//   * k_plan_build plans each distinct recipe once, from the representative tile's actual masks;
//   * k_plan_slots packs each recipe's C nonzeros into "slots" (chains of nonzeros) of nearly equal product count;
//   * the symbolic step becomes a 68-byte copy per C tile (k_symbolic_from_plans) and the numeric step walks an
//     L1-resident plan, one lane per slot, every iteration a product, each nonzero summed in the serial SPA's order
//     (k_numeric_from_plans_rows): bit-identical to the generic kernels' results.
// Nothing here waits on another thread: insert = one atomicCAS on the key word, owner of a slot = atomicMin of the item
// indices that landed in it (deterministic), and a true 64-bit hash collision, too many patterns / recipes, or a plan
// that outgrows its buffer raise the fail flag -- the generic kernels (spgemm.cu step 2, numeric.cu) then run instead.
// Irregular matrices (R-MAT: 891 020 recipes for 933 730 C tiles at scale 14) never get here: the path is attempted only
// when both operands hold few patterns (PLANS_MAX_PATTERNS) and no tile-row is heavy. TSG_PLANS=0 switches it off.
// Replaces, on such inputs, what reference src/tilespgemm-cuda.h:394-773 (symbolic) and :1273-1952 (numeric) compute.
#include "common.cuh"
#include "scan.cuh"
#include "kernels.h"
#include "plans.cuh"

namespace tsg {

using namespace plans;

// ---------------------------------------------------------------------------------------------
// Pattern ids (per tiled matrix, at csr2tile / upload time)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pattern_insert(int numtile, const uint16_t *__restrict__ mask, unsigned long long *keys, int *owner, int *count,
                 int *__restrict__ pat_id, int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numtile) return;
    const uint4 *mp = reinterpret_cast<const uint4 *>(mask + (size_t)t * TS);
    const uint4 x = mp[0], y = mp[1];
    unsigned long long h = 0x243F6A8885A308D3ull;
    h = mix64(h, ((unsigned long long)x.x << 32) | x.y); h = mix64(h, ((unsigned long long)x.z << 32) | x.w);
    h = mix64(h, ((unsigned long long)y.x << 32) | y.y); h = mix64(h, ((unsigned long long)y.z << 32) | y.w);
    const int slot = table_insert(keys, PCAP, h, count, PCAP / 2, fail);
    pat_id[t] = slot;
    if (slot >= 0 && t < owner[slot]) atomicMin(&owner[slot], t);  // owner only ever decreases: a stale read costs one atomic
}

__global__ void __launch_bounds__(256)
k_pattern_verify(int numtile, const uint16_t *__restrict__ mask, const int *__restrict__ pat_id, const int *__restrict__ owner, int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numtile || *(volatile int *)fail) return;
    const int slot = pat_id[t];
    if (slot < 0) { *fail = 1; return; }
    const uint4 *op = reinterpret_cast<const uint4 *>(mask + (size_t)owner[slot] * TS);
    const uint4 *mp = reinterpret_cast<const uint4 *>(mask + (size_t)t * TS);
    const uint4 x = mp[0], y = mp[1], a = op[0], b = op[1];
    if (!(a.x == x.x && a.y == x.y && a.z == x.z && a.w == x.w && b.x == y.x && b.y == y.y && b.z == y.z && b.w == y.w)) *fail = 2;
}

constexpr int RP_TMAX = 256;             // tile-row templates that get a row plan (more: the per-recipe plans are walked)
constexpr int RP_SCAP = 2048;            // slots of one tile-row
constexpr int RP_GCAP = RP_SCAP / 32;    // groups of 32 slots = one warp's share
constexpr int RP_WCAP = 1 << 15;         // plan words of one tile-row template

struct PlanCtx {
    // pattern table scratch (reused by every csr2tile)
    unsigned long long *pkeys = nullptr;
    int *powner = nullptr;
    // recipe table + plans (reused by every spgemm call)
    unsigned long long *rkeys = nullptr;
    int *rowner = nullptr, *rflags = nullptr, *rdense = nullptr, *rep_tile = nullptr;
    uint16_t *plan_mask = nullptr, *plan_ptr = nullptr;
    int *plan_nnz = nullptr, *plan_tot = nullptr, *plan_off = nullptr;
    unsigned *plan_ent = nullptr;
    uint16_t *plan_cnt = nullptr;
    uint8_t *plan_col = nullptr;
    unsigned *plan_slot = nullptr;   // per recipe, 128 slots: products | first chain entry << 16
    unsigned *plan_jslot = nullptr;  // per recipe, 256 nonzeros: slot | first iteration << 8
    uint16_t *plan_chain = nullptr;  // per recipe, 256 chain entries: position in the tile | column << 8
    int *plan_nslots = nullptr;
    int *ctl = nullptr;  // [0] pattern count, [1] pattern fail, [2] recipe count, [3] recipe / plan fail, [4] row plan fail
    // row plans (k_numeric_from_rowplans): per tile-row template, the slots of the whole tile-row and their plan words laid out
    // warp by warp
    int4 *rp_rec = nullptr;      // [RP_TMAX][RP_SCAP] per slot: products, first chain entry, C nnz offset of its tile in the row, pair offset
    int2 *rp_src = nullptr;      // [RP_TMAX][RP_SCAP] per slot: first plan word, stride (the recipe's slot count)
    unsigned *rp_words = nullptr;  // [RP_TMAX][RP_WCAP] plan words, group (32 slots) after group, iteration-major inside a group
    int *rp_goff = nullptr;      // [RP_TMAX][RP_GCAP + 1] first word of every group
    int *rp_nslots = nullptr;    // [RP_TMAX]
    int device = -1;
};
static PlanCtx g_plan;

static int plan_ctx_init()
{
    Ctx &c = ctx();
    if (g_plan.device == c.device && g_plan.pkeys) return TSG_OK;
    g_plan = PlanCtx();
    PlanCtx &p = g_plan;
    p.pkeys = dalloc_n<unsigned long long>(PCAP);
    p.powner = dalloc_n<int>(PCAP);
    p.rkeys = dalloc_n<unsigned long long>(RCAP);
    p.rowner = dalloc_n<int>(RCAP);
    p.rflags = dalloc_n<int>(RCAP + 1);
    p.rdense = dalloc_n<int>(RCAP + 1);
    p.rep_tile = dalloc_n<int>(RMAX);
    p.plan_mask = dalloc_n<uint16_t>((size_t)RMAX * TS);
    p.plan_ptr = dalloc_n<uint16_t>((size_t)RMAX * TS);
    p.plan_nnz = dalloc_n<int>(RMAX);
    p.plan_tot = dalloc_n<int>(RMAX + 1);
    p.plan_off = dalloc_n<int>(RMAX + 1);
    p.plan_cnt = dalloc_n<uint16_t>((size_t)RMAX * 256);
    p.plan_col = dalloc_n<uint8_t>((size_t)RMAX * 256);
    p.plan_slot = dalloc_n<unsigned>((size_t)RMAX * 128);
    p.plan_jslot = dalloc_n<unsigned>((size_t)RMAX * 256);
    p.plan_chain = dalloc_n<uint16_t>((size_t)RMAX * 256);
    p.plan_nslots = dalloc_n<int>(RMAX);
    p.plan_ent = dalloc_n<unsigned>(PLAN_ENT_CAP);
    p.ctl = dalloc_n<int>(8);
    if (!p.pkeys || !p.powner || !p.rkeys || !p.rowner || !p.rflags || !p.rdense || !p.rep_tile || !p.plan_mask || !p.plan_ptr ||
        !p.plan_nnz || !p.plan_tot || !p.plan_off || !p.plan_cnt || !p.plan_col || !p.plan_slot || !p.plan_jslot || !p.plan_chain || !p.plan_nslots || !p.plan_ent || !p.ctl) {
        g_plan = PlanCtx();
        return last_error();
    }
    p.device = c.device;
    return TSG_OK;
}

void plans_shutdown() { g_plan = PlanCtx(); }  // the buffers belong to the stream-ordered pool, which tsg_shutdown releases

static bool plans_env_on()
{
    const char *e = getenv("TSG_PLANS");
    return !(e && *e == '0');
}

__global__ void k_gather_int(int n, const int *__restrict__ idx, const int *__restrict__ src, int *__restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[idx[t]];
}

// Pattern id of every tile of T into T->pat -- in ROW-MAJOR tile order (for a column-major T that is the order of
// rm2csc, the order step 1 walks B's tile-rows in: k_s1_fill reads the ids beside B's tile columns, coalesced) -- and the
// number of distinct patterns into T->npat (-1 when there are more than PCAP/2 or a 64-bit hash collision was detected).
// One read-back; runs at csr2tile / upload time.
int tile_patterns_device(tsg_dtile *T)
{
    Ctx &c = ctx();
    T->npat = -1;
    if (!T->pat || T->numtile <= 0 || !plans_env_on()) return TSG_OK;
    int rc = plan_ctx_init();
    if (rc) return rc;
    PlanCtx &p = g_plan;
    CK(cudaMemsetAsync(p.pkeys, 0, (size_t)PCAP * 8, c.stream));
    CK(cudaMemsetAsync(p.powner, 0x7f, (size_t)PCAP * 4, c.stream));
    CK(cudaMemsetAsync(p.ctl, 0, 2 * sizeof(int), c.stream));
    const int blocks = ceil_div(T->numtile, 256);
    const bool permute = T->col_major && T->rm2csc;
    int *ids = permute ? dalloc_n<int>((size_t)T->numtile) : T->pat;  // storage order first
    if (!ids) return last_error();
    k_pattern_insert<<<blocks, 256, 0, c.stream>>>(T->numtile, T->mask, p.pkeys, p.powner, p.ctl, ids, p.ctl + 1);
    k_pattern_verify<<<blocks, 256, 0, c.stream>>>(T->numtile, T->mask, ids, p.powner, p.ctl + 1);
    if (permute) {
        k_gather_int<<<blocks, 256, 0, c.stream>>>(T->numtile, T->rm2csc, ids, T->pat);
        dfree(ids);  // stream-ordered: released after the gather
    }
    CK_LAUNCH();
    rc = publish_words(&c.h_scalars[13], p.ctl, 2);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    const int *h = (const int *)&c.h_scalars[13];
    T->npat = h[1] ? -1 : h[0];
    return TSG_OK;
}

bool plans_wanted(const tsg_dtile *A, const tsg_dtile *B)
{
    const char *forced = getenv("TSG_STEP3");  // a forced numeric kernel (A/B measurements, tests) means the generic path
    if (forced && *forced && strcmp(forced, "auto")) return false;
    // matrices made of well-filled tiles only (block-FEM: 96 entries per A tile) are the dense accumulator's: measured on
    // config 4, the plan path is 4 % slower than k_step3_dense there (552.8 vs 573.8 GFLOP/s, profiles/README.md r2t / r2s);
    // a half-and-half mix of block-FEM and stencil tiles (45 entries per tile on average) is 2.1x FASTER through the plans
    if (!(getenv("TSG_PLANS") && *getenv("TSG_PLANS") == '2') && A->nnz >= 64ll * A->numtile) return false;
    return plans_env_on() && A->pat && B->pat && A->npat > 0 && B->npat > 0 && A->npat <= PLANS_MAX_PATTERNS && B->npat <= PLANS_MAX_PATTERNS;
}

// ---------------------------------------------------------------------------------------------
// Recipes and plans (per spgemm call, after k_s1_fill has inserted every C tile's recipe hash)
// ---------------------------------------------------------------------------------------------
// flags[slot] = 1 where the slot has an owner; scanned into dense recipe numbers (rdense).
__global__ void __launch_bounds__(256)
k_recipe_flags(const int *__restrict__ owner, int *__restrict__ flags)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < RCAP) flags[s] = owner[s] != NO_OWNER;
}

__global__ void __launch_bounds__(256)
k_recipe_reps(const int *__restrict__ owner, const int *__restrict__ rdense, int *__restrict__ rep_tile)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < RCAP && owner[s] != NO_OWNER && rdense[s] < RMAX) rep_tile[rdense[s]] = owner[s];
}

// every C tile compares its (A pattern, B pattern) sequence (pair_pat, written by k_s1_fill beside the pair lists) with
// its recipe's representative: a 64-bit collision fails.
__global__ void __launch_bounds__(256)
k_recipe_verify(int numblkC, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end, const unsigned *__restrict__ pair_pat,
                const int *__restrict__ rslot, const int *__restrict__ owner, const int *__restrict__ rdense,
                int *__restrict__ recipe_id, int *fail, int skip_unset)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numblkC) return;
    const int slot = rslot[t];
    if (skip_unset && slot == -1) return;  // tile-row templates: this tile's row takes its recipes from its representative
    recipe_id[t] = 0;
    if (*(volatile int *)fail) return;
    if (slot < 0 || slot >= RCAP) { *fail = 1; return; }
    const int u = owner[slot];
    if (u < 0 || u >= numblkC || rdense[slot] >= RMAX) { *fail = 1; return; }
    const int p0 = pair_ptr[t], n = pair_end[t] - p0, q0 = pair_ptr[u];
    bool same = pair_end[u] - q0 == n;
    if (u != t)
        for (int i = 0; i < n && same; i++) same = pair_pat[p0 + i] == pair_pat[q0 + i];
    if (!same) *fail = 2;
    recipe_id[t] = rdense[slot];
}

// A HALF-WARP per distinct recipe, lane = row r of the recipe's representative C tile. FILL = false: masks / Ptr / nnz of
// the tile and the number of products of every C nonzero (plan_cnt). k_plan_slots then pairs the nonzeros into slots and
// sizes the recipe's entries (plan_tot). FILL = true: plan_off = exclusive scan of plan_tot; writes the entries
// (last product of its nonzero << 31 | pair index << 16 | position in B's tile << 8 | position in A's tile), each
// nonzero's in the serial SPA's order, ITERATION-MAJOR over slots: product i of the nonzero that starts at iteration
// `start` of slot v lives at plan_off[R] + (start + i) * nslots + v, so the lanes of the numeric kernel -- consecutive
// slots of a tile, all at the same iteration -- read consecutive words.
constexpr int PB_SPLIT = 4;
template <bool FILL>
__global__ void __launch_bounds__(128)
k_plan_build(const int *__restrict__ nrec_p, const int *__restrict__ rep_tile, const int *__restrict__ pair_ptr,
             const int *__restrict__ pair_end, const int *__restrict__ pair_a, const int *__restrict__ pair_b,
             const uint16_t *__restrict__ a_mask, const uint16_t *__restrict__ a_ptr, const uint16_t *__restrict__ b_mask,
             const uint16_t *__restrict__ b_ptr, uint16_t *plan_mask, uint16_t *plan_ptr, int *plan_nnz,
             const int *__restrict__ plan_off, uint16_t *plan_cnt, uint8_t *plan_col, const unsigned *__restrict__ plan_jslot,
             const int *__restrict__ plan_nslots, unsigned *plan_ent, int *fail)
{
    // PB_SPLIT half-warps per recipe: each computes the row masks and offsets (cheap), then walks every PB_SPLIT-th C nonzero of
    // its rows -- the chain of dependent loads per thread is what this kernel's time is made of (42 recipes on config 2: 21
    // warps in all, 43 + 60 us for the two passes before the split, a quarter of the non-numeric time of a 64^3-sized slab)
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = gt / (16 * PB_SPLIT), r = threadIdx.x & 15, sub = (gt >> 4) % PB_SPLIT;
    const unsigned hm = 0xFFFFu << (threadIdx.x & 16);
    const int nrec = *nrec_p;
    if (*(volatile int *)fail) return;
    if (nrec > RMAX) { if (R == 0 && r == 0) *fail = 1; return; }
    if (R >= nrec) return;  // whole half-warps leave together
    if (FILL && plan_off[nrec] > PLAN_ENT_CAP) { if (R == 0 && r == 0) *fail = 3; return; }
    const int t = rep_tile[R];
    const int p0 = pair_ptr[t], p1 = pair_end[t];
    if (p1 - p0 > 0x7FFF) { *fail = 1; return; }  // the plan entry keeps the pair index in 15 bits
    unsigned cm = 0;
    for (int p = p0; p < p1; p++) {
        const int a = pair_a[p], b = pair_b[p];
        unsigned am = a_mask[(size_t)a * TS + r];
        while (am) {
            const int k = __clz(am) - 16;
            am ^= 0x8000u >> k;
            cm |= b_mask[(size_t)b * TS + k];
        }
    }
    const int n = __popc(cm);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(hm, incl, o, 16);
        if (r >= o) incl += v;
    }
    const int rowbase = incl - n, nnz = __shfl_sync(hm, incl, 15, 16);
    if (!FILL && sub == 0) {
        plan_ptr[R * TS + r] = (uint16_t)rowbase;
        plan_mask[R * TS + r] = (uint16_t)cm;
        if (r == 0) plan_nnz[R] = nnz;
    }
    const unsigned base = FILL ? (unsigned)plan_off[R] : 0u, nslots = FILL ? (unsigned)plan_nslots[R] : 0u;
    int j = rowbase;
    unsigned rowm = cm;
    while (rowm) {
        const int c = __clz(rowm) - 16;
        rowm ^= 0x8000u >> c;
        const unsigned cbit = 0x8000u >> c;
        if ((j - rowbase) % PB_SPLIT != sub) { j++; continue; }  // another half-warp's nonzero
        const unsigned js = FILL ? plan_jslot[(size_t)R * 256 + j] : 0u;  // slot | first iteration << 8
        unsigned i = js >> 8;
        const unsigned i0 = i, ilast = FILL ? i0 + plan_cnt[(size_t)R * 256 + j] - 1u : 0u;
        for (int p = p0; p < p1; p++) {
            const int a = pair_a[p], b = pair_b[p];
            unsigned am = a_mask[(size_t)a * TS + r];
            unsigned ia = a_ptr[(size_t)a * TS + r];
            while (am) {
                const int k = __clz(am) - 16;
                am ^= 0x8000u >> k;
                const unsigned bm = b_mask[(size_t)b * TS + k];
                if (bm & cbit) {
                    if (FILL) {
                        const unsigned posb = (unsigned)b_ptr[(size_t)b * TS + k] + __popc(bm >> (16 - c));
                        plan_ent[base + i * nslots + (js & 255u)] = ((unsigned)(i == ilast) << 31) | ((unsigned)(p - p0) << 16) | (posb << 8) | ia;
                    }
                    i++;
                }
                ia++;
            }
        }
        if (!FILL) {
            if (i - i0 > 0xFFFFu) *fail = 1;
            plan_cnt[(size_t)R * 256 + j] = (uint16_t)(i - i0);
            plan_col[(size_t)R * 256 + j] = (uint8_t)c;
        }
        j++;
    }
}

// A WARP per distinct recipe: the C nonzeros of the tile differ in how many products they sum (1 ... 27 on the 27-point
// stencil), and a warp whose lanes each walked one nonzero would run as long as its longest list with, measured, 15 of 32
// lanes busy. So the nonzeros are packed into SLOTS -- chains of nonzeros walked one after the other by one lane of the
// numeric kernel -- of nearly equal total length: nslots = ceil(all products / T) with T = max(the longest list, tmin),
// filled longest-first into the slot with the least so far (LPT). Per slot: plan_slot = products | first chain entry << 16.
// Per chain entry (slot after slot, in walking order): plan_chain = position of the nonzero in the tile | its column << 8.
// Per nonzero: plan_jslot = slot | first iteration << 8. plan_tot = nslots * the longest slot = words of plan entries.
__global__ void __launch_bounds__(128)
k_plan_slots(const int *__restrict__ nrec_p, int tmin, const int *__restrict__ plan_nnz, const uint16_t *__restrict__ plan_cnt,
             const uint8_t *__restrict__ plan_col, unsigned *__restrict__ plan_jslot, unsigned *__restrict__ plan_slot,
             uint16_t *__restrict__ plan_chain, int *__restrict__ plan_nslots, int *__restrict__ plan_tot, int *fail)
{
    __shared__ uint16_t s_cnt[4][256], s_start[4][256];
    __shared__ uint8_t s_sorted[4][256], s_bin[4][256], s_k[4][256];
    __shared__ int s_load[4][128], s_len[4][128];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, R = blockIdx.x * 4 + w;
    const int nrec = *nrec_p;
    if (*(volatile int *)fail || nrec > RMAX || R >= nrec) return;  // whole warps leave together
    const int nnz = plan_nnz[R];
    int total = 0;
    for (int j = lane; j < nnz; j += 32) { const int cj = plan_cnt[(size_t)R * 256 + j]; s_cnt[w][j] = (uint16_t)cj; total += cj; }
#pragma unroll
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(FULL_MASK, total, o);
    __syncwarp();
    for (int j = lane; j < nnz; j += 32) {  // rank = how many nonzeros come before j in (count descending, index ascending) order
        const int cj = s_cnt[w][j];
        int rank = 0;
        for (int k = 0; k < nnz; k++) {
            const int ck = s_cnt[w][k];
            rank += (ck > cj) || (ck == cj && k < j);
        }
        s_sorted[w][rank] = (uint8_t)j;
    }
    for (int b = lane; b < 128; b += 32) { s_load[w][b] = 0; s_len[w][b] = 0; }
    __syncwarp();
    int nslots = 0;
    if (nnz > 0) {
        const int T = max((int)s_cnt[w][s_sorted[w][0]], tmin);
        nslots = min(128, (total + T - 1) / T);  // <= nnz: every list holds at least one product
    }
    for (int r = 0; r < nnz; r++) {  // LPT: the next-longest list goes to the slot with the fewest products so far
        unsigned best = 0xFFFFFFFFu;
        for (int b = lane; b < nslots; b += 32) best = min(best, ((unsigned)s_load[w][b] << 8) | (unsigned)b);
#pragma unroll
        for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(FULL_MASK, best, o));
        if (lane == 0) {
            const int j = s_sorted[w][r], b = (int)(best & 255u), ld = s_load[w][b];
            if (ld + (int)s_cnt[w][j] > 0xFFFF) *fail = 1;
            s_bin[w][j] = (uint8_t)b;
            s_start[w][j] = (uint16_t)ld;
            s_k[w][j] = (uint8_t)s_len[w][b];
            s_load[w][b] = ld + s_cnt[w][j];
            s_len[w][b]++;
        }
        __syncwarp();
    }
    int longest = 0, carry = 0;
    for (int b0 = 0; b0 < nslots; b0 += 32) {  // first chain entry of every slot: exclusive scan of the chain lengths
        const int b = b0 + lane, len = b < nslots ? s_len[w][b] : 0;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (b < nslots) {
            const int first = carry + incl - len;
            plan_slot[(size_t)R * 128 + b] = (unsigned)s_load[w][b] | ((unsigned)first << 16);
            longest = max(longest, s_load[w][b]);
            s_len[w][b] = first;
        }
        carry += __shfl_sync(FULL_MASK, incl, 31);
    }
    __syncwarp();
    for (int j = lane; j < nnz; j += 32) {
        const int b = s_bin[w][j];
        plan_jslot[(size_t)R * 256 + j] = (unsigned)b | ((unsigned)s_start[w][j] << 8);
        plan_chain[(size_t)R * 256 + s_len[w][b] + s_k[w][j]] = (uint16_t)(j | ((int)plan_col[(size_t)R * 256 + j] << 8));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) longest = max(longest, __shfl_xor_sync(FULL_MASK, longest, o));
    if (lane == 0) { plan_nslots[R] = nslots; plan_tot[R] = longest * nslots; }
}

// C tile metadata from the plan: one thread per C tile copies its recipe's 32 + 32 bytes and its nnz.
__global__ void __launch_bounds__(256)
k_symbolic_from_plans(int numblkC, const int *__restrict__ recipe_id, const uint16_t *__restrict__ plan_mask,
                      const uint16_t *__restrict__ plan_ptr, const int *__restrict__ plan_nnz, uint16_t *__restrict__ c_mask,
                      uint16_t *__restrict__ c_ptr, int *__restrict__ c_cnt, const int *__restrict__ fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numblkC) return;
    if (*fail) { c_cnt[t] = 0; return; }  // the generic symbolic step will overwrite everything
    const int R = recipe_id[t];
    const uint4 *sm = reinterpret_cast<const uint4 *>(plan_mask + (size_t)R * TS), *sp = reinterpret_cast<const uint4 *>(plan_ptr + (size_t)R * TS);
    uint4 *dm = reinterpret_cast<uint4 *>(c_mask + (size_t)t * TS), *dp = reinterpret_cast<uint4 *>(c_ptr + (size_t)t * TS);
    dm[0] = sm[0]; dm[1] = sm[1];
    dp[0] = sp[0]; dp[1] = sp[1];
    c_cnt[t] = plan_nnz[R];
}

// The numeric step: a CTA per C TILE-ROW. The tile-row's share of A's values (contiguous in row-major tile order), the
// offsets / recipes / pair ranges / slot ranges of its C tiles and, per pair, the position of the A tile's values in that
// staged copy and the base of the B tile's values are put into shared memory first. Then one lane per SLOT of the
// tile-row (k_plan_slots: a chain of C nonzeros of one tile): every iteration is a product -- two global loads (the plan
// word, coalesced, and B's value) and two shared ones (the pair's bases, A's value) -- and the product that ends a
// nonzero's list (bit 31 of the plan word) stores the sum and moves on to the chain's next nonzero. Each nonzero is summed
// in the serial SPA's order, so the values are bit-identical to the generic kernels'.
// Tile-rows that do not fit `smem_cap` take one lane per nonzero and read everything from global memory (same results).
struct PlanRows {
    int trow0, smem_cap;
    const int *a_tile_ptr, *a_tile_nnz;
    const double *a_val;
    const int *b_tile_nnz;
    const double *b_val;
    const int *c_tile_ptr, *c_tile_nnz, *wptr, *pair_ptr, *pair_a, *pair_b, *recipe_id;
    const int *plan_off, *plan_nslots;
    const uint16_t *plan_cnt, *plan_chain;
    const uint8_t *plan_col;
    const unsigned *plan_slot, *plan_jslot, *plan_ent;
    uint16_t *c_col;
    double *c_val;
};

__host__ __device__ __forceinline__ size_t plan_rows_need(int nnzA, int numJ, int W)
{
    return (((size_t)nnzA * 8 + 15) & ~(size_t)15) + 2 * (((size_t)(numJ + 1) * 4 + 15) & ~(size_t)15) + 2 * (((size_t)numJ * 4 + 15) & ~(size_t)15) +
           (((size_t)W * 8 + 15) & ~(size_t)15);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_numeric_from_plans_rows(const __grid_constant__ PlanRows P)
{
    extern __shared__ __align__(16) unsigned char pr_smem[];
    const int i = blockIdx.x, tid = threadIdx.x, I = P.trow0 + i;
    const int c0 = P.c_tile_ptr[i], numJ = P.c_tile_ptr[i + 1] - c0;
    if (numJ == 0) return;
    const int n0 = P.c_tile_nnz[c0], nnzC = P.c_tile_nnz[c0 + numJ] - n0;
    if (nnzC == 0) return;
    const int a0 = P.a_tile_ptr[I], a1 = P.a_tile_ptr[I + 1];
    const int av0 = P.a_tile_nnz[a0], nnzA = P.a_tile_nnz[a1] - av0;
    const int w0 = P.wptr[i], W = P.wptr[i + 1] - w0;
    if (plan_rows_need(nnzA, numJ, W) > (size_t)P.smem_cap) {  // uniform over the CTA: lane per nonzero, nothing staged
        for (int g = tid; g < nnzC; g += THREADS) {
            int lo = 0, hi = numJ - 1;  // the tile holding nonzero g of the tile-row: largest s with nnz offset <= g
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (P.c_tile_nnz[c0 + mid] - n0 <= g) lo = mid; else hi = mid - 1;
            }
            const int off = g - (P.c_tile_nnz[c0 + lo] - n0), R = P.recipe_id[c0 + lo];
            const unsigned js = P.plan_jslot[(size_t)R * 256 + off];
            const int n = P.plan_cnt[(size_t)R * 256 + off], nsl = P.plan_nslots[R], pp = P.pair_ptr[c0 + lo];
            const unsigned *ent = P.plan_ent + P.plan_off[R] + (size_t)(js >> 8) * nsl + (js & 255u);
            double acc = 0.0;
            for (int it = 0; it < n; it++, ent += nsl) {
                const unsigned e = *ent;
                const int p = pp + (int)((e >> 16) & 0x7FFFu);
                acc = fma(P.a_val[P.a_tile_nnz[P.pair_a[p]] + (int)(e & 255u)], P.b_val[P.b_tile_nnz[P.pair_b[p]] + (int)((e >> 8) & 255u)], acc);
            }
            P.c_val[n0 + g] = acc;
            P.c_col[n0 + g] = P.plan_col[(size_t)R * 256 + off];
        }
        return;
    }
    size_t off = 0;
    auto carve = [&](size_t bytes) { unsigned char *p = pr_smem + off; off += (bytes + 15) & ~(size_t)15; return p; };
    double *s_aval = (double *)carve((size_t)nnzA * 8);
    int *s_cnnz = (int *)carve((size_t)(numJ + 1) * 4);
    int *s_slot0 = (int *)carve((size_t)(numJ + 1) * 4);
    int *s_rec = (int *)carve((size_t)numJ * 4);
    int *s_pp = (int *)carve((size_t)numJ * 4);
    int2 *s_base = (int2 *)carve((size_t)W * 8);
    for (int k = tid; k < nnzA; k += THREADS) s_aval[k] = P.a_val[av0 + k];
    for (int k = tid; k <= numJ; k += THREADS) {
        s_cnnz[k] = P.c_tile_nnz[c0 + k] - n0;
        if (k < numJ) {
            const int R = P.recipe_id[c0 + k];
            s_rec[k] = R;
            s_slot0[k] = P.plan_nslots[R];
            s_pp[k] = P.pair_ptr[c0 + k] - w0;
        }
    }
    for (int k = tid; k < W; k += THREADS) s_base[k] = make_int2(P.a_tile_nnz[P.pair_a[w0 + k]] - av0, P.b_tile_nnz[P.pair_b[w0 + k]]);
    __syncthreads();
    if (tid < 32) {  // first slot of every tile: exclusive scan of the slot counts over the tile-row's tiles
        int carry = 0;
        for (int k0 = 0; k0 < numJ; k0 += 32) {
            const int k = k0 + tid, v = k < numJ ? s_slot0[k] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL_MASK, incl, o);
                if (tid >= o) incl += u;
            }
            if (k < numJ) s_slot0[k] = carry + incl - v;
            carry += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (tid == 0) s_slot0[numJ] = carry;
    }
    __syncthreads();
    const int nslot_row = s_slot0[numJ];
    for (int q = tid; q < nslot_row; q += THREADS) {
        int lo = 0, hi = numJ - 1;  // the tile holding slot q of the tile-row: largest s with first slot <= q
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_slot0[mid] <= q) lo = mid; else hi = mid - 1;
        }
        const int sidx = lo, v = q - s_slot0[sidx], R = s_rec[sidx];
        const int nsl = s_slot0[sidx + 1] - s_slot0[sidx], pp = s_pp[sidx];
        const unsigned si = __ldg(P.plan_slot + R * 128 + v);
        const int n = (int)(si & 0xFFFFu), out = n0 + s_cnnz[sidx];  // nnz(C) of a slab fits int32 (spgemm_device checks)
        int ch = R * 256 + (int)(si >> 16), eo = P.plan_off[R] + v;   // next chain entry, next plan word
        unsigned oc = __ldg(P.plan_chain + ch);  // position of the chain's current nonzero in the tile | its column << 8
        double acc = 0.0;
        for (int it = 0; it < n; it += 4, eo += 4 * nsl) {  // four products at a time: all their loads first, then the sums in order
            unsigned e[4];
            double av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] = it + u < n ? __ldg(P.plan_ent + (eo + u * nsl)) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int2 base = s_base[pp + (int)((e[u] >> 16) & 0x7FFFu)];
                av[u] = s_aval[base.x + (int)(e[u] & 255u)];
                bv[u] = it + u < n ? __ldg(P.b_val + (base.y + (int)((e[u] >> 8) & 255u))) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (it + u < n) acc = fma(av[u], bv[u], acc);
                if (e[u] >> 31) {  // the nonzero's last product
                    const int o = out + (int)(oc & 255u);
                    P.c_val[o] = acc;
                    P.c_col[o] = (uint16_t)(oc >> 8);
                    acc = 0.0;
                    if (it + u + 1 < n) oc = __ldg(P.plan_chain + ++ch);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
// Before k_s1_fill<HASH>: clear the recipe table and hand its pointers to the fill kernel.
int plans_begin(PlanTable *out)
{
    Ctx &c = ctx();
    int rc = plan_ctx_init();
    if (rc) return rc;
    PlanCtx &p = g_plan;
    CK(cudaMemsetAsync(p.rkeys, 0, (size_t)RCAP * 8, c.stream));
    CK(cudaMemsetAsync(p.rowner, 0x7f, (size_t)RCAP * 4, c.stream));
    CK(cudaMemsetAsync(p.ctl + 2, 0, 2 * sizeof(int), c.stream));
    CK(cudaMemsetAsync(p.plan_tot, 0, ((size_t)RMAX + 1) * 4, c.stream));
    out->keys = p.rkeys; out->owner = p.rowner; out->count = p.ctl + 2; out->fail = p.ctl + 3;
    return TSG_OK;
}

// After k_s1_fill<HASH>: dense recipe numbers, verification, the plans, and C's masks / Ptr / tile nnz counts from them.
// Everything is enqueued; *d_fail is the device flag the caller reads back with nnz(C).
int plans_symbolic_device(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, const PairLists &pl, const unsigned *pair_pat, const int *rslot,
                          int *recipe_id, const RowTemplates *rt, const int **d_fail)
{
    Ctx &c = ctx();
    PlanCtx &p = g_plan;
    const int numblkC = C->numtile;
    int *fail = p.ctl + 3;
    *d_fail = fail;
    k_recipe_flags<<<ceil_div(RCAP, 256), 256, 0, c.stream>>>(p.rowner, p.rflags);
    CK_LAUNCH();
    int rc = exclusive_scan<int>(p.rflags, p.rdense, RCAP);
    if (rc) return rc;
    k_recipe_reps<<<ceil_div(RCAP, 256), 256, 0, c.stream>>>(p.rowner, p.rdense, p.rep_tile);
    CK_LAUNCH();
    k_recipe_verify<<<ceil_div(numblkC, 256), 256, 0, c.stream>>>(numblkC, pl.ptr, pl.end, pair_pat, rslot, p.rowner, p.rdense, recipe_id,
                                                                  fail, rt != nullptr);
    CK_LAUNCH();
    if (rt) {  // the other tile-rows: C's tile columns, pair lists and recipe ids from their representatives (rowplans.cu)
        rc = rowplans_instantiate(A, B, C, *rt, recipe_id);
        if (rc) return rc;
    }
    const int *nrec = p.rdense + RCAP;
    k_plan_build<false><<<ceil_div(RMAX * 16 * PB_SPLIT, 128), 128, 0, c.stream>>>(nrec, p.rep_tile, pl.ptr, pl.end, pl.a, pl.b, A->mask, A->ptr, B->mask,
                                                                        B->ptr, p.plan_mask, p.plan_ptr, p.plan_nnz, nullptr, p.plan_cnt,
                                                                        p.plan_col, nullptr, nullptr, nullptr, fail);
    CK_LAUNCH();
    const char *ch = getenv("TSG_PLANS_CHAIN");  // shortest slot length aimed at (the longest product list of the recipe if that is more)
    const int tmin = ch && *ch ? atoi(ch) : PLANS_CHAIN_MIN;
    k_plan_slots<<<ceil_div(RMAX, 4), 128, 0, c.stream>>>(nrec, tmin, p.plan_nnz, p.plan_cnt, p.plan_col, p.plan_jslot, p.plan_slot, p.plan_chain,
                                                          p.plan_nslots, p.plan_tot, fail);
    CK_LAUNCH();
    rc = exclusive_scan<int>(p.plan_tot, p.plan_off, RMAX);
    if (rc) return rc;
    k_plan_build<true><<<ceil_div(RMAX * 16 * PB_SPLIT, 128), 128, 0, c.stream>>>(nrec, p.rep_tile, pl.ptr, pl.end, pl.a, pl.b, A->mask, A->ptr, B->mask,
                                                                       B->ptr, p.plan_mask, p.plan_ptr, p.plan_nnz, p.plan_off, p.plan_cnt,
                                                                       p.plan_col, p.plan_jslot, p.plan_nslots, p.plan_ent, fail);
    CK_LAUNCH();
    k_symbolic_from_plans<<<ceil_div(numblkC, 256), 256, 0, c.stream>>>(numblkC, recipe_id, p.plan_mask, p.plan_ptr, p.plan_nnz, C->mask, C->ptr,
                                                                        C->tile_nnz, fail);
    CK_LAUNCH();
    return TSG_OK;
}

// ---------------------------------------------------------------------------------------------
// Row plans. With tile-row templates (rowplans.cu) the sequence of recipes along a tile-row is the template's, so the
// slots of the whole tile-row -- and the plan words they walk -- can be laid out once per template in the order the
// numeric kernel's warps take them: group g = slots 32g .. 32g+31, its words iteration-major (word of slot 32g+l at
// iteration it: goff[g] + 32 it + l). A warp then reads ONE 128-byte line per iteration instead of one piece per tile it
// spans (measured on config 2: 3.4 lines), and a slot's constants come from one 16-byte record instead of a binary
// search over the tile-row's tiles and four table lookups.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_rowplan_slots(int ntpl, const int *__restrict__ rep_list, const int *__restrict__ c_tile_ptr, const int *__restrict__ c_tile_nnz,
                const int *__restrict__ wptr, const int *__restrict__ pair_ptr, const int *__restrict__ recipe_id,
                const int *__restrict__ plan_nslots, const int *__restrict__ plan_off, const unsigned *__restrict__ plan_slot,
                int4 *__restrict__ rec, int2 *__restrict__ src, int *__restrict__ nslots, int *__restrict__ tau, int *fail)
{
    const int tpl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (tpl >= ntpl) return;  // whole warps
    const int i = rep_list[tpl];
    if (lane == 0) tau[i] = tpl;
    const int c0 = c_tile_ptr[i], numJ = c_tile_ptr[i + 1] - c0;
    const int n0 = c_tile_nnz[c0], w0 = wptr[i];
    int carry = 0;
    for (int k0 = 0; k0 < numJ; k0 += 32) {
        const int k = k0 + lane;
        const int R = k < numJ ? recipe_id[c0 + k] : -1;
        const int nsl = R >= 0 ? plan_nslots[R] : 0;
        int incl = nsl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        const int s0 = carry + incl - nsl;
        carry += __shfl_sync(FULL_MASK, incl, 31);
        if (R >= 0 && s0 + nsl <= RP_SCAP) {
            const int outoff = c_tile_nnz[c0 + k] - n0, pp = pair_ptr[c0 + k] - w0, po = plan_off[R];
            for (int v = 0; v < nsl; v++) {
                const unsigned si = plan_slot[R * 128 + v];
                rec[(size_t)tpl * RP_SCAP + s0 + v] = make_int4((int)(si & 0xFFFFu), R * 256 + (int)(si >> 16), outoff, pp);
                src[(size_t)tpl * RP_SCAP + s0 + v] = make_int2(po + v, nsl);
            }
        }
    }
    if (lane == 0) {
        nslots[tpl] = carry;
        if (carry > RP_SCAP) *fail = 1;
    }
}

__global__ void __launch_bounds__(128)
k_rowplan_groups(int ntpl, const int4 *__restrict__ rec, const int *__restrict__ nslots, int *__restrict__ goff, int *fail)
{
    const int tpl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (tpl >= ntpl) return;
    const int S = nslots[tpl];
    if (S > RP_SCAP) return;
    const int G = (S + 31) >> 5;
    int carry = 0;
    for (int g0 = 0; g0 < G; g0 += 32) {
        const int g = g0 + lane;
        int len = 0;
        if (g < G)
            for (int j = 0; j < 32; j++) {
                const int q = g * 32 + j;
                if (q < S) len = max(len, rec[(size_t)tpl * RP_SCAP + q].x);
            }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (g < G) goff[tpl * (RP_GCAP + 1) + g] = 32 * (carry + incl - len);
        carry += __shfl_sync(FULL_MASK, incl, 31);
    }
    if (lane == 0) {
        goff[tpl * (RP_GCAP + 1) + G] = 32 * carry;
        if (32 * carry > RP_WCAP) *fail = 1;
    }
}

__global__ void __launch_bounds__(256)
k_rowplan_words(const int4 *__restrict__ rec, const int2 *__restrict__ src, const int *__restrict__ nslots, const int *__restrict__ goff,
                const unsigned *__restrict__ plan_ent, unsigned *__restrict__ words, const int *__restrict__ fail)
{
    if (*fail) return;
    const int tpl = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
    const int S = nslots[tpl], G = (S + 31) >> 5, g = q >> 5;
    if (g >= G) return;
    const int o0 = goff[tpl * (RP_GCAP + 1) + g], glen = (goff[tpl * (RP_GCAP + 1) + g + 1] - o0) >> 5;
    int n = 0, base = 0, stride = 0;
    if (q < S) {
        n = rec[(size_t)tpl * RP_SCAP + q].x;
        const int2 sr = src[(size_t)tpl * RP_SCAP + q];
        base = sr.x; stride = sr.y;
    }
    unsigned *dst = words + (size_t)tpl * RP_WCAP + o0 + (q & 31);
    for (int it = 0; it < glen; it++) dst[it * 32] = it < n ? plan_ent[base + it * stride] : 0u;
}

struct RowPlanRows {
    int trow0;
    const int *a_tile_ptr, *a_tile_nnz;
    const double *a_val;
    const int *b_tile_nnz;
    const double *b_val;
    const int *c_tile_ptr, *c_tile_nnz, *wptr, *pair_a, *pair_b, *rep_of, *tau;
    const int4 *rec;
    const unsigned *words;
    const int *goff, *nslots;
    const uint16_t *plan_chain;
    uint16_t *c_col;
    double *c_val;
};

__host__ __device__ __forceinline__ size_t rowplan_rows_need(int nnzA, int W)
{
    return (((size_t)nnzA * 8 + 15) & ~(size_t)15) + (((size_t)W * 8 + 15) & ~(size_t)15);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_numeric_from_rowplans(const __grid_constant__ RowPlanRows P)
{
    extern __shared__ __align__(16) unsigned char rr_smem[];
    const int i = blockIdx.x, tid = threadIdx.x, I = P.trow0 + i;
    const int rep = P.rep_of[i];
    if (rep < 0) return;  // a tile-row without pairs
    const int tpl = P.tau[rep];
    const int c0 = P.c_tile_ptr[i], numJ = P.c_tile_ptr[i + 1] - c0;
    if (numJ == 0) return;
    const int n0 = P.c_tile_nnz[c0];
    if (P.c_tile_nnz[c0 + numJ] == n0) return;
    const int a0 = P.a_tile_ptr[I], a1 = P.a_tile_ptr[I + 1];
    const int av0 = P.a_tile_nnz[a0], nnzA = P.a_tile_nnz[a1] - av0;
    const int w0 = P.wptr[i], W = P.wptr[i + 1] - w0;
    double *s_aval = (double *)rr_smem;
    int2 *s_base = (int2 *)(rr_smem + (((size_t)nnzA * 8 + 15) & ~(size_t)15));
    for (int k = tid; k < nnzA; k += THREADS) s_aval[k] = P.a_val[av0 + k];
    for (int k = tid; k < W; k += THREADS) s_base[k] = make_int2(P.a_tile_nnz[P.pair_a[w0 + k]] - av0, P.b_tile_nnz[P.pair_b[w0 + k]]);
    __syncthreads();
    const int S = P.nslots[tpl];
    const int4 *rec = P.rec + (size_t)tpl * RP_SCAP;
    const unsigned *words = P.words + (size_t)tpl * RP_WCAP;
    const int *goff = P.goff + tpl * (RP_GCAP + 1);
    for (int q = tid; q < S; q += THREADS) {
        const int4 rc = __ldg(rec + q);
        const int n = rc.x, out = n0 + rc.z, pp = rc.w;
        int ch = rc.y;
        const unsigned *wp = words + __ldg(goff + (q >> 5)) + (q & 31);
        unsigned oc = __ldg(P.plan_chain + ch);  // position of the chain's current nonzero in the tile | its column << 8
        double acc = 0.0;
        for (int it = 0; it < n; it += 4, wp += 128) {  // four products at a time: all their loads first, then the sums in order
            unsigned e[4];
            double av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] = it + u < n ? __ldg(wp + u * 32) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int2 base = s_base[pp + (int)((e[u] >> 16) & 0x7FFFu)];
                av[u] = s_aval[base.x + (int)(e[u] & 255u)];
                bv[u] = it + u < n ? __ldg(P.b_val + (base.y + (int)((e[u] >> 8) & 255u))) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (it + u < n) acc = fma(av[u], bv[u], acc);
                if (e[u] >> 31) {  // the nonzero's last product
                    const int o = out + (int)(oc & 255u);
                    P.c_val[o] = acc;
                    P.c_col[o] = (uint16_t)(oc >> 8);
                    acc = 0.0;
                    if (it + u + 1 < n) oc = __ldg(P.plan_chain + ++ch);
                }
            }
        }
    }
}

static bool rowplan_numeric_env_on()
{
    const char *e = getenv("TSG_ROWPLAN_NUMERIC");
    return !(e && *e == '0');
}

// Builds the row plans of the slab's templates (after the scan of C's tile nnz); *d_fail = the device flag that says
// "caps exceeded: walk the per-recipe plans". Returns with *d_fail = nullptr when row plans are not attempted.
int plans_rowplan_build(const tsg_dtile *C, const RowTemplates &rt, const int *recipe_id, const int *pair_ptr, int **d_tau, const int **d_fail)
{
    Ctx &c = ctx();
    PlanCtx &p = g_plan;
    *d_fail = nullptr;
    *d_tau = nullptr;
    if (!rowplan_numeric_env_on() || rt.n <= 0 || rt.n > RP_TMAX) return TSG_OK;
    if (!p.rp_rec) {
        p.rp_rec = dalloc_n<int4>((size_t)RP_TMAX * RP_SCAP);
        p.rp_src = dalloc_n<int2>((size_t)RP_TMAX * RP_SCAP);
        p.rp_words = dalloc_n<unsigned>((size_t)RP_TMAX * RP_WCAP);
        p.rp_goff = dalloc_n<int>((size_t)RP_TMAX * (RP_GCAP + 1));
        p.rp_nslots = dalloc_n<int>(RP_TMAX);
        if (!p.rp_rec || !p.rp_src || !p.rp_words || !p.rp_goff || !p.rp_nslots) { p.rp_rec = nullptr; return last_error(); }
    }
    int *tau = dalloc_n<int>((size_t)rt.ntr + 1);
    if (!tau) return last_error();
    int *fail = p.ctl + 4;
    CK(cudaMemsetAsync(fail, 0, sizeof(int), c.stream));
    k_rowplan_slots<<<ceil_div(rt.n * 32, 128), 128, 0, c.stream>>>(rt.n, rt.rep_list, C->tile_ptr, C->tile_nnz, rt.wptr, pair_ptr, recipe_id,
                                                                    p.plan_nslots, p.plan_off, p.plan_slot, p.rp_rec, p.rp_src, p.rp_nslots, tau, fail);
    CK_LAUNCH();
    k_rowplan_groups<<<ceil_div(rt.n * 32, 128), 128, 0, c.stream>>>(rt.n, p.rp_rec, p.rp_nslots, p.rp_goff, fail);
    CK_LAUNCH();
    k_rowplan_words<<<dim3(RP_SCAP / 256, rt.n), 256, 0, c.stream>>>(p.rp_rec, p.rp_src, p.rp_nslots, p.rp_goff, p.plan_ent, p.rp_words, fail);
    CK_LAUNCH();
    *d_fail = fail;
    *d_tau = tau;
    return TSG_OK;
}

int plans_numeric_device(const tsg_dtile *A, const tsg_dtile *B, tsg_dtile *C, const PairLists &pl, const int *recipe_id, int trow0, int ntr,
                         const int *wptr, int max_need, tsg_stats *stats, const RowTemplates *rt, const int *tau, int max_nnzA_row, int wmax)
{
    Ctx &c = ctx();
    PlanCtx &p = g_plan;
    if (C->nnz <= 0 || C->numtile <= 0) return TSG_OK;
    if (rt && tau) {  // row plans: every tile-row walks its template's slots, laid out warp by warp (see k_rowplan_slots)
        size_t smem = rowplan_rows_need(max_nnzA_row, wmax);
        smem = (smem + 1023) & ~(size_t)1023;
        if (smem <= (size_t)64 * 1024) {
            RowPlanRows P{trow0, A->tile_ptr, A->tile_nnz, A->val, B->tile_nnz, B->val, C->tile_ptr, C->tile_nnz, wptr, pl.a, pl.b, rt->rep_of, tau,
                          p.rp_rec, p.rp_words, p.rp_goff, p.rp_nslots, p.plan_chain, C->col, C->val};
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_numeric_from_rowplans<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const char *cv = getenv("TSG_PLANS_CARVEOUT");
            int pct = cv && *cv ? atoi(cv) : (int)((6 * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            if (pct >= 0) CK(cudaFuncSetAttribute(k_numeric_from_rowplans<256, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            k_numeric_from_rowplans<256, 6><<<ntr, 256, smem, c.stream>>>(P);
            CK_LAUNCH();
            if (stats) stats->plan_recipes = 1;
            return TSG_OK;
        }
    }
    const char *kb = getenv("TSG_PLANS_SMEM_KB");  // tests: a small budget sends tile-rows down the unstaged branch
    const size_t cap = kb && *kb ? (size_t)atoi(kb) * 1024 : (size_t)64 * 1024;
    size_t smem = (size_t)max_need < cap ? (size_t)max_need : cap;
    smem = (smem + 1023) & ~(size_t)1023;
    PlanRows P{trow0, (int)smem, A->tile_ptr, A->tile_nnz, A->val, B->tile_nnz, B->val, C->tile_ptr, C->tile_nnz, wptr, pl.ptr, pl.a, pl.b,
               recipe_id, p.plan_off, p.plan_nslots, p.plan_cnt, p.plan_chain, p.plan_col, p.plan_slot, p.plan_jslot, p.plan_ent, C->col, C->val};
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_numeric_from_plans_rows<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {   // shared memory the 6 resident CTAs need, as a share of the SM's 228 KB: the rest stays L1 for the gathers of B's values
        const char *cv = getenv("TSG_PLANS_CARVEOUT");
        int pct = cv && *cv ? atoi(cv) : (int)((6 * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        if (pct >= 0) CK(cudaFuncSetAttribute(k_numeric_from_plans_rows<256, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    k_numeric_from_plans_rows<256, 6><<<ntr, 256, smem, c.stream>>>(P);
    CK_LAUNCH();
    if (stats) stats->plan_recipes = 1;  // the caller fills in the count it read back
    return TSG_OK;
}

// largest shared-memory need of k_numeric_from_plans_rows over the slab's tile-rows (host side, from sizes known after step 1)
size_t plans_rows_need_bound(int max_nnzA_row, int maxJ, int wmax) { return plan_rows_need(max_nnzA_row, maxJ, wmax); }

const int *plans_recipe_count_ptr() { return g_plan.rdense ? g_plan.rdense + RCAP : nullptr; }

}  // namespace tsg
