// plans.cuh -- DRAFT for round 2: recipe-cached symbolic + numeric steps.
// Not compiled into libtilespgemm_b200.so and not yet run on a GPU. The index logic is checked on the host by
// scratch/next/test_plans_emul.py (serial emulation through emul.h, bit-exact against the oracle); `nvcc -c` of
// plans_compile_check.cu checks that it is valid CUDA for sm_100a. The deduplication never makes a thread wait for another
// (insert = one atomicCAS on the key word, owner = atomicMin, full-key verification in a second kernel), so what the
// emulation leaves unchecked is performance, not protocol.
//
// Pipeline (after step 1 has produced the pair lists, with its fused symbolic switched off):
//   k_pattern_insert / _verify   tile -> pattern id: the 32-byte block of 16 row masks, deduplicated through a device hash table
//   k_recipe_insert / _flags / _reps / _verify   C tile -> recipe id: the sequence of (A pattern, B pattern) over its pairs
//   k_plan_build    one thread per distinct recipe, from a representative C tile: the tile's masks / Ptr / nnz and, per C
//                   nonzero in storage order, the (pair, position in A's tile, position in B's tile) sources in the serial
//                   SPA's order (ascending pair, then ascending k)   [count pass, scan, fill pass]
//   k_symbolic_from_plans   C tile masks / Ptr / nnz = a 68-byte copy from the plan
//   k_numeric_from_plans    one lane per C nonzero walks its plan entries: every iteration is a product
// Any overflow (too many patterns / recipes, > 65535 pairs in a C tile) raises *fail and the caller falls back to the
// generic kernels.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#include "emul.h"  // scratch/next: serial host emulation of these kernels
#endif

#ifndef TS
#define TS 16
#endif

namespace plans {

constexpr int PCAP = 1 << 13;      // pattern table slots; more than PCAP/2 distinct patterns => fail
constexpr int RCAP = 1 << 15;      // recipe table slots;  more than RCAP/2 distinct recipes  => fail
constexpr int PLAN_ROWS = 257;     // plan_start entries per recipe (<= 256 nonzeros per C tile, plus the end)

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v)
{
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}

// ---- deduplication without waiting: (1) insert the 64-bit hash with one atomicCAS on the key word, (2) the smallest item
// index that landed in a slot becomes its owner (atomicMin: deterministic), (3) every item compares its full key with
// the owner's; a true 64-bit collision raises *fail (the generic kernels then run instead). No thread ever waits for another.
constexpr int NO_OWNER = 0x7f7f7f7f;  // what cudaMemset(0x7f) leaves; item indices stay below it

__device__ __forceinline__ int table_insert(unsigned long long *keys, int cap, unsigned long long h, int *count, int limit, int *fail)
{
    if (h == 0ull) h = 1ull;  // 0 marks an empty slot
    unsigned slot = (unsigned)(h >> 20) & (unsigned)(cap - 1);
    for (int probe = 0; probe < cap; probe++, slot = (slot + 1) & (unsigned)(cap - 1)) {
        if (*(volatile int *)fail) return -1;
        unsigned long long old = *(volatile unsigned long long *)&keys[slot];  // millions of items, a few dozen hot slots: read first
        if (old == 0ull) old = atomicCAS(&keys[slot], 0ull, h);
        if (old == 0ull) {
            if (atomicAdd(count, 1) >= limit) *fail = 1;
            return (int)slot;
        }
        if (old == h) return (int)slot;
    }
    *fail = 1;
    return -1;
}

// Tiles of A (index t) and of B (index numtileA + t) share one table: `first` is 0 for A's launch, numtileA for B's.
__global__ void __launch_bounds__(256)
k_pattern_insert(int numtile, int first, const uint16_t *__restrict__ mask, unsigned long long *keys, int *owner, int *count,
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
    if (slot >= 0 && first + t < *(volatile int *)&owner[slot]) atomicMin(&owner[slot], first + t);  // owner only ever decreases
}

__global__ void __launch_bounds__(256)
k_pattern_verify(int numtile, const uint16_t *__restrict__ mask, const int *__restrict__ pat_id, const int *__restrict__ owner,
                 int numtileA, const uint16_t *__restrict__ a_mask, const uint16_t *__restrict__ b_mask, int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numtile) return;
    const int slot = pat_id[t];
    if (slot < 0) { *fail = 1; return; }
    const int o = owner[slot];
    const uint4 *op = reinterpret_cast<const uint4 *>(o < numtileA ? a_mask + (size_t)o * TS : b_mask + (size_t)(o - numtileA) * TS);
    const uint4 *mp = reinterpret_cast<const uint4 *>(mask + (size_t)t * TS);
    const uint4 x = mp[0], y = mp[1], a = op[0], b = op[1];
    if (!(a.x == x.x && a.y == x.y && a.z == x.z && a.w == x.w && b.x == y.x && b.y == y.y && b.z == y.z && b.w == y.w)) *fail = 2;
}

// Recipe of a C tile = the (A pattern, B pattern) sequence of its pairs.
__global__ void __launch_bounds__(256)
k_recipe_insert(int numblkC, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end, const int *__restrict__ pair_a,
                const int *__restrict__ pair_b, const int *__restrict__ patA, const int *__restrict__ patB,
                unsigned long long *keys, int *owner, int *count, int *__restrict__ rslot, int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numblkC) return;
    rslot[t] = -1;
    const int p0 = pair_ptr[t], p1 = pair_end[t];
    if (p1 - p0 > 0xFFFF) { *fail = 1; return; }  // the plan entry keeps the pair index in 16 bits
    unsigned long long h = mix64(0x13198A2E03707344ull, (unsigned long long)(p1 - p0));
    for (int p = p0; p < p1; p++) h = mix64(h, ((unsigned long long)(unsigned)patA[pair_a[p]] << 32) | (unsigned)patB[pair_b[p]]);
    const int slot = table_insert(keys, RCAP, h, count, RCAP / 2, fail);
    rslot[t] = slot;
    if (slot >= 0 && t < *(volatile int *)&owner[slot]) atomicMin(&owner[slot], t);
}

// flags[slot] = 1 where the slot has an owner; the caller scans them into dense recipe numbers (rdense).
__global__ void __launch_bounds__(256)
k_recipe_flags(const int *__restrict__ owner, int *__restrict__ flags)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < RCAP) flags[s] = owner[s] != NO_OWNER;
}

// rep_tile holds RCAP/2 entries: inserts that raced past the limit before they saw *fail must not write beyond it
__global__ void __launch_bounds__(256)
k_recipe_reps(const int *__restrict__ owner, const int *__restrict__ rdense, int *__restrict__ rep_tile)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < RCAP && owner[s] != NO_OWNER && rdense[s] < RCAP / 2) rep_tile[rdense[s]] = owner[s];
}

__global__ void __launch_bounds__(256)
k_recipe_verify(int numblkC, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end, const int *__restrict__ pair_a,
                const int *__restrict__ pair_b, const int *__restrict__ patA, const int *__restrict__ patB,
                const int *__restrict__ rslot, const int *__restrict__ owner, const int *__restrict__ rdense,
                int *__restrict__ recipe_id, int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numblkC) return;
    const int slot = rslot[t];
    if (slot < 0) { *fail = 1; return; }
    const int u = owner[slot];
    const int p0 = pair_ptr[t], n = pair_end[t] - p0, q0 = pair_ptr[u];
    bool same = pair_end[u] - q0 == n;
    for (int i = 0; i < n && same; i++)
        same = patA[pair_a[p0 + i]] == patA[pair_a[q0 + i]] && patB[pair_b[p0 + i]] == patB[pair_b[q0 + i]];
    if (!same) *fail = 2;
    recipe_id[t] = rdense[slot];
}

// One thread per distinct recipe. FILL = false: masks / Ptr / nnz of the recipe's C tile and the number of plan entries
// (plan_tot). FILL = true: plan_off = exclusive scan of plan_tot; writes plan_start, plan_col and the entries
// (pair index << 16 | position in B's tile << 8 | position in A's tile).
template <bool FILL>
__global__ void __launch_bounds__(64)
k_plan_build(int nrec, const int *__restrict__ rep_tile, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end,
             const int *__restrict__ pair_a, const int *__restrict__ pair_b, const uint16_t *__restrict__ a_mask,
             const uint16_t *__restrict__ a_ptr, const uint16_t *__restrict__ b_mask, const uint16_t *__restrict__ b_ptr,
             uint16_t *plan_mask, uint16_t *plan_ptr, int *plan_nnz, int *plan_tot, const int *__restrict__ plan_off,
             unsigned *plan_start, uint8_t *plan_col, unsigned *plan_ent)
{
    const int R = blockIdx.x * blockDim.x + threadIdx.x;
    if (R >= nrec) return;
    const int t = rep_tile[R];
    const int p0 = pair_ptr[t], p1 = pair_end[t];
    unsigned cm[TS];
    for (int r = 0; r < TS; r++) cm[r] = 0;
    for (int p = p0; p < p1; p++) {
        const int a = pair_a[p], b = pair_b[p];
        for (int r = 0; r < TS; r++) {
            unsigned am = a_mask[(size_t)a * TS + r];
            while (am) {
                const int k = __clz(am) - 16;
                am ^= 0x8000u >> k;
                cm[r] |= b_mask[(size_t)b * TS + k];
            }
        }
    }
    if (!FILL) {
        int run = 0;
        for (int r = 0; r < TS; r++) {
            plan_ptr[R * TS + r] = (uint16_t)run;
            plan_mask[R * TS + r] = (uint16_t)cm[r];
            run += __popc(cm[r]);
        }
        plan_nnz[R] = run;
    }
    unsigned e = FILL ? (unsigned)plan_off[R] : 0u;
    int j = 0;
    for (int r = 0; r < TS; r++) {
        unsigned rowm = cm[r];
        while (rowm) {
            const int c = __clz(rowm) - 16;
            rowm ^= 0x8000u >> c;
            const unsigned cbit = 0x8000u >> c;
            if (FILL) { plan_start[(size_t)R * PLAN_ROWS + j] = e; plan_col[(size_t)R * 256 + j] = (uint8_t)c; }
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
                            plan_ent[e] = ((unsigned)(p - p0) << 16) | (posb << 8) | ia;
                        }
                        e++;
                    }
                    ia++;
                }
            }
            j++;
        }
    }
    if (FILL) plan_start[(size_t)R * PLAN_ROWS + j] = e;
    else plan_tot[R] = (int)e;
}

// C tile metadata from the plan: thread per (tile, row).
__global__ void __launch_bounds__(256)
k_symbolic_from_plans(int numblkC, const int *__restrict__ recipe_id, const uint16_t *__restrict__ plan_mask,
                      const uint16_t *__restrict__ plan_ptr, const int *__restrict__ plan_nnz, uint16_t *__restrict__ c_mask,
                      uint16_t *__restrict__ c_ptr, int *__restrict__ c_cnt)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = (int)(gid >> 4), r = (int)(gid & 15);
    if (t >= numblkC) return;
    const int R = recipe_id[t];
    c_mask[(size_t)t * TS + r] = plan_mask[R * TS + r];
    c_ptr[(size_t)t * TS + r] = plan_ptr[R * TS + r];
    if (r == 0) c_cnt[t] = plan_nnz[R];
}

// One lane per C nonzero g. blk2tile as in the gather kernel of the product (tile holding nonzero 32*(g/32)).
__global__ void __launch_bounds__(256)
k_numeric_from_plans(int numblkC, int nnzC, const int *__restrict__ blk2tile, const int *__restrict__ c_tile_nnz,
                     const int *__restrict__ recipe_id, const unsigned *__restrict__ plan_start,
                     const uint8_t *__restrict__ plan_col, const unsigned *__restrict__ plan_ent,
                     const int *__restrict__ pair_ptr, const int *__restrict__ pair_a, const int *__restrict__ pair_b,
                     const int *__restrict__ a_tile_nnz, const double *__restrict__ a_val, const int *__restrict__ b_tile_nnz,
                     const double *__restrict__ b_val, uint16_t *__restrict__ c_col, double *__restrict__ c_val)
{
    const long long gl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= nnzC) return;
    const int g = (int)gl;
    const int blk = g >> 5, nblk = (nnzC + 31) >> 5;
    int lo = blk2tile[blk], hi = blk + 1 < nblk ? blk2tile[blk + 1] : numblkC - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (c_tile_nnz[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const int t = lo, off = g - c_tile_nnz[t];
    const int R = recipe_id[t];
    const unsigned s1 = plan_start[(size_t)R * PLAN_ROWS + off + 1];
    const int pp = pair_ptr[t];
    double acc = 0.0;
    for (unsigned s = plan_start[(size_t)R * PLAN_ROWS + off]; s < s1; s++) {
        const unsigned e = plan_ent[s];
        const int a = pair_a[pp + (int)(e >> 16)], b = pair_b[pp + (int)(e >> 16)];
        acc = fma(a_val[a_tile_nnz[a] + (int)(e & 255u)], b_val[b_tile_nnz[b] + (int)((e >> 8) & 255u)], acc);
    }
    c_val[g] = acc;
    c_col[g] = plan_col[(size_t)R * 256 + off];
}

// Variant B of the numeric step: the value bases of a pair's two tiles are looked up once per pair by a streaming kernel
// (k_pair_bases) instead of twice per product through pair_a -> a_tile_nnz / pair_b -> b_tile_nnz.
struct int2_ { int x, y; };

__global__ void __launch_bounds__(256)
k_pair_bases(long long npairs, const int *__restrict__ pair_a, const int *__restrict__ pair_b, const int *__restrict__ a_tile_nnz,
             const int *__restrict__ b_tile_nnz, int2_ *__restrict__ pair_base)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    int2_ v;
    v.x = a_tile_nnz[pair_a[p]];
    v.y = b_tile_nnz[pair_b[p]];
    pair_base[p] = v;
}

__global__ void __launch_bounds__(256)
k_numeric_from_plans_b(int numblkC, int nnzC, const int *__restrict__ blk2tile, const int *__restrict__ c_tile_nnz,
                       const int *__restrict__ recipe_id, const unsigned *__restrict__ plan_start,
                       const uint8_t *__restrict__ plan_col, const unsigned *__restrict__ plan_ent,
                       const int *__restrict__ pair_ptr, const int2_ *__restrict__ pair_base, const double *__restrict__ a_val,
                       const double *__restrict__ b_val, uint16_t *__restrict__ c_col, double *__restrict__ c_val)
{
    const long long gl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= nnzC) return;
    const int g = (int)gl;
    const int blk = g >> 5, nblk = (nnzC + 31) >> 5;
    int lo = blk2tile[blk], hi = blk + 1 < nblk ? blk2tile[blk + 1] : numblkC - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (c_tile_nnz[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const int t = lo, off = g - c_tile_nnz[t];
    const int R = recipe_id[t];
    const unsigned s1 = plan_start[(size_t)R * PLAN_ROWS + off + 1];
    const int2_ *pb = pair_base + pair_ptr[t];
    double acc = 0.0;
    for (unsigned s = plan_start[(size_t)R * PLAN_ROWS + off]; s < s1; s++) {
        const unsigned e = plan_ent[s];
        const int2_ base = pb[e >> 16];
        acc = fma(a_val[base.x + (int)(e & 255u)], b_val[base.y + (int)((e >> 8) & 255u)], acc);
    }
    c_val[g] = acc;
    c_col[g] = plan_col[(size_t)R * 256 + off];
}

}  // namespace plans
