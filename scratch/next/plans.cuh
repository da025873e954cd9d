// plans.cuh -- DRAFT for round 2: recipe-cached symbolic + numeric steps.
// Not compiled into libtilespgemm_b200.so and not yet run on a GPU. The index logic is checked on the host by
// scratch/next/test_plans_emul.py (serial emulation through emul.h, bit-exact against the oracle); `nvcc -c` of
// plans_compile_check.cu checks that it is valid CUDA for sm_100a. What the emulation cannot check is the concurrent
// behaviour of the two hash-table insert protocols (k_pattern_ids, k_recipe_ids).
//
// Pipeline (after step 1 has produced the pair lists, with its fused symbolic switched off):
//   k_pattern_ids   tile -> pattern id: the 32-byte block of 16 row masks, deduplicated in a device hash table (full compare)
//   k_recipe_ids    C tile -> recipe id: the sequence of (A pattern, B pattern) over its pairs, deduplicated likewise
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
#include "emul.h"
#define LD_CG(p) (*(p))
#else
#define LD_CG(p) __ldcg(p)
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

// state[slot]: 0 empty, 1 being written, 2 ready. blk: two uint4 per slot (the 16 row masks).
__global__ void __launch_bounds__(256)
k_pattern_ids(int numtile, const uint16_t *__restrict__ mask, unsigned *state, uint4 *blk, int *count, int *__restrict__ pat_id,
              int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numtile) return;
    pat_id[t] = -1;
    const uint4 *mp = reinterpret_cast<const uint4 *>(mask + (size_t)t * TS);
    const uint4 x = mp[0], y = mp[1];
    unsigned long long h = 0x243F6A8885A308D3ull;
    h = mix64(h, ((unsigned long long)x.x << 32) | x.y); h = mix64(h, ((unsigned long long)x.z << 32) | x.w);
    h = mix64(h, ((unsigned long long)y.x << 32) | y.y); h = mix64(h, ((unsigned long long)y.z << 32) | y.w);
    unsigned slot = (unsigned)h & (PCAP - 1);
    for (int probe = 0; probe < PCAP; probe++, slot = (slot + 1) & (PCAP - 1)) {
        if (*(volatile int *)fail) return;
        unsigned st = atomicCAS(&state[slot], 0u, 1u);
        if (st == 0u) {
            blk[2 * slot] = x; blk[2 * slot + 1] = y;
            __threadfence();
            atomicExch(&state[slot], 2u);
            if (atomicAdd(count, 1) >= PCAP / 2) *fail = 1;
            pat_id[t] = (int)slot;
            return;
        }
        while (st == 1u) st = *(volatile unsigned *)&state[slot];
        __threadfence();
        const uint4 a = LD_CG(&blk[2 * slot]), b = LD_CG(&blk[2 * slot + 1]);
        if (a.x == x.x && a.y == x.y && a.z == x.z && a.w == x.w && b.x == y.x && b.y == y.y && b.z == y.z && b.w == y.w) {
            pat_id[t] = (int)slot;
            return;
        }
    }
    *fail = 1;
}

// rhash / rrep / rdense per slot: hash of the sequence, representative C tile, dense recipe number.
__global__ void __launch_bounds__(256)
k_recipe_ids(int numblkC, const int *__restrict__ pair_ptr, const int *__restrict__ pair_end, const int *__restrict__ pair_a,
             const int *__restrict__ pair_b, const int *__restrict__ patA, const int *__restrict__ patB, unsigned *state,
             unsigned long long *rhash, int *rrep, int *rdense, int *count, int *__restrict__ recipe_id, int *rep_tile, int *fail)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numblkC) return;
    recipe_id[t] = -1;
    const int p0 = pair_ptr[t], p1 = pair_end[t];
    if (p1 - p0 > 0xFFFF) { *fail = 1; return; }  // the plan entry keeps the pair index in 16 bits
    unsigned long long h = mix64(0x13198A2E03707344ull, (unsigned long long)(p1 - p0));
    for (int p = p0; p < p1; p++) h = mix64(h, ((unsigned long long)(unsigned)patA[pair_a[p]] << 32) | (unsigned)patB[pair_b[p]]);
    unsigned slot = (unsigned)h & (RCAP - 1);
    for (int probe = 0; probe < RCAP; probe++, slot = (slot + 1) & (RCAP - 1)) {
        if (*(volatile int *)fail) return;
        unsigned st = atomicCAS(&state[slot], 0u, 1u);
        if (st == 0u) {
            const int d = atomicAdd(count, 1);
            rhash[slot] = h; rrep[slot] = t; rdense[slot] = d;
            if (d >= RCAP / 2) *fail = 1; else rep_tile[d] = t;
            __threadfence();
            atomicExch(&state[slot], 2u);
            recipe_id[t] = d;
            return;
        }
        while (st == 1u) st = *(volatile unsigned *)&state[slot];
        __threadfence();
        if (LD_CG(&rhash[slot]) != h) continue;
        const int u = LD_CG(&rrep[slot]);
        const int q0 = pair_ptr[u];
        if (pair_end[u] - q0 != p1 - p0) continue;
        bool same = true;
        for (int i = 0; i < p1 - p0 && same; i++)
            same = patA[pair_a[p0 + i]] == patA[pair_a[q0 + i]] && patB[pair_b[p0 + i]] == patB[pair_b[q0 + i]];
        if (same) { recipe_id[t] = LD_CG(&rdense[slot]); return; }
    }
    *fail = 1;
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

}  // namespace plans
