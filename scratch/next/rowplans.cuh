// rowplans.cuh -- DRAFT for round 2/3: step 1 (tile-level symbolic) from TILE-ROW recipes.
// Not compiled into the library, never run on a GPU; checked by serial host emulation (test_rowplans_emul.py).
//
// On structured matrices the tile-level structure repeats from tile-row to tile-row up to a translation: the tile columns of
// A's tile-row I relative to I, and for each of them the tile columns of B's tile-row K relative to K. Two tile-rows with the
// same relative structure have the same C tile columns relative to I and the same pair lists in local coordinates
// (i = index of the A tile inside the tile-row, j = index of the B tile inside B's tile-row K_i). So:
//   k_brow_insert/_verify   B tile-row K -> id of its relative column sequence (J - K)           [per B, cacheable]
//   k_arow_insert/_verify   A tile-row I -> row recipe = sequence of (K - I, id of B tile-row K)
//   k_rowplan_build         one thread per distinct row recipe: sorted distinct D = J - I (C tile columns relative to I),
//                           pairs per C tile, and the pair list as (i, j) in C-tile-major, ascending-i order
//   k_expand_tiles/_pairs   every tile-row: C tile columns = I + D, pair ranges, pair_a = tile_ptrA[I] + i,
//                           pair_b = rm2csc[tile_ptrB[K_i] + j]  -- streaming writes, no bitmap, no atomics
// Limits (else *fail and the generic step 1 runs): <= RROW_CAP/2 distinct recipes of either kind, <= MAXW pairs and
// <= 65535 tiles per tile-row.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#include "emul.h"
#endif
#include "plans.cuh"  // mix64, table_insert, NO_OWNER

namespace rowplans {

using plans::mix64;
using plans::table_insert;
using plans::NO_OWNER;

constexpr int RROW_CAP = 1 << 14;  // slots of the B-row and A-row tables
constexpr int MAXW = 4096;         // pairs per tile-row a row plan may hold

__global__ void __launch_bounds__(256)
k_brow_insert(int tilemB, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, unsigned long long *keys, int *owner,
              int *count, int *__restrict__ brow_id, int *fail)
{
    const int K = blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= tilemB) return;
    const int s = b_tile_ptr[K], e = b_tile_ptr[K + 1];
    if (e - s > 0xFFFF) { *fail = 1; brow_id[K] = -1; return; }
    unsigned long long h = mix64(0x452821E638D01377ull, (unsigned long long)(e - s));
    for (int q = s; q < e; q++) h = mix64(h, (unsigned long long)(long long)(b_tile_col[q] - K));
    const int slot = table_insert(keys, RROW_CAP, h, count, RROW_CAP / 2, fail);
    brow_id[K] = slot;
    if (slot >= 0 && K < *(volatile int *)&owner[slot]) atomicMin(&owner[slot], K);
}

__global__ void __launch_bounds__(256)
k_brow_verify(int tilemB, const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, const int *__restrict__ brow_id,
              const int *__restrict__ owner, int *fail)
{
    const int K = blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= tilemB) return;
    const int slot = brow_id[K];
    if (slot < 0) { *fail = 1; return; }
    const int O = owner[slot];
    const int s = b_tile_ptr[K], n = b_tile_ptr[K + 1] - s, so = b_tile_ptr[O];
    bool same = b_tile_ptr[O + 1] - so == n;
    for (int q = 0; q < n && same; q++) same = b_tile_col[s + q] - K == b_tile_col[so + q] - O;
    if (!same) *fail = 2;
}

// A tile-rows of the slab [trow0, trow0 + ntr). w[I - trow0] = pairs of the tile-row (the step-1 weight).
__global__ void __launch_bounds__(256)
k_arow_insert(int ntr, int trow0, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col, const int *__restrict__ b_tile_ptr,
              const int *__restrict__ brow_id, unsigned long long *keys, int *owner, int *count, int *__restrict__ arow_slot,
              int *__restrict__ w, int *fail)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= ntr) return;
    const int I = trow0 + x;
    const int s = a_tile_ptr[I], e = a_tile_ptr[I + 1];
    arow_slot[x] = -1;
    long long pairs = 0;
    unsigned long long h = mix64(0xBE5466CF34E90C6Cull, (unsigned long long)(e - s));
    for (int q = s; q < e; q++) {
        const int K = a_tile_col[q];
        pairs += b_tile_ptr[K + 1] - b_tile_ptr[K];
        h = mix64(h, ((unsigned long long)(unsigned)(K - I) << 32) | (unsigned)brow_id[K]);
    }
    w[x] = (int)(pairs < 0x7fffffff ? pairs : 0x7fffffff);
    if (pairs > MAXW || e - s > 0xFFFF) { *fail = 1; return; }
    const int slot = table_insert(keys, RROW_CAP, h, count, RROW_CAP / 2, fail);
    arow_slot[x] = slot;
    if (slot >= 0 && x < *(volatile int *)&owner[slot]) atomicMin(&owner[slot], x);
}

__global__ void __launch_bounds__(256)
k_arow_verify(int ntr, int trow0, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col, const int *__restrict__ brow_id,
              const int *__restrict__ arow_slot, const int *__restrict__ owner, const int *__restrict__ dense,
              int *__restrict__ arow_recipe, int *fail)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= ntr) return;
    const int slot = arow_slot[x];
    if (slot < 0) { *fail = 1; return; }
    const int I = trow0 + x, O = trow0 + owner[slot];
    const int s = a_tile_ptr[I], n = a_tile_ptr[I + 1] - s, so = a_tile_ptr[O];
    bool same = a_tile_ptr[O + 1] - so == n;
    for (int q = 0; q < n && same; q++)
        same = a_tile_col[s + q] - I == a_tile_col[so + q] - O && brow_id[a_tile_col[s + q]] == brow_id[a_tile_col[so + q]];
    if (!same) *fail = 2;
    arow_recipe[x] = dense[slot];
}

__global__ void __launch_bounds__(256)
k_flags(int cap, const int *__restrict__ owner, int *__restrict__ flags)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) flags[s] = owner[s] != NO_OWNER;
}

__global__ void __launch_bounds__(256)
k_reps(int cap, const int *__restrict__ owner, const int *__restrict__ dense, int *__restrict__ rep)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && owner[s] != NO_OWNER && dense[s] < cap / 2) rep[dense[s]] = owner[s];  // rep holds cap/2 entries
}

// One thread per distinct row recipe (representative tile-row x = rep[R], relative to trow0). Per recipe, MAXW-strided:
//   rp_D[R*MAXW + s]      s-th C tile column relative to I (ascending), s < rp_numJ[R]
//   rp_poff[R*MAXW + s]   start of C tile s's pairs inside the tile-row's pair range; rp_poff[.. + numJ] = pairs of the row
//   rp_pair[R*MAXW + q]   q-th pair of the tile-row in (C tile, ascending i) order: i << 16 | j
__global__ void __launch_bounds__(64)
k_rowplan_build(int nrec, int trow0, const int *__restrict__ rep, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col,
                const int *__restrict__ b_tile_ptr, const int *__restrict__ b_tile_col, int *__restrict__ rp_numJ, int *__restrict__ rp_D,
                int *__restrict__ rp_poff, unsigned *__restrict__ rp_pair)
{
    const int R = blockIdx.x * blockDim.x + threadIdx.x;
    if (R >= nrec) return;
    const int I = trow0 + rep[R];
    const int as = a_tile_ptr[I], na = a_tile_ptr[I + 1] - as;
    int *D = rp_D + (size_t)R * MAXW;
    int *poff = rp_poff + (size_t)R * (MAXW + 1);
    unsigned *pr = rp_pair + (size_t)R * MAXW;
    // sorted distinct D = J - I by insertion (numJ <= pairs <= MAXW)
    int numJ = 0;
    for (int i = 0; i < na; i++) {
        const int K = a_tile_col[as + i];
        for (int q = b_tile_ptr[K]; q < b_tile_ptr[K + 1]; q++) {
            const int d = b_tile_col[q] - I;
            int lo = 0, hi = numJ;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (D[mid] < d) lo = mid + 1; else hi = mid; }
            if (lo < numJ && D[lo] == d) continue;
            for (int z = numJ; z > lo; z--) D[z] = D[z - 1];
            D[lo] = d;
            numJ++;
        }
    }
    rp_numJ[R] = numJ;
    // pairs per C tile, ascending i: B's tile-row K_i holds column I + D[s] at most once (binary search)
    int run = 0;
    for (int s = 0; s < numJ; s++) {
        poff[s] = run;
        const int J = I + D[s];
        for (int i = 0; i < na; i++) {
            const int K = a_tile_col[as + i];
            int lo = b_tile_ptr[K], hi = b_tile_ptr[K + 1];
            const int base = lo;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (b_tile_col[mid] < J) lo = mid + 1; else hi = mid; }
            if (lo < b_tile_ptr[K + 1] && b_tile_col[lo] == J) pr[run++] = ((unsigned)i << 16) | (unsigned)(lo - base);
        }
    }
    poff[numJ] = run;
}

// c_cnt[x] = C tiles of tile-row x (scanned by the caller into c_tile_ptr)
__global__ void __launch_bounds__(256)
k_row_counts(int ntr, const int *__restrict__ arow_recipe, const int *__restrict__ rp_numJ, int *__restrict__ c_cnt)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < ntr) c_cnt[x] = rp_numJ[arow_recipe[x]];
}

// thread per (tile-row x, slot s < maxJ): C tile t = c_tile_ptr[x] + s
__global__ void __launch_bounds__(256)
k_expand_tiles(int ntr, int trow0, int maxJ, const int *__restrict__ arow_recipe, const int *__restrict__ rp_numJ,
               const int *__restrict__ rp_D, const int *__restrict__ rp_poff, const int *__restrict__ c_tile_ptr,
               const int *__restrict__ wptr, int *__restrict__ c_tile_col, int *__restrict__ c_tile_row, int *__restrict__ pair_ptr,
               int *__restrict__ pair_end)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(gid / maxJ), s = (int)(gid % maxJ);
    if (x >= ntr) return;
    const int R = arow_recipe[x];
    if (s >= rp_numJ[R]) return;
    const int t = c_tile_ptr[x] + s;
    c_tile_col[t] = trow0 + x + rp_D[(size_t)R * MAXW + s];
    c_tile_row[t] = trow0 + x;
    pair_ptr[t] = wptr[x] + rp_poff[(size_t)R * (MAXW + 1) + s];
    pair_end[t] = wptr[x] + rp_poff[(size_t)R * (MAXW + 1) + s + 1];
}

// thread per (tile-row x, pair q < maxW)
__global__ void __launch_bounds__(256)
k_expand_pairs(int ntr, int trow0, int maxW, const int *__restrict__ arow_recipe, const int *__restrict__ w, const unsigned *__restrict__ rp_pair,
               const int *__restrict__ wptr, const int *__restrict__ a_tile_ptr, const int *__restrict__ a_tile_col,
               const int *__restrict__ b_tile_ptr, const int *__restrict__ b_rm2csc, int *__restrict__ pair_a, int *__restrict__ pair_b)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(gid / maxW), q = (int)(gid % maxW);
    if (x >= ntr || q >= w[x]) return;
    const unsigned e = rp_pair[(size_t)arow_recipe[x] * MAXW + q];
    const int a = a_tile_ptr[trow0 + x] + (int)(e >> 16);
    const int K = a_tile_col[a];
    pair_a[wptr[x] + q] = a;
    pair_b[wptr[x] + q] = b_rm2csc[b_tile_ptr[K] + (int)(e & 0xFFFFu)];
}

}  // namespace rowplans
