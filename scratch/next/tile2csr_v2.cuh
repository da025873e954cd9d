// tile2csr_v2.cuh -- DRAFT for round 2: tiles -> CSR without the serial walk of a tile-row by one half-warp
// (the long pole of the R-MAT end-to-end path: hub tile-rows hold 10^5 tiles; 7.3 ms on config 2).
// Not compiled into the library; checked by serial host emulation (test_tile2csr_emul.py) against the oracle.
//
// Idea: lay the per-(tile, row) entry counts out so that ONE global exclusive scan yields every CSR offset. For tile-row I
// with tiles [s, e) the count of (tile t, row r) goes to  perm = 16*s + r*(e-s) + (t-s):  the 16 rows of the tile-row one
// after the other, each row's tiles in ascending tile column. The scan value at that position is then the number of
// nonzeros in all earlier rows plus those of the same row in earlier tiles -- exactly where the tile's row segment starts
// in the CSR arrays; the value at the head of a run is the row pointer. Every kernel is one thread per (tile, row).
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#include "emul.h"
#endif
#ifndef TS
#define TS 16
#endif

namespace t2c {

__device__ __forceinline__ size_t perm_index(const int *__restrict__ tile_ptr, int I, int t, int r)
{
    const int s = tile_ptr[I], len = tile_ptr[I + 1] - s;
    return (size_t)TS * s + (size_t)r * len + (t - s);
}

// counts[perm] = entries of row r in tile t (tile_rowidx gives the tile-row of a tile, relative to the slab)
__global__ void __launch_bounds__(256)
k_counts(int numtile, const int *__restrict__ tile_ptr, const int *__restrict__ tile_rowidx, int trow0, const int *__restrict__ tile_nnz,
         const uint16_t *__restrict__ ptr, int *__restrict__ counts)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = (int)(gid >> 4), r = (int)(gid & 15);
    if (t >= numtile) return;
    const int tnnz = tile_nnz[t + 1] - tile_nnz[t];
    const int p0 = ptr[(size_t)t * TS + r], p1 = r < TS - 1 ? (int)ptr[(size_t)t * TS + r + 1] : tnnz;
    counts[perm_index(tile_ptr, tile_rowidx[t] - trow0, t, r)] = tnnz ? p1 - p0 : 0;
}

// rowptr[row] = scan value at the head of the row's run (thread per matrix row; rows of empty tile-rows read the head of
// the next block, which is the same running total). offs[16*numtile] holds the total.
__global__ void __launch_bounds__(256)
k_rowptr(int m, const int *__restrict__ tile_ptr, const int *__restrict__ offs, int numtile, int base, int *__restrict__ rowptr)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row > m) return;
    if (row == m) { rowptr[m] = offs[(size_t)TS * numtile] + base; return; }
    const int I = row >> 4, r = row & 15;
    const int s = tile_ptr[I], len = tile_ptr[I + 1] - s;
    rowptr[row] = offs[(size_t)TS * s + (size_t)r * len] + base;
}

__global__ void __launch_bounds__(256)
k_fill(int numtile, const int *__restrict__ tile_ptr, const int *__restrict__ tile_rowidx, int trow0, const int *__restrict__ tile_col,
       const int *__restrict__ tile_nnz, const uint16_t *__restrict__ ptr, const uint16_t *__restrict__ col,
       const double *__restrict__ val, const int *__restrict__ offs, int *__restrict__ out_col, double *__restrict__ out_val)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = (int)(gid >> 4), r = (int)(gid & 15);
    if (t >= numtile) return;
    const int b = tile_nnz[t], tnnz = tile_nnz[t + 1] - b;
    if (tnnz == 0) return;
    const int p0 = ptr[(size_t)t * TS + r], p1 = r < TS - 1 ? (int)ptr[(size_t)t * TS + r + 1] : tnnz;
    int dst = offs[perm_index(tile_ptr, tile_rowidx[t] - trow0, t, r)];
    const int cb = tile_col[t] * TS;
    for (int j = p0; j < p1; j++, dst++) {
        out_col[dst] = cb + col[b + j];
        out_val[dst] = val[b + j];
    }
}

}  // namespace t2c
