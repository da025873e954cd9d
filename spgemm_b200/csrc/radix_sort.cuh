// radix_sort.cuh -- stable LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits per pass.
//
// Used for the two stable transpositions on the path:
//   * CSR -> CSC of a whole matrix (reference matrix_transposition, src/utils.h:161-198: the
//     counting sort there is stable, so each output column lists rows in ascending order), and
//   * the tile-level transposition inside csr2tile_col_major (reference src/csr2tile.h:300-347,
//     which gets CSC-tile order by tiling B^T).
// Keys arrive ordered by (row, col), so a stable sort by column alone yields (col, row) order.
//
// A pass is histogram -> exclusive scan (scan.cuh) -> stable scatter. The unit of work is a WARP
// owning a contiguous chunk of RS_ITEMS_PER_WARP items; its 256 digit counters live in shared
// memory. Stability inside a warp comes from __match_any_sync: lanes holding the same digit are
// ranked by lane id, strips of 32 items are processed in order.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace tsg {

constexpr int RS_WARPS = 8;                      // warps per CTA
constexpr int RS_STRIPS = 64;                    // strips of 32 items per warp
constexpr int RS_ITEMS_PER_WARP = 32 * RS_STRIPS;

__global__ void __launch_bounds__(RS_WARPS * 32)
rs_hist_kernel(const uint32_t *__restrict__ keys, long long n, int shift, int nwarps_total, int *__restrict__ table)
{
    __shared__ int h[RS_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = lane; i < 256; i += 32) h[w][i] = 0;
    __syncwarp();
    const long long gw = (long long)blockIdx.x * RS_WARPS + w;
    if (gw < nwarps_total) {
        const long long base = gw * RS_ITEMS_PER_WARP;
        for (int s = 0; s < RS_STRIPS; s++) {
            long long i = base + s * 32 + lane;
            if (i < n) atomicAdd(&h[w][(keys[i] >> shift) & 255], 1);
        }
        __syncwarp();
        for (int d = lane; d < 256; d += 32) table[(long long)d * nwarps_total + gw] = h[w][d];
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
rs_scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, long long n, int shift,
                  int nwarps_total, const int *__restrict__ table, uint32_t *__restrict__ keys_out,
                  uint32_t *__restrict__ vals_out)
{
    __shared__ int cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * RS_WARPS + w;
    if (gw >= nwarps_total) return;
    for (int d = lane; d < 256; d += 32) cnt[w][d] = table[(long long)d * nwarps_total + gw];
    __syncwarp();
    const long long base = gw * RS_ITEMS_PER_WARP;
    const unsigned lt = (1u << lane) - 1;
    for (int s = 0; s < RS_STRIPS; s++) {
        long long i = base + s * 32 + lane;
        bool valid = i < n;
        unsigned active = __ballot_sync(FULL_MASK, valid);
        if (!active) break;
        if (valid) {
            uint32_t k = keys[i], v = vals ? vals[i] : (uint32_t)i;
            int d = (k >> shift) & 255;
            unsigned peers = __match_any_sync(active, d);
            int rank = __popc(peers & lt);
            int pos = cnt[w][d];
            __syncwarp(active);
            if (rank == 0) cnt[w][d] = pos + __popc(peers);
            __syncwarp(active);
            keys_out[pos + rank] = k;
            vals_out[pos + rank] = v;
        }
    }
}

// Sorts n pairs by the low `key_bits` bits of the key, stably. keys_a/vals_a hold the input
// (vals_a == nullptr means the identity permutation 0..n-1); *_b are scratch of the same size.
// On return *keys_res / *vals_res point at whichever buffer holds the result (a or b).
// n must be < 2^31 (positions are int).
static int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, long long n,
                            int key_bits, uint32_t **keys_res, uint32_t **vals_res, uint32_t *vals_a_storage)
{
    Ctx &c = ctx();
    *keys_res = keys_a;
    *vals_res = vals_a ? vals_a : vals_a_storage;
    if (n <= 0) return TSG_OK;
    int npass = (key_bits + 7) / 8;
    if (npass < 1) npass = 1;
    int nwarps = ceil_div(n, RS_ITEMS_PER_WARP);
    int nblocks = ceil_div(nwarps, RS_WARPS);
    int *table = dalloc_n<int>((size_t)256 * nwarps + 1);
    if (!table) return last_error();
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int p = 0; p < npass; p++) {
        int shift = 8 * p;
        rs_hist_kernel<<<nblocks, RS_WARPS * 32, 0, c.stream>>>(kin, n, shift, nwarps, table);
        CK_LAUNCH();
        int rc = exclusive_scan<int>(table, table, (long long)256 * nwarps);
        if (rc) return rc;
        rs_scatter_kernel<<<nblocks, RS_WARPS * 32, 0, c.stream>>>(kin, vin, n, shift, nwarps, table, kout, vout);
        CK_LAUNCH();
        // ping-pong; after the first pass the "a" value buffer is free to be written
        uint32_t *tk = kin, *tv = vin ? vin : vals_a_storage;
        kin = kout; vin = vout; kout = tk; vout = tv;
    }
    *keys_res = kin;
    *vals_res = vin;
    dfree(table);
    return TSG_OK;
}

}  // namespace tsg
