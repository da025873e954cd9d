// scan.cuh -- single-pass exclusive prefix sum with decoupled look-back.
// Replaces the reference's thrust::exclusive_scan calls (src/utils_cuda_scan.h:267-286, used at
// tilespgemm-cuda.h:2403,2602) and its host exclusive_scan (src/utils.h:36-51).
//
// Semantics (same as the reference's in-place convention): given counts in[0..n), writes
// out[i] = sum_{j<i} in[j] for i in [0, n] -- n+1 outputs, out[n] = total. out may alias in
// (the array then needs n+1 slots, the last one being ignored on input).
// Output type int or long long (64-bit where the total can pass 2^31).
//
// One pass over the data: each CTA takes a 2048-element tile in ticket order, publishes its
// aggregate, and resolves its exclusive prefix by looking back over the published
// (flag | value) words of its predecessors, 32 at a time.
#pragma once
#include "common.cuh"

namespace tsg {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long scan_ld_state(const unsigned long long *p)
{
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

template <typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(const int *in, OutT *out, long long n, unsigned long long *state, int *ticket, long long *total)
{
    __shared__ int s_tile;
    __shared__ long long s_warp[SCAN_THREADS / 32];
    __shared__ long long s_prefix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const long long base = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;

    int v[SCAN_ITEMS];
    long long sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long long idx = base + k;
        v[k] = idx < n ? in[idx] : 0;
        sum += v[k];
    }
    long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long warp_off = 0, aggregate = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        long long t = s_warp[w];
        if (w < warp) warp_off += t;
        aggregate += t;
    }
    if (warp == 0) {
        long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&state[0], (2ull << 62) | ((unsigned long long)aggregate & SCAN_VAL_MASK));
        } else {
            if (lane == 0) atomicExch(&state[tile], (1ull << 62) | ((unsigned long long)aggregate & SCAN_VAL_MASK));
            int look = tile - 1;
            while (true) {
                int idx = look - lane;
                unsigned long long s = 2ull << 62;  // before the first tile: inclusive prefix 0
                if (idx >= 0) {
                    do { s = scan_ld_state(&state[idx]); } while ((s >> 62) == 0);
                }
                unsigned incl_mask = __ballot_sync(FULL_MASK, (s >> 62) == 2);
                long long val = (long long)(s & SCAN_VAL_MASK);
                int first = incl_mask ? (__ffs(incl_mask) - 1) : 31;
                long long contrib = lane <= first ? val : 0;
#pragma unroll
                for (int o = 16; o; o >>= 1) contrib += __shfl_xor_sync(FULL_MASK, contrib, o);
                prefix += contrib;
                if (incl_mask) break;
                look -= 32;
            }
            if (lane == 0)
                atomicExch(&state[tile], (2ull << 62) | ((unsigned long long)(prefix + aggregate) & SCAN_VAL_MASK));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    long long run = s_prefix + warp_off + (incl - sum);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long long idx = base + k;
        if (idx <= n) out[idx] = (OutT)run;
        if (idx == n && total) *total = run;  // the 64-bit total next to narrowed 32-bit offsets
        run += v[k];
    }
}

// Host wrapper; enqueues on the library stream. Returns TSG_OK or an error code.
// total64 (optional, device pointer): receives the sum as a 64-bit value even when OutT is int, so that one scan gives
// 32-bit offsets for the kernels and an overflow-proof total for the host (the reference scans int totals).
template <typename OutT>
static int exclusive_scan(const int *in, OutT *out, long long n, long long *total64 = nullptr)
{
    Ctx &c = ctx();
    long long ntiles = (n + 1 + SCAN_TILE - 1) / SCAN_TILE;
    if ((size_t)ntiles > c.scan_state_cap) {
        if (c.scan_state) dfree(c.scan_state);
        size_t cap = (size_t)ntiles * 2;
        c.scan_state = (unsigned long long *)dalloc(cap * sizeof(unsigned long long));
        if (!c.scan_state) return last_error();
        c.scan_state_cap = cap;
    }
    CK(cudaMemsetAsync(c.scan_state, 0, (size_t)ntiles * sizeof(unsigned long long), c.stream));
    CK(cudaMemsetAsync(c.scan_ticket, 0, sizeof(int), c.stream));
    scan_lookback_kernel<OutT><<<(unsigned)ntiles, SCAN_THREADS, 0, c.stream>>>(in, out, n, c.scan_state, c.scan_ticket, total64);
    CK_LAUNCH();
    return TSG_OK;
}

}  // namespace tsg
