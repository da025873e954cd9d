// emul.h -- run thread-per-item CUDA kernels on the host, one "thread" after the other, to check their index logic
// against the oracle without a GPU. Only for kernels that use no warp intrinsics and no __syncthreads: with serial
// execution an atomic protocol never observes an intermediate state, so spin loops fall through.
// NOT part of the product; see scratch/next/README.md.
#pragma once
#ifndef __CUDACC__
#include <stdint.h>
#include <string.h>
#include <cmath>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct uint3_e { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static uint3_e threadIdx, blockIdx, blockDim, gridDim;

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline unsigned __brev(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline void __threadfence() {}
template <typename T> static inline T atomicCAS(T *p, T cmp, T val) { T old = *p; if (old == cmp) *p = val; return old; }
template <typename T> static inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
template <typename T> static inline T atomicExch(T *p, T v) { T old = *p; *p = v; return old; }
template <typename T> static inline T atomicMax(T *p, T v) { T old = *p; if (v > old) *p = v; return old; }
template <typename T> static inline T atomicMin(T *p, T v) { T old = *p; if (v < old) *p = v; return old; }

// LAUNCH(kernel, blocks, threads, args...): every thread of every block, serially
#define LAUNCH(kern, nblocks, nthreads, ...)                                                   \
    do {                                                                                       \
        gridDim = {(unsigned)(nblocks), 1, 1}; blockDim = {(unsigned)(nthreads), 1, 1};        \
        for (unsigned b_ = 0; b_ < (unsigned)(nblocks); b_++)                                  \
            for (unsigned t_ = 0; t_ < (unsigned)(nthreads); t_++) {                           \
                blockIdx = {b_, 0, 0}; threadIdx = {t_, 0, 0};                                 \
                kern(__VA_ARGS__);                                                             \
            }                                                                                  \
    } while (0)
#endif
