// gentile_emu.h -- TEST INFRASTRUCTURE: lets spgemm_b200/csrc/gentile.cu compile as plain C++ (GT_EMULATE) so that the
// CPU tests can run its kernels, one emulated thread after another, and all of its host orchestration against the oracle.
// "Device" memory is host memory; a launch of n items runs thread ids 0 .. roundup(n, 256) - 1 in order, so the bounds
// guard of every kernel is exercised too. What this cannot show is a data race: the kernels of gentile.cu are written
// so that every output element has one owner, and the only cross-thread operations are integer atomicAdd / atomicOr.
// The library's look-back scan and radix sort (scan.cuh, radix_sort.cuh -- warp-level code, validated on the GPU by the
// 16 x 16 path's tests) are replaced by their specification: a serial exclusive scan and std::stable_sort.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/tilespgemm.h"

#define GT_KERNEL static void
#define GT_DEVICE static inline
#define GT_TID (tsg::gt_tid)
#define GT_POPC(x) __builtin_popcount((unsigned)(x))
#define GT_LAUNCH(kern, n, ...)                                                   \
    do {                                                                          \
        long long n_ = (long long)(n);                                            \
        if (n_ > 0) {                                                             \
            const long long pad_ = (n_ + 255) / 256 * 256;                        \
            for (tsg::gt_tid = 0; tsg::gt_tid < pad_; tsg::gt_tid++) kern(__VA_ARGS__); \
            tsg::ctx().launches++;                                                \
        }                                                                         \
    } while (0)

typedef int cudaStream_t;
enum { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline int cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { if (n) memmove(d, s, n); return 0; }
static inline int cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
#define CK(call) do { if ((call) != 0) return tsg::last_error(); } while (0)

static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }

namespace tsg {

static long long gt_tid = 0;
struct Ctx { cudaStream_t stream = 0; long long launches = 0; };
static Ctx g_emu_ctx;
static inline Ctx &ctx() { return g_emu_ctx; }

static int g_emu_err = 0;
static char g_emu_msg[512] = "";
static inline int last_error() { return g_emu_err; }
static inline void set_error(int code, const char *fmt, ...)
{
    if (g_emu_err) return;
    g_emu_err = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_emu_msg, sizeof(g_emu_msg), fmt, ap);
    va_end(ap);
}

// every block is padded by a guard zone filled with a pattern, checked on free: an emulated thread that writes past
// the end of an array fails the test instead of corrupting the heap
static const size_t GUARD = 64;
static inline void *dalloc(size_t bytes)
{
    if (!bytes) bytes = 256;
    char *p = (char *)malloc(bytes + GUARD + 16);
    if (!p) { set_error(TSG_ERR_NOMEM, "emu: malloc(%zu)", bytes); return nullptr; }
    *(size_t *)p = bytes;
    memset(p + 16, 0xA5, bytes);          // "uninitialised device memory"
    memset(p + 16 + bytes, 0x5C, GUARD);
    return p + 16;
}
static inline void dfree(void *q)
{
    if (!q) return;
    char *p = (char *)q - 16;
    const size_t bytes = *(size_t *)p;
    for (size_t k = 0; k < GUARD; k++)
        if ((unsigned char)p[16 + bytes + k] != 0x5C) { set_error(TSG_ERR_CUDA, "emu: write past the end of a %zu-byte block", bytes); break; }
    free(p);
}
template <typename T> static inline T *dalloc_n(size_t n) { return (T *)dalloc((n ? n : 1) * sizeof(T)); }

template <typename OutT> static int exclusive_scan(const int *in, OutT *out, long long n, long long *total64 = nullptr)
{
    long long run = 0;
    for (long long i = 0; i < n; i++) { const int v = in[i]; out[i] = (OutT)run; run += v; }
    out[n] = (OutT)run;
    if (total64) *total64 = run;
    return 0;
}
static inline int read_back_i32(const int *d, int *out) { *out = *d; return 0; }
static inline int read_back_i64(const long long *d, long long *out) { *out = *d; return 0; }

// specification of radix_sort_pairs (stable, by the low key_bits bits of the key); the result lands in the "b" buffers
// after an odd number of 8-bit passes and in the "a" buffers after an even number, like the device code
static inline int sort_pairs_device(uint32_t *ka, uint32_t *va, uint32_t *kb, uint32_t *vb, long long n, int key_bits, uint32_t **kres,
                                    uint32_t **vres)
{
    *kres = ka; *vres = va;
    if (n <= 0) return 0;
    int npass = (key_bits + 7) / 8;
    if (npass < 1) npass = 1;
    std::vector<uint32_t> idx((size_t)n);
    for (long long i = 0; i < n; i++) idx[(size_t)i] = (uint32_t)i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return ka[x] < ka[y]; });
    uint32_t *ko = (npass & 1) ? kb : ka, *vo = (npass & 1) ? vb : va;
    std::vector<uint32_t> k2((size_t)n), v2((size_t)n);
    for (long long i = 0; i < n; i++) { k2[(size_t)i] = ka[idx[(size_t)i]]; v2[(size_t)i] = va[idx[(size_t)i]]; }
    // the ping-pong leaves garbage in the other pair of buffers
    for (long long i = 0; i < n; i++) { ka[i] = 0xDEADBEEF; va[i] = 0xDEADBEEF; kb[i] = 0xDEADBEEF; vb[i] = 0xDEADBEEF; }
    memcpy(ko, k2.data(), (size_t)n * 4);
    memcpy(vo, v2.data(), (size_t)n * 4);
    *kres = ko; *vres = vo;
    return 0;
}

struct GtTimer {
    void mark(int) {}
    double ms(int, int) { return 0.0; }
};

}  // namespace tsg
