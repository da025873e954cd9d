// common.cuh -- library context, error latch, stream-ordered allocation, launch accounting.
// B200 (sm_100a) only. No CPU fallback anywhere: a failed CUDA call latches an error and the
// C-ABI entry point returns it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/tilespgemm.h"

#define TS 16             // tile edge (reference src/common.h:36 BLOCK_SIZE; MaskBits = 16)
#define FULL_MASK 0xffffffffu

namespace tsg {

struct Ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    int num_sms = 148;
    size_t smem_optin = 0;
    long long launches = 0;
    // scan workspace (decoupled look-back tile states + ticket)
    unsigned long long *scan_state = nullptr;
    size_t scan_state_cap = 0;
    int *scan_ticket = nullptr;
    // grow-only bump arenas for the per-call scratch of spgemm_device (row-sized arrays, pair lists,
    // block->tile map): after the largest slab has been seen no allocation happens in a step at all
    struct Arena { char *base = nullptr; size_t cap = 0, off = 0; } arena[3];
    // one-deep cache of the two buffers of a C slab (metadata, payload): a slab loop frees a C slab and asks for
    // the next one of similar size; handing the same buffer back avoids 20-GB pool re-allocations
    struct SlabCache { void *p = nullptr; size_t cap = 0; } cslab[2];
    // end-to-end path (tsg_spgemm_to_host): a second stream for the device->host copies and two grow-only landing
    // buffers, so that the CSR of slab s crosses PCIe while slab s+1 is computed
    cudaStream_t copy_stream = nullptr;
    struct Landing { char *p = nullptr; size_t cap = 0; cudaEvent_t ready = nullptr, copied = nullptr; bool busy = false; } land[2];
    // small pinned host scratch for scalar read-backs
    long long *h_scalars = nullptr;   // pinned + mapped, 32 slots
    long long *h_scalars_dev = nullptr;  // the same memory as the device sees it
    long long *d_scalars = nullptr;   // device, 32 slots
};

Ctx &ctx();
bool ctx_ready();

void set_error(int code, const char *fmt, ...);
int last_error();

// Returns false (and latches TSG_ERR_CUDA) on failure.
bool cuda_ok(cudaError_t e, const char *what, const char *file, int line);
#define CK(call)                                                       \
    do {                                                               \
        if (!tsg::cuda_ok((call), #call, __FILE__, __LINE__)) return tsg::last_error(); \
    } while (0)
#define CKV(call)                                                      \
    do {                                                               \
        if (!tsg::cuda_ok((call), #call, __FILE__, __LINE__)) return;  \
    } while (0)
// after a kernel launch
#define CK_LAUNCH()                                                    \
    do {                                                               \
        tsg::ctx().launches++;                                         \
        if (!tsg::cuda_ok(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)) return tsg::last_error(); \
    } while (0)

// Stream-ordered allocation from the library pool (no cudaMalloc/cudaFree synchronisation in the
// hot loop -- the reference calls cudaMalloc ~25x inside its timed region, tilespgemm-cuda.h:2417-2638).
void *dalloc(size_t bytes);
void dfree(void *p);
template <typename T> static inline T *dalloc_n(size_t n) { return (T *)dalloc((n ? n : 1) * sizeof(T)); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Arena `which`: make room for `bytes` (grow-only, stream-ordered re-allocation) and rewind it.
bool arena_reserve(int which, size_t bytes);
// Bump-allocate n elements (256-byte aligned) from arena `which`; nullptr + latched error if it does not fit.
void *arena_take_bytes(int which, size_t bytes);
template <typename T> static inline T *arena_take(int which, size_t n) { return (T *)arena_take_bytes(which, (n ? n : 1) * sizeof(T)); }
static inline size_t arena_need(size_t n, size_t elem) { return (((n ? n : 1) * elem) + 255) & ~(size_t)255; }

// C slab buffers: take a cached buffer of >= bytes (or allocate one with headroom; *cap gets its capacity) and
// give one back on free (the larger of cached / returned is kept).
void *cslab_take(int which, size_t bytes, size_t *cap);
void cslab_give(int which, void *p, size_t cap);

// Enqueue a copy of `nwords` 32-bit words from device memory to a slot of ctx().h_scalars, done by a KERNEL that
// stores into the mapped host memory. Scalar read-backs must not use the copy engine: while the end-to-end path is
// streaming a C slab to the host, a 4-byte D2H copy would queue behind hundreds of MB and stall the next slab.
int publish_words(void *h_dst, const void *d_src, int nwords);
// Small device->device copy done by a kernel, for the same reason (inside spgemm_device only).
int copy_words(void *d_dst, const void *d_src, size_t nwords);
// Read one device int / long long back (publish + stream sync). Used only where a size is needed for an
// allocation (numblkC, nnzC): two per SpGEMM call instead of the reference's ~8.
int read_back_i32(const int *d, int *out);
int read_back_i64(const long long *d, long long *out);

}  // namespace tsg
