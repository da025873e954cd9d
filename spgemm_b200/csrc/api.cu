// api.cu -- the C ABI of libtilespgemm_b200.so (include/tilespgemm.h): library context, the
// reference-named drop-in entry points (host buffers) and the device-resident tsg_* API.
#include <stdarg.h>
#include <unordered_map>
#include <vector>
#include "common.cuh"
#include "kernels.h"

namespace tsg {

static Ctx g_ctx;
static bool g_ready = false;
static int g_err = TSG_OK;
static char g_errmsg[512] = "";

Ctx &ctx() { return g_ctx; }
bool ctx_ready() { return g_ready; }
int last_error() { return g_err; }

void set_error(int code, const char *fmt, ...)
{
    if (g_err != TSG_OK) return;  // keep the first error
    g_err = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_errmsg, sizeof(g_errmsg), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[tilespgemm_b200] ERROR %d: %s\n", code, g_errmsg);
}

bool cuda_ok(cudaError_t e, const char *what, const char *file, int line)
{
    if (e == cudaSuccess) return true;
    set_error(e == cudaErrorMemoryAllocation ? TSG_ERR_NOMEM : TSG_ERR_CUDA, "%s failed at %s:%d: %s", what, file, line,
              cudaGetErrorString(e));
    return false;
}

// Big blocks (>= 32 MB) are recycled by this small exact-fit cache instead of going back to the stream-ordered
// pool: the pool splits cached blocks to serve smaller requests, so a steady-state loop that allocates the same
// multi-GB sizes every step would now and then fall through to the driver (hundreds of ms for a 3-GB block;
// measured in bench.py's end-to-end leg). All work is ordered on the one library stream, so a block freed here
// may be handed out again immediately.
struct BigBlock { void *p; size_t cap; };
static std::vector<BigBlock> g_big_free;                // cached free blocks, oldest first
static std::unordered_map<void *, size_t> g_big_live;   // live big blocks -> capacity
static size_t g_big_free_bytes = 0;
constexpr size_t BIG_MIN = (size_t)32 << 20, BIG_CACHE_MAX_BYTES = (size_t)64 << 30;
constexpr size_t BIG_CACHE_MAX_BLOCKS = 48;

static void big_cache_trim(size_t max_bytes, size_t max_blocks)
{
    while (!g_big_free.empty() && (g_big_free_bytes > max_bytes || g_big_free.size() > max_blocks)) {
        cudaFreeAsync(g_big_free.front().p, g_ctx.stream);
        g_big_free_bytes -= g_big_free.front().cap;
        g_big_free.erase(g_big_free.begin());
    }
}

static void *dalloc_impl(size_t bytes, bool quiet);
void *dalloc(size_t bytes) { return dalloc_impl(bytes, false); }

// quiet: a failure returns nullptr without latching an error (the caller has a smaller request to fall back to)
static void *dalloc_impl(size_t bytes, bool quiet)
{
    void *p = nullptr;
    if (bytes == 0) bytes = 256;
    if (bytes >= BIG_MIN) {
        int best = -1;
        for (size_t k = 0; k < g_big_free.size(); k++)
            if (g_big_free[k].cap >= bytes && g_big_free[k].cap <= bytes + bytes / 2 && (best < 0 || g_big_free[k].cap < g_big_free[best].cap))
                best = (int)k;
        if (best >= 0) {
            BigBlock b = g_big_free[best];
            g_big_free.erase(g_big_free.begin() + best);
            g_big_free_bytes -= b.cap;
            g_big_live[b.p] = b.cap;
            return b.p;
        }
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, g_ctx.stream);
    if (e != cudaSuccess && !g_big_free.empty()) {  // give the cached blocks back and retry once
        cudaGetLastError();
        big_cache_trim(0, 0);
        cudaStreamSynchronize(g_ctx.stream);
        e = cudaMallocAsync(&p, bytes, g_ctx.stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (!quiet) set_error(TSG_ERR_NOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    if (bytes >= BIG_MIN) g_big_live[p] = bytes;
    return p;
}

void dfree(void *p)
{
    if (!p) return;
    auto it = g_big_live.find(p);
    if (it == g_big_live.end()) { cudaFreeAsync(p, g_ctx.stream); return; }
    g_big_free.push_back(BigBlock{p, it->second});
    g_big_free_bytes += it->second;
    g_big_live.erase(it);
    big_cache_trim(BIG_CACHE_MAX_BYTES, BIG_CACHE_MAX_BLOCKS);
}

bool arena_reserve(int which, size_t bytes)
{
    Ctx::Arena &a = g_ctx.arena[which];
    a.off = 0;
    if (bytes <= a.cap) return true;
    if (a.base) dfree(a.base);
    a.base = nullptr; a.cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    a.base = (char *)dalloc(want);
    if (!a.base) return false;
    a.cap = want;
    return true;
}

void *arena_take_bytes(int which, size_t bytes)
{
    Ctx::Arena &a = g_ctx.arena[which];
    size_t need = (bytes + 255) & ~(size_t)255;
    if (!a.base || a.off + need > a.cap) {
        set_error(TSG_ERR_NOMEM, "internal: scratch arena %d overflow (%zu + %zu > %zu)", which, a.off, need, a.cap);
        return nullptr;
    }
    void *p = a.base + a.off;
    a.off += need;
    return p;
}

void *cslab_take(int which, size_t bytes, size_t *cap)
{
    Ctx::SlabCache &s = g_ctx.cslab[which];
    if (s.p && s.cap >= bytes) {
        void *p = s.p;
        *cap = s.cap;
        s.p = nullptr; s.cap = 0;
        return p;
    }
    size_t want = bytes + bytes / 8 + 4096;
    void *p = dalloc_impl(want, true);  // headroom is optional: only the exact-size retry may latch an error
    if (!p) { want = bytes; p = dalloc(want); }
    *cap = p ? want : 0;
    return p;
}

void cslab_give(int which, void *p, size_t cap)
{
    if (!p) return;
    Ctx::SlabCache &s = g_ctx.cslab[which];
    if (!s.p || cap > s.cap) {
        if (s.p) dfree(s.p);
        s.p = p; s.cap = cap;
    } else {
        dfree(p);
    }
}

__global__ void k_publish(const int *__restrict__ src, int *__restrict__ dst, int nwords)
{
    if ((int)threadIdx.x < nwords) dst[threadIdx.x] = src[threadIdx.x];
}

int publish_words(void *h_dst, const void *d_src, int nwords)
{
    int *dst = (int *)((char *)g_ctx.h_scalars_dev + ((char *)h_dst - (char *)g_ctx.h_scalars));
    k_publish<<<1, 32, 0, g_ctx.stream>>>((const int *)d_src, dst, nwords);
    CK_LAUNCH();
    return TSG_OK;
}

__global__ void k_copy_words(const int *__restrict__ src, int *__restrict__ dst, size_t nwords)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int copy_words(void *d_dst, const void *d_src, size_t nwords)
{
    if (!nwords) return TSG_OK;
    const size_t blocks = (nwords + 255) / 256;
    k_copy_words<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, g_ctx.stream>>>((const int *)d_src, (int *)d_dst, nwords);
    CK_LAUNCH();
    return TSG_OK;
}

int read_back_i32(const int *d, int *out)
{
    int rc = publish_words(&g_ctx.h_scalars[15], d, 1);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    *out = *(const volatile int *)&g_ctx.h_scalars[15];
    return TSG_OK;
}

int read_back_i64(const long long *d, long long *out)
{
    int rc = publish_words(&g_ctx.h_scalars[15], d, 2);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    *out = *(const volatile long long *)&g_ctx.h_scalars[15];
    return TSG_OK;
}

static int ensure_init()
{
    if (g_ready) return TSG_OK;
    int dev = 0;
    if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice (is a CUDA device present?)", __FILE__, __LINE__)) return g_err;
    return tsg_init(dev);
}

// 16 x 16 runs the tuned kernels; every other representable size the general-tile path (gentile.cu)
static bool tiles_16(int tm, int tn) { return tm == TS && tn == TS; }
static bool tiles_general(int tm, int tn)
{
    if (gtile_size_ok(tm, tn)) return true;
    set_error(TSG_ERR_UNSUPPORTED, "tile size %d x %d: tile_size_m and tile_size_n must be multiples of 16 between 16 and 128", tm, tn);
    return false;
}

}  // namespace tsg

using namespace tsg;

extern "C" {

int tilespgemm_last_error(void) { return g_err; }
const char *tilespgemm_last_error_string(void) { return g_errmsg; }
void tilespgemm_clear_error(void) { g_err = TSG_OK; g_errmsg[0] = 0; }

int tsg_init(int device)
{
    if (g_ready && g_ctx.device == device) return TSG_OK;
    if (g_ready) tsg_shutdown();
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error(TSG_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return g_err;
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    g_ctx = Ctx();
    g_ctx.device = device;
    g_ctx.num_sms = prop.multiProcessorCount;
    g_ctx.smem_optin = prop.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    CK(cudaDeviceGetDefaultMemPool(&g_ctx.pool, device));
    unsigned long long thresh = ~0ull;  // keep freed blocks cached: no cudaMalloc/cudaFree in steady state
    CK(cudaMemPoolSetAttribute(g_ctx.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    CK(cudaMalloc(&g_ctx.scan_ticket, 256));
    CK(cudaMalloc(&g_ctx.d_scalars, 32 * sizeof(long long)));
    CK(cudaHostAlloc(&g_ctx.h_scalars, 32 * sizeof(long long), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&g_ctx.h_scalars_dev, g_ctx.h_scalars, 0));
    g_ready = true;
    return TSG_OK;
}

void tsg_shutdown(void)
{
    if (!g_ready) return;
    cudaStreamSynchronize(g_ctx.stream);
    plans_shutdown();
    rowplans_shutdown();
    if (g_ctx.scan_state) cudaFreeAsync(g_ctx.scan_state, g_ctx.stream);
    for (int k = 0; k < 3; k++)
        if (g_ctx.arena[k].base) cudaFreeAsync(g_ctx.arena[k].base, g_ctx.stream);
    for (int k = 0; k < 2; k++)
        if (g_ctx.cslab[k].p) cudaFreeAsync(g_ctx.cslab[k].p, g_ctx.stream);
    if (g_ctx.copy_stream) cudaStreamSynchronize(g_ctx.copy_stream);
    for (int k = 0; k < 2; k++) {
        if (g_ctx.land[k].p) dfree(g_ctx.land[k].p);
        if (g_ctx.land[k].ready) cudaEventDestroy(g_ctx.land[k].ready);
        if (g_ctx.land[k].copied) cudaEventDestroy(g_ctx.land[k].copied);
    }
    if (g_ctx.copy_stream) cudaStreamDestroy(g_ctx.copy_stream);
    big_cache_trim(0, 0);
    g_big_live.clear();
    cudaStreamSynchronize(g_ctx.stream);
    cudaFree(g_ctx.scan_ticket);
    cudaFree(g_ctx.d_scalars);
    cudaFreeHost(g_ctx.h_scalars);
    cudaStreamDestroy(g_ctx.stream);
    g_ctx = Ctx();
    g_ready = false;
}

void *tsg_stream(void) { return g_ready ? (void *)g_ctx.stream : nullptr; }

int tsg_sync(void)
{
    if (ensure_init()) return g_err;
    CK(cudaStreamSynchronize(g_ctx.stream));
    return TSG_OK;
}

long long tsg_launch_count(void) { return g_ctx.launches; }

static cudaEvent_t g_timer[2] = {nullptr, nullptr};

int tsg_timer_start(void)
{
    if (ensure_init()) return g_err;
    for (int k = 0; k < 2; k++)
        if (!g_timer[k]) CK(cudaEventCreate(&g_timer[k]));
    CK(cudaEventRecord(g_timer[0], g_ctx.stream));
    return TSG_OK;
}

int tsg_timer_stop(double *ms)
{
    if (ensure_init()) return g_err;
    if (!g_timer[0]) { set_error(TSG_ERR_UNSUPPORTED, "tsg_timer_stop without tsg_timer_start"); return g_err; }
    CK(cudaEventRecord(g_timer[1], g_ctx.stream));
    CK(cudaEventSynchronize(g_timer[1]));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, g_timer[0], g_timer[1]));
    *ms = f;
    return TSG_OK;
}

/* ------------------------------- CSR on the device ------------------------------- */

static int csr_alloc(int m, int n, long long nnz, tsg_dcsr *out)
{
    memset(out, 0, sizeof(*out));
    size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    size_t o_ci = (((size_t)m + 1) * 4 + 255) & ~(size_t)255, o_v = o_ci + ((nz * 4 + 255) & ~(size_t)255);
    char *base = (char *)dalloc(o_v + nz * 8);
    if (!base) return g_err;
    out->m = m; out->n = n; out->nnz = nnz; out->owner = base;
    out->rowptr = (int *)base; out->colidx = (int *)(base + o_ci); out->val = (double *)(base + o_v);
    return TSG_OK;
}

int tsg_csr_upload(int m, int n, const int *rowptr, const int *colidx, const double *val, tsg_dcsr *out)
{
    if (ensure_init()) return g_err;
    if (m < 0 || n < 0 || !rowptr) { set_error(TSG_ERR_INPUT, "tsg_csr_upload: bad arguments"); return g_err; }
    long long nnz = rowptr[m];
    if (nnz < 0) { set_error(TSG_ERR_INPUT, "tsg_csr_upload: rowptr[m] = %lld", nnz); return g_err; }
    int rc = csr_alloc(m, n, nnz, out);
    if (rc) return rc;
    if (!cuda_ok(cudaMemcpyAsync(out->rowptr, rowptr, ((size_t)m + 1) * 4, cudaMemcpyHostToDevice, g_ctx.stream), "upload rowptr", __FILE__, __LINE__)) {
        tsg_csr_free(out);
        return g_err;
    }
    if (nnz > 0 && (!cuda_ok(cudaMemcpyAsync(out->colidx, colidx, (size_t)nnz * 4, cudaMemcpyHostToDevice, g_ctx.stream), "upload colidx", __FILE__, __LINE__) ||
                    !cuda_ok(cudaMemcpyAsync(out->val, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, g_ctx.stream), "upload val", __FILE__, __LINE__))) {
        tsg_csr_free(out);
        return g_err;
    }
    rc = csr_check_device(out);  // column range / row pointer monotonicity, before any kernel indexes by them (synchronises)
    if (rc) { tsg_csr_free(out); return rc; }
    return TSG_OK;
}

int tsg_csr_wrap(int m, int n, long long nnz, int *d_rowptr, int *d_colidx, double *d_val, tsg_dcsr *out)
{
    if (ensure_init()) return g_err;
    memset(out, 0, sizeof(*out));
    out->m = m; out->n = n; out->nnz = nnz; out->rowptr = d_rowptr; out->colidx = d_colidx; out->val = d_val; out->owner = nullptr;
    int rc = csr_check_device(out);
    if (rc) memset(out, 0, sizeof(*out));
    return rc;
}

int tsg_csr_row_slice(const tsg_dcsr *a, int row0, int row1, tsg_dcsr *out)
{
    if (ensure_init()) return g_err;
    return csr_row_slice_device(a, row0, row1, out);
}

int tsg_csr_canonicalize(const tsg_dcsr *a, int dup_policy, tsg_dcsr *out)
{
    if (ensure_init()) return g_err;
    if (dup_policy != 0 && dup_policy != 1) { set_error(TSG_ERR_INPUT, "canonicalize: dup_policy must be 0 (sum) or 1 (keep first)"); return g_err; }
    int rc = csr_canonicalize_device(a, dup_policy, out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    return TSG_OK;
}

int tsg_csr_download(const tsg_dcsr *a, int *rowptr, int *colidx, double *val)
{
    if (ensure_init()) return g_err;
    if (rowptr) CK(cudaMemcpyAsync(rowptr, a->rowptr, ((size_t)a->m + 1) * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    if (a->nnz > 0) {
        if (colidx) CK(cudaMemcpyAsync(colidx, a->colidx, (size_t)a->nnz * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
        if (val) CK(cudaMemcpyAsync(val, a->val, (size_t)a->nnz * 8, cudaMemcpyDeviceToHost, g_ctx.stream));
    }
    CK(cudaStreamSynchronize(g_ctx.stream));
    return TSG_OK;
}

void tsg_csr_free(tsg_dcsr *a)
{
    if (!a) return;
    if (a->owner && g_ready) dfree(a->owner);
    memset(a, 0, sizeof(*a));
}

int tsg_csr_validate(const tsg_dcsr *a)
{
    // The contract is checked by the scatter kernel of csr2tile (it sees every entry and its left
    // neighbour anyway); a throw-away row-major conversion is the validation.
    if (ensure_init()) return g_err;
    tsg_dtile t;
    int rc = csr2tile_device(a, 0, &t);
    tsg_tile_free(&t);
    if (rc == TSG_OK) CK(cudaStreamSynchronize(g_ctx.stream));
    return rc;
}

int tsg_transpose(const tsg_dcsr *a, tsg_dcsr *at)
{
    if (ensure_init()) return g_err;
    int rc = transpose_device(a, at);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    return TSG_OK;
}

int tsg_nnzcub(const tsg_dcsr *a, const tsg_dcsr *b, unsigned long long *out)
{
    if (ensure_init()) return g_err;
    if (a->n != b->m) { set_error(TSG_ERR_UNSUPPORTED, "nnzcub: inner dimensions differ"); return g_err; }
    return nnzcub_device(a, b, out);
}

/* ------------------------------- tiles on the device ------------------------------- */

int tsg_csr2tile(const tsg_dcsr *a, int col_major, tsg_dtile *out)
{
    if (ensure_init()) return g_err;
    int rc = csr2tile_device(a, col_major, out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    return TSG_OK;
}

int tsg_tile_alloc(int m, int n, int numtile, long long nnz, int col_major, tsg_dtile *out)
{
    if (ensure_init()) return g_err;
    return tile_alloc_layout(m, n, numtile, nnz, col_major, out);
}

void tsg_tile_free(tsg_dtile *t)
{
    if (!t) return;
    for (int k = 0; k < 4; k++) {
        if (!t->slab[k] || !g_ready) continue;
        if (t->cached && k < 2) cslab_give(k, t->slab[k], t->slab_bytes[k]);  // a C slab of tsg_spgemm: keep for the next slab
        else dfree(t->slab[k]);
    }
    memset(t, 0, sizeof(*t));
}

int tsg_tile_upload(const SMatrix *h, int col_major, tsg_dtile *out)
{
    if (ensure_init()) return g_err;
    if (h->numtile < 0 || h->nnz < 0) { set_error(TSG_ERR_INPUT, "tsg_tile_upload: negative sizes"); return g_err; }
    int rc = tile_alloc_layout(h->m, h->n, h->numtile, h->nnz, col_major, out);
    if (rc) return rc;
    out->tilem = h->tilem; out->tilen = h->tilen;
    cudaStream_t s = g_ctx.stream;
    const size_t nt = (size_t)h->numtile, nz = (size_t)h->nnz;
    CK(cudaMemcpyAsync(out->tile_ptr, h->tile_ptr, ((size_t)h->tilem + 1) * 4, cudaMemcpyHostToDevice, s));
    if (nt) {
        CK(cudaMemcpyAsync(out->tile_columnidx, h->tile_columnidx, nt * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(out->ptr, h->tile_csr_Ptr, nt * TS * 2, cudaMemcpyHostToDevice, s));
        if (h->mask) CK(cudaMemcpyAsync(out->mask, h->mask, nt * TS * 2, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(out->tile_rowidx, 0, nt * 4, s));
    }
    CK(cudaMemcpyAsync(out->tile_nnz, h->tile_nnz, (nt + 1) * 4, cudaMemcpyHostToDevice, s));
    if (nz) {
        CK(cudaMemcpyAsync(out->val, h->tile_csr_Value, nz * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(out->col, h->tile_csr_Col, nz * 2, cudaMemcpyHostToDevice, s));
    }
    if (nt && !h->mask) {  // steps 1-3 read the row masks of A and B: rebuild them from Ptr / Col
        rc = masks_from_tiles_device(out);
        if (rc) return rc;
    }
    if (col_major) {
        CK(cudaMemcpyAsync(out->csc_tile_ptr, h->csc_tile_ptr, ((size_t)h->tilen + 1) * 4, cudaMemcpyHostToDevice, s));
        if (nt) CK(cudaMemcpyAsync(out->csc_tile_rowidx, h->csc_tile_rowidx, nt * 4, cudaMemcpyHostToDevice, s));
        rc = build_rm2csc_device(out);
        if (rc) return rc;
    }
    rc = tile_patterns_device(out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));
    return TSG_OK;
}

int tsg_tile_download(const tsg_dtile *t, SMatrix *h)
{
    if (ensure_init()) return g_err;
    if (t->nnz >= (1ll << 31)) { set_error(TSG_ERR_OVERFLOW, "tile download: nnz %lld does not fit SMatrix.nnz", t->nnz); return g_err; }
    cudaStream_t s = g_ctx.stream;
    const size_t nt = (size_t)t->numtile, nz = (size_t)t->nnz;
    h->m = t->m; h->n = t->n; h->tilem = t->tilem; h->tilen = t->tilen; h->numtile = t->numtile; h->nnz = (int)t->nnz;
    h->tile_ptr = (int *)malloc(((size_t)t->tilem + 1) * 4);
    h->tile_columnidx = (int *)malloc((nt ? nt : 1) * 4);
    h->tile_rowidx = (int *)calloc(nt ? nt : 1, 4);
    h->tile_nnz = (int *)malloc((nt + 1) * 4);
    h->tile_csr_Value = (double *)malloc((nz ? nz : 1) * 8);
    h->tile_csr_Col = (uint16_t *)malloc((nz ? nz : 1) * 2);
    h->tile_csr_Ptr = (uint16_t *)malloc((nt ? nt : 1) * TS * 2);
    h->mask = (uint16_t *)malloc((nt ? nt : 1) * TS * 2);
    h->csc_tile_ptr = nullptr; h->csc_tile_rowidx = nullptr;
    if (!h->tile_ptr || !h->tile_columnidx || !h->tile_rowidx || !h->tile_nnz || !h->tile_csr_Value || !h->tile_csr_Col ||
        !h->tile_csr_Ptr || !h->mask) {
        set_error(TSG_ERR_NOMEM, "tile download: host allocation failed");
        return g_err;
    }
    CK(cudaMemcpyAsync(h->tile_ptr, t->tile_ptr, ((size_t)t->tilem + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->tile_nnz, t->tile_nnz, (nt + 1) * 4, cudaMemcpyDeviceToHost, s));
    if (nt) {
        CK(cudaMemcpyAsync(h->tile_columnidx, t->tile_columnidx, nt * 4, cudaMemcpyDeviceToHost, s));
        // reference csr2tile_col_major allocates tile_rowidx and leaves it zero (src/csr2tile.h:336-337)
        if (!t->col_major) CK(cudaMemcpyAsync(h->tile_rowidx, t->tile_rowidx, nt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->tile_csr_Ptr, t->ptr, nt * TS * 2, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->mask, t->mask, nt * TS * 2, cudaMemcpyDeviceToHost, s));
    }
    if (nz) {
        CK(cudaMemcpyAsync(h->tile_csr_Value, t->val, nz * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->tile_csr_Col, t->col, nz * 2, cudaMemcpyDeviceToHost, s));
    }
    if (t->col_major) {
        h->csc_tile_ptr = (int *)malloc(((size_t)t->tilen + 1) * 4);
        h->csc_tile_rowidx = (int *)malloc((nt ? nt : 1) * 4);
        if (!h->csc_tile_ptr || !h->csc_tile_rowidx) { set_error(TSG_ERR_NOMEM, "tile download: host allocation failed"); return g_err; }
        CK(cudaMemcpyAsync(h->csc_tile_ptr, t->csc_tile_ptr, ((size_t)t->tilen + 1) * 4, cudaMemcpyDeviceToHost, s));
        if (nt) CK(cudaMemcpyAsync(h->csc_tile_rowidx, t->csc_tile_rowidx, nt * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return TSG_OK;
}

int tsg_tilerow_weights(const tsg_dtile *a, const tsg_dtile *b, long long *w_host)
{
    if (ensure_init()) return g_err;
    if (a->n != b->m) { set_error(TSG_ERR_UNSUPPORTED, "weights: inner dimensions differ"); return g_err; }
    int *w = nullptr, *jlo = nullptr, *jhi = nullptr;
    int rc = tilerow_weights_device(a, b, &w, &jlo, &jhi);
    std::vector<int> hw((size_t)a->tilem + 1);
    int overflow = 0;
    if (!rc && a->tilem > 0 &&
        !cuda_ok(cudaMemcpyAsync(hw.data(), w, (size_t)a->tilem * 4, cudaMemcpyDeviceToHost, g_ctx.stream), "weights D2H", __FILE__, __LINE__))
        rc = g_err;
    if (!rc) rc = read_back_i32((int *)g_ctx.d_scalars + 1, &overflow);  // k_step1_weights raises scal[1] when a row's weight passes int32
    if (!rc && overflow) {
        set_error(TSG_ERR_OVERFLOW, "weights: a tile-row has more than 2^31-1 matched tile pairs");
        rc = g_err;
    }
    if (!rc)
        for (int i = 0; i < a->tilem; i++) w_host[i] = hw[i];
    dfree(w); dfree(jlo); dfree(jhi);
    return rc;
}

int tsg_spgemm(const tsg_dtile *a, const tsg_dtile *b, int trow0, int trow1, tsg_dtile *c, tsg_stats *stats)
{
    if (ensure_init()) return g_err;
    int rc = spgemm_device(a, b, trow0, trow1, c, stats);
    if (rc) { tsg_tile_free(c); return rc; }
    return TSG_OK;
}

int tsg_tile2csr(const tsg_dtile *t, tsg_dcsr *out)
{
    if (ensure_init()) return g_err;
    int rc = tile2csr_device(t, out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    return TSG_OK;
}

int tsg_tile_rowsums(const tsg_dtile *t, double *sums_host, long long *counts_host)
{
    if (ensure_init()) return g_err;
    const size_t m = (size_t)(t->m > 0 ? t->m : 1);
    double *d_s = dalloc_n<double>(m);
    long long *d_c = dalloc_n<long long>(m);
    int rc = (!d_s || !d_c) ? g_err : tile_rowsums_device(t, d_s, d_c);
    if (!rc && sums_host && t->m > 0 &&
        !cuda_ok(cudaMemcpyAsync(sums_host, d_s, (size_t)t->m * 8, cudaMemcpyDeviceToHost, g_ctx.stream), "rowsums D2H", __FILE__, __LINE__))
        rc = g_err;
    if (!rc && counts_host && t->m > 0 &&
        !cuda_ok(cudaMemcpyAsync(counts_host, d_c, (size_t)t->m * 8, cudaMemcpyDeviceToHost, g_ctx.stream), "rowcounts D2H", __FILE__, __LINE__))
        rc = g_err;
    if (!cuda_ok(cudaStreamSynchronize(g_ctx.stream), "rowsums sync", __FILE__, __LINE__)) rc = g_err;
    dfree(d_s); dfree(d_c);
    return rc;
}

int tsg_spgemm_csr_host(int m, int k, int n, const int *a_rowptr, const int *a_colidx, const double *a_val,
                        const int *b_rowptr, const int *b_colidx, const double *b_val, int aat,
                        int **c_rowptr, int **c_colidx, double **c_val, long long *c_nnz, tsg_stats *stats)
{
    if (ensure_init()) return g_err;
    tsg_dcsr A, B, Cc;
    tsg_dtile tA, tB, tC;
    memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B)); memset(&Cc, 0, sizeof(Cc));
    memset(&tA, 0, sizeof(tA)); memset(&tB, 0, sizeof(tB)); memset(&tC, 0, sizeof(tC));
    int rc = tsg_csr_upload(m, k, a_rowptr, a_colidx, a_val, &A);
    const tsg_dcsr *Bp = &A;
    if (!rc && aat) { rc = tsg_transpose(&A, &B); Bp = &B; }
    else if (!rc && b_rowptr) { rc = tsg_csr_upload(k, n, b_rowptr, b_colidx, b_val, &B); Bp = &B; }
    if (!rc && Bp->n != n) { set_error(TSG_ERR_UNSUPPORTED, "spgemm_csr_host: B has %d columns, expected %d", Bp->n, n); rc = g_err; }
    if (!rc) rc = tsg_csr2tile(&A, 0, &tA);
    if (!rc) rc = tsg_csr2tile(Bp, 1, &tB);
    if (!rc) rc = tsg_spgemm(&tA, &tB, 0, -1, &tC, stats);
    if (!rc) rc = tsg_tile2csr(&tC, &Cc);
    if (!rc) {
        *c_nnz = Cc.nnz;
        *c_rowptr = (int *)malloc(((size_t)Cc.m + 1) * 4);
        *c_colidx = (int *)malloc((size_t)(Cc.nnz > 0 ? Cc.nnz : 1) * 4);
        *c_val = (double *)malloc((size_t)(Cc.nnz > 0 ? Cc.nnz : 1) * 8);
        if (!*c_rowptr || !*c_colidx || !*c_val) { set_error(TSG_ERR_NOMEM, "spgemm_csr_host: host allocation failed"); rc = g_err; }
        else rc = tsg_csr_download(&Cc, *c_rowptr, *c_colidx, *c_val);
    }
    tsg_csr_free(&A); tsg_csr_free(&B); tsg_csr_free(&Cc);
    tsg_tile_free(&tA); tsg_tile_free(&tB); tsg_tile_free(&tC);
    return rc;
}

/* ---------------- end to end with overlap: C leaves the device slab by slab while the next slab is computed ---------------- */

// Cut tile-rows [r0, r1) into slabs of about equal step-1 weight (tile pairs). nslabs <= 0: up to 16 slabs of >= 2^20
// pairs each. No slab exceeds 2^28 pairs unless a single tile-row does (the bound spgemm_device's scratch is sized for).
static void plan_slabs(const long long *w, int r0, int r1, int nslabs, std::vector<int> &cuts)
{
    long long total = 0;
    for (int i = r0; i < r1; i++) total += w[i];
    long long S = nslabs > 0 ? nslabs : total >> 20;
    if (nslabs <= 0 && S > 16) S = 16;
    const long long max_pairs = 1ll << 28;
    if (S < (total + max_pairs - 1) / max_pairs) S = (total + max_pairs - 1) / max_pairs;
    if (S > r1 - r0) S = r1 - r0;  // more slabs than tile-rows means one per row
    if (S < 1) S = 1;
    cuts.clear();
    cuts.push_back(r0);
    long long acc = 0, done = 0, k = 1;  // slab k ends with the tile-row that takes the running total past k/S of the total
    for (int i = r0; i < r1; i++) {
        if (acc > 0 && acc + w[i] > max_pairs) { cuts.push_back(i); acc = 0; }  // hard bound (a heavier single row stands alone)
        acc += w[i];
        done += w[i];
        if (i + 1 < r1 && total > 0 && (__int128)done * S >= (__int128)k * total) {
            cuts.push_back(i + 1);
            acc = 0;
            k = (long long)((__int128)done * S / total) + 1;
        }
    }
    if (r1 > r0) cuts.push_back(r1);
}

int tsg_plan_slabs(const long long *weights, int trow0, int trow1, int nslabs, int *cuts, int cuts_cap)
{
    if (!weights || !cuts || trow0 < 0 || trow1 < trow0) { set_error(TSG_ERR_INPUT, "plan_slabs: bad arguments"); return -1; }
    std::vector<int> c;
    plan_slabs(weights, trow0, trow1, nslabs, c);
    if ((int)c.size() > cuts_cap) { set_error(TSG_ERR_NOMEM, "plan_slabs: %zu boundaries, room for %d", c.size(), cuts_cap); return -1; }
    for (size_t i = 0; i < c.size(); i++) cuts[i] = c[i];
    return (int)c.size() - 1;
}

/* ---------------- slab runner: C tile-rows in slabs of bounded size, each handed to a callback ---------------- */

int tsg_spgemm_slabs(const tsg_dtile *a, const tsg_dtile *b, int trow0, int trow1, long long max_pairs, tsg_slab_sink sink, void *user,
                     tsg_stats *totals, int *nslabs_out)
{
    if (ensure_init()) return g_err;
    if (trow1 < 0) trow1 = a->tilem;
    if (trow0 < 0 || trow0 > trow1 || trow1 > a->tilem) {
        set_error(TSG_ERR_INPUT, "spgemm_slabs: tile-row range [%d,%d) outside [0,%d]", trow0, trow1, a->tilem);
        return g_err;
    }
    if (a->n != b->m) { set_error(TSG_ERR_UNSUPPORTED, "spgemm_slabs: inner dimensions differ (%d vs %d)", a->n, b->m); return g_err; }
    if (max_pairs <= 0 || max_pairs > (1ll << 30)) max_pairs = 1ll << 28;  // what spgemm_device's 32-bit offsets are sized for
    if (totals) memset(totals, 0, sizeof(*totals));
    if (nslabs_out) *nslabs_out = 0;
    std::vector<long long> w((size_t)a->tilem + 1);
    int rc = tsg_tilerow_weights(a, b, w.data());
    if (rc) return rc;
    int nslabs = 0;
    for (int r0 = trow0; r0 < trow1 && !rc;) {
        long long acc = 0;
        int r1 = r0;
        while (r1 < trow1 && (r1 == r0 || acc + w[r1] <= max_pairs)) acc += w[r1++];  // a heavier single row stands alone
        tsg_dtile tC;
        tsg_stats st;
        memset(&tC, 0, sizeof(tC));
        rc = spgemm_device(a, b, r0, r1, &tC, &st);
        if (!rc && totals) {
            totals->ms_step1 += st.ms_step1; totals->ms_step2 += st.ms_step2; totals->ms_step3 += st.ms_step3;
            totals->ms_alloc += st.ms_alloc; totals->ms_total += st.ms_total; totals->numblkC += st.numblkC;
            totals->nnzC += st.nnzC; totals->pairs += st.pairs; totals->launches += st.launches;
            totals->algorithmic_bytes += st.algorithmic_bytes;
            totals->rows_staged += st.rows_staged; totals->rows_gather += st.rows_gather; totals->tiles_dense += st.tiles_dense;
            totals->tiles_nonempty += st.tiles_nonempty;
            if (st.rows_smem > totals->rows_smem) totals->rows_smem = st.rows_smem;
            if (st.plan_recipes > totals->plan_recipes) totals->plan_recipes = st.plan_recipes;
            if (st.row_templates > totals->row_templates) totals->row_templates = st.row_templates;
        }
        if (!rc && sink && sink(&tC, &st, user)) {
            set_error(TSG_ERR_INPUT, "spgemm_slabs: the slab callback stopped the run at tile-rows [%d,%d)", r0, r1);
            rc = g_err;
        }
        tsg_tile_free(&tC);
        nslabs++;
        r0 = r1;
    }
    if (nslabs_out) *nslabs_out = nslabs;
    return rc;
}

static int landing_init()
{
    Ctx &c = g_ctx;
    if (c.copy_stream) return TSG_OK;
    CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreateWithFlags(&c.land[k].ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.land[k].copied, cudaEventDisableTiming));
    }
    return TSG_OK;
}

int tsg_spgemm_to_host(const tsg_dtile *a, const tsg_dtile *b, int trow0, int trow1, int nslabs, int *c_rowptr, int *c_colidx,
                       double *c_val, long long c_cap, long long *c_nnz, tsg_stats *stats)
{
    if (ensure_init()) return g_err;
    Ctx &c = g_ctx;
    if (trow1 < 0) trow1 = a->tilem;
    if (trow0 < 0 || trow0 > trow1 || trow1 > a->tilem) {
        set_error(TSG_ERR_INPUT, "spgemm_to_host: tile-row range [%d,%d) outside [0,%d]", trow0, trow1, a->tilem);
        return g_err;
    }
    if (a->n != b->m) { set_error(TSG_ERR_UNSUPPORTED, "spgemm_to_host: inner dimensions differ (%d vs %d)", a->n, b->m); return g_err; }
    if (stats) memset(stats, 0, sizeof(*stats));
    if (c_nnz) *c_nnz = 0;
    c_rowptr[0] = 0;
    if (trow0 == trow1) return TSG_OK;
    int rc = landing_init();
    if (rc) return rc;
    std::vector<long long> w((size_t)a->tilem + 1);
    rc = tsg_tilerow_weights(a, b, w.data());
    if (rc) return rc;
    std::vector<int> cuts;
    plan_slabs(w.data(), trow0, trow1, nslabs, cuts);

    long long off = 0;
    bool full = false;
    for (size_t s = 0; s + 1 < cuts.size() && !rc; s++) {
        const int r0 = cuts[s], r1 = cuts[s + 1];
        tsg_dtile tC;
        tsg_stats st;
        memset(&tC, 0, sizeof(tC));
        rc = spgemm_device(a, b, r0, r1, &tC, &st);
        if (rc) { tsg_tile_free(&tC); break; }
        if (stats) {
            stats->ms_step1 += st.ms_step1; stats->ms_step2 += st.ms_step2; stats->ms_step3 += st.ms_step3;
            stats->ms_alloc += st.ms_alloc; stats->ms_total += st.ms_total; stats->numblkC += st.numblkC;
            stats->nnzC += st.nnzC; stats->pairs += st.pairs; stats->launches += st.launches;
            stats->algorithmic_bytes += st.algorithmic_bytes;
            stats->rows_staged += st.rows_staged; stats->rows_gather += st.rows_gather; stats->tiles_dense += st.tiles_dense;
            stats->tiles_nonempty += st.tiles_nonempty;
            if (st.rows_smem > stats->rows_smem) stats->rows_smem = st.rows_smem;
            if (st.plan_recipes > stats->plan_recipes) stats->plan_recipes = st.plan_recipes;
            if (st.row_templates > stats->row_templates) stats->row_templates = st.row_templates;
        }
        const long long nz = tC.nnz;
        const int rows = tC.m;
        if (off + nz >= (1ll << 31)) {
            set_error(TSG_ERR_OVERFLOW, "spgemm_to_host: nnz(C) passes 2^31 in tile-rows [%d,%d); 32-bit CSR row pointers cannot hold it", r0, r1);
            rc = g_err;
        } else if (off + nz > c_cap) {
            full = true;  // keep counting so that the caller learns the capacity it needs
        } else if (!full) {
            Ctx::Landing &L = c.land[s & 1];
            const size_t nzs = (size_t)(nz > 0 ? nz : 1);
            const size_t o_ci = (((size_t)rows + 1) * 4 + 255) & ~(size_t)255, o_v = o_ci + ((nzs * 4 + 255) & ~(size_t)255);
            const size_t need = o_v + nzs * 8;
            if (L.cap < need) {  // grow: the copy that still reads the old buffer has to finish first
                if (L.busy && !cuda_ok(cudaEventSynchronize(L.copied), "landing wait", __FILE__, __LINE__)) rc = g_err;
                if (L.p) dfree(L.p);
                L.cap = 0; L.busy = false;
                L.p = rc ? nullptr : (char *)dalloc(need + need / 8);
                if (L.p) L.cap = need + need / 8; else rc = g_err;
            } else if (L.busy && !cuda_ok(cudaStreamWaitEvent(c.stream, L.copied, 0), "landing wait", __FILE__, __LINE__)) {
                rc = g_err;
            }
            if (!rc) rc = tile2csr_into(&tC, (int *)L.p, (int *)(L.p + o_ci), (double *)(L.p + o_v), (int)off);
            if (!rc) {
                const size_t row_at = (size_t)(r0 - trow0) * TS;
                bool ok = cuda_ok(cudaEventRecord(L.ready, c.stream), "landing ready", __FILE__, __LINE__) &&
                          cuda_ok(cudaStreamWaitEvent(c.copy_stream, L.ready, 0), "landing ready", __FILE__, __LINE__) &&
                          cuda_ok(cudaMemcpyAsync(c_rowptr + row_at, L.p, ((size_t)rows + 1) * 4, cudaMemcpyDeviceToHost, c.copy_stream),
                                  "landing rowptr", __FILE__, __LINE__);
                if (ok && nz > 0)
                    ok = cuda_ok(cudaMemcpyAsync(c_colidx + off, L.p + o_ci, (size_t)nz * 4, cudaMemcpyDeviceToHost, c.copy_stream),
                                 "landing colidx", __FILE__, __LINE__) &&
                         cuda_ok(cudaMemcpyAsync(c_val + off, L.p + o_v, (size_t)nz * 8, cudaMemcpyDeviceToHost, c.copy_stream),
                                 "landing val", __FILE__, __LINE__);
                ok = ok && cuda_ok(cudaEventRecord(L.copied, c.copy_stream), "landing copied", __FILE__, __LINE__);
                L.busy = ok;
                if (!ok) rc = g_err;
            }
        }
        tsg_tile_free(&tC);
        off += nz;
    }
    cudaError_t e = cudaStreamSynchronize(c.copy_stream);  // drain on every path: the landing buffers are reused by the next call
    c.land[0].busy = c.land[1].busy = false;
    if (!rc && !cuda_ok(e, "landing drain", __FILE__, __LINE__)) rc = g_err;
    if (rc) return rc;
    if (c_nnz) *c_nnz = off;
    if (full) {
        set_error(TSG_ERR_NOMEM, "spgemm_to_host: C has %lld entries, the output buffers hold %lld", off, c_cap);
        return g_err;
    }
    return TSG_OK;
}

int tsg_spgemm_csr_host_into(int m, int k, int n, const int *a_rowptr, const int *a_colidx, const double *a_val,
                             const int *b_rowptr, const int *b_colidx, const double *b_val, int aat, int *c_rowptr,
                             int *c_colidx, double *c_val, long long c_cap, long long *c_nnz, tsg_stats *stats)
{
    if (ensure_init()) return g_err;
    tsg_dcsr A, B;
    tsg_dtile tA, tB;
    memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B)); memset(&tA, 0, sizeof(tA)); memset(&tB, 0, sizeof(tB));
    int rc = tsg_csr_upload(m, k, a_rowptr, a_colidx, a_val, &A);
    const tsg_dcsr *Bp = &A;
    if (!rc && aat) { rc = tsg_transpose(&A, &B); Bp = &B; }
    else if (!rc && b_rowptr) { rc = tsg_csr_upload(k, n, b_rowptr, b_colidx, b_val, &B); Bp = &B; }
    if (!rc && Bp->n != n) { set_error(TSG_ERR_UNSUPPORTED, "spgemm_csr_host_into: B has %d columns, expected %d", Bp->n, n); rc = g_err; }
    if (!rc) rc = tsg_csr2tile(&A, 0, &tA);
    if (!rc) rc = tsg_csr2tile(Bp, 1, &tB);
    if (!rc) rc = tsg_spgemm_to_host(&tA, &tB, 0, -1, 0, c_rowptr, c_colidx, c_val, c_cap, c_nnz, stats);
    tsg_csr_free(&A); tsg_csr_free(&B);
    tsg_tile_free(&tA); tsg_tile_free(&tB);
    return rc;
}

/* ------------------------------- drop-in entry points ------------------------------- */

// general tile sizes: tiles of A are tm x tn, tiles of B tn x tm (reference src/main.cu:84, src/csr2tile.h:282-283)
static void csr2tile_host_general(SMatrix *mat, int tm, int tn, int col_major)
{
    const int TR = col_major ? tn : tm, TC = col_major ? tm : tn;
    tsg_dcsr A, An;
    tsg_gtile t;
    memset(&t, 0, sizeof(t)); memset(&An, 0, sizeof(An));
    if (tsg_csr_upload(mat->m, mat->n, mat->rowpointer, mat->columnindex, mat->value, &A)) return;
    int rc = gtile_csr2tile_device(&A, col_major, TR, TC, &t);
    if (rc == TSG_ERR_INPUT && gtile_last_input_flags() == 2) {  // unsorted rows / duplicates: canonicalise and retry, as for 16 x 16
        tilespgemm_clear_error();
        gtile_free(&t);
        const char *pol = getenv("TSG_DUP_POLICY");
        rc = tsg_csr_canonicalize(&A, pol && !strcmp(pol, "first") ? 1 : 0, &An);
        if (!rc) rc = gtile_csr2tile_device(&An, col_major, TR, TC, &t);
    }
    if (rc == TSG_OK) gtile_download(&t, mat);
    gtile_free(&t);
    tsg_csr_free(&A);
    tsg_csr_free(&An);
}

static void csr2tile_host(SMatrix *mat, int tm, int tn, int col_major)
{
    if (ensure_init()) return;
    if (!tiles_16(tm, tn)) {
        if (tiles_general(tm, tn)) csr2tile_host_general(mat, tm, tn, col_major);
        return;
    }
    tsg_dcsr A, An;
    tsg_dtile t;
    memset(&t, 0, sizeof(t)); memset(&An, 0, sizeof(An));
    if (tsg_csr_upload(mat->m, mat->n, mat->rowpointer, mat->columnindex, mat->value, &A)) return;
    int rc = tsg_csr2tile(&A, col_major, &t);
    if (rc == TSG_ERR_INPUT && (last_input_flags() & 2) && !(last_input_flags() & 1)) {  // unsorted rows (which also derail the tile search: flag 4)
        // The reference's loader neither sorts rows nor merges duplicates (src/mmio_highlevel.h:593-759) and its csr2tile
        // takes whatever order it is given (src/csr2tile.h:152-168). The kernels here need sorted, duplicate-free rows:
        // bring the matrix into that form on the device (duplicates summed; TSG_DUP_POLICY=first keeps the first) and retry.
        // With duplicates the tiled matrix then holds fewer entries than mat->nnz says; mat->nnz is updated.
        static bool told = false;
        if (!told) { fprintf(stderr, "[tilespgemm_b200] note: CSR rows unsorted or with duplicates; canonicalised on the device\n"); told = true; }
        tilespgemm_clear_error();
        tsg_tile_free(&t);
        const char *pol = getenv("TSG_DUP_POLICY");
        rc = tsg_csr_canonicalize(&A, pol && !strcmp(pol, "first") ? 1 : 0, &An);
        if (!rc) rc = tsg_csr2tile(&An, col_major, &t);
    }
    if (rc == TSG_OK) {
        // tsg_tile_download overwrites the size fields with identical values and fills the tile arrays;
        // the CSR members of *mat (caller-owned, possibly aliased by B, src/main.cu:145-151) are untouched.
        tsg_tile_download(&t, mat);
    }
    tsg_tile_free(&t);
    tsg_csr_free(&A);
    tsg_csr_free(&An);
}

void csr2tile_row_major(SMatrix *matrix, int tile_size_m, int tile_size_n) { csr2tile_host(matrix, tile_size_m, tile_size_n, 0); }
void csr2tile_col_major(SMatrix *matrix, int tile_size_m, int tile_size_n) { csr2tile_host(matrix, tile_size_m, tile_size_n, 1); }

void tilespgemm(SMatrix *matrixA, SMatrix *matrixB, SMatrix *matrixC, unsigned int *blk_intersec_bitmask_A,
                unsigned int *blk_intersec_bitmask_B, int blk_intersec_bitmask_len, double densityA, double densityB,
                unsigned long long int nnzCub, unsigned long long int *nnzC_computed, double *compression_rate,
                double *time_tile, double *gflops_tile, char *filename, double *time_step1, double *time_step2,
                double *time_step3, double *time_malloc, int tile_size_m, int tile_size_n)
{
    (void)blk_intersec_bitmask_A; (void)blk_intersec_bitmask_B; (void)blk_intersec_bitmask_len;  // dense tile bitmaps: never needed
    (void)densityA; (void)densityB; (void)filename;
    if (ensure_init()) return;
    if (!tiles_16(tile_size_m, tile_size_n)) {
        if (!tiles_general(tile_size_m, tile_size_n)) return;
        tsg_gtile gA, gB, gC;
        tsg_stats st;
        memset(&gA, 0, sizeof(gA)); memset(&gB, 0, sizeof(gB)); memset(&gC, 0, sizeof(gC)); memset(&st, 0, sizeof(st));
        if (gtile_upload(matrixA, 0, tile_size_m, tile_size_n, &gA) == TSG_OK && gtile_upload(matrixB, 1, tile_size_n, tile_size_m, &gB) == TSG_OK &&
            gtile_spgemm_device(&gA, &gB, &gC, &st) == TSG_OK) {
            int keep_sym = matrixC->isSymmetric;
            if (gtile_download(&gC, matrixC) == TSG_OK) {
                matrixC->isSymmetric = keep_sym;
                if (nnzC_computed) *nnzC_computed = (unsigned long long)st.nnzC;
                if (compression_rate) *compression_rate = st.nnzC ? (double)nnzCub / (double)st.nnzC : 0.0;
                if (time_tile) *time_tile = st.ms_total;
                if (gflops_tile) *gflops_tile = st.ms_total > 0 ? 2.0 * (double)nnzCub / (st.ms_total * 1e6) : 0.0;
                if (time_step1) *time_step1 = st.ms_step1;
                if (time_step2) *time_step2 = st.ms_step2;
                if (time_step3) *time_step3 = st.ms_step3;
                if (time_malloc) *time_malloc = st.ms_alloc;
                printf("Non-empty tiles of C = %i\n", (int)st.numblkC);
                printf("nnzC = %i\n", (int)st.nnzC);
                printf("CUDA  TileSpGEMM runtime is %4.2f ms, gflops = %4.2f\n", st.ms_total,
                       st.ms_total > 0 ? 2.0 * (double)nnzCub / (st.ms_total * 1e6) : 0.0);
            }
        }
        gtile_free(&gA); gtile_free(&gB); gtile_free(&gC);
        return;
    }
    tsg_dtile tA, tB, tC;
    tsg_stats st;
    memset(&tA, 0, sizeof(tA)); memset(&tB, 0, sizeof(tB)); memset(&tC, 0, sizeof(tC)); memset(&st, 0, sizeof(st));
    if (tsg_tile_upload(matrixA, 0, &tA) == TSG_OK && tsg_tile_upload(matrixB, 1, &tB) == TSG_OK &&
        tsg_spgemm(&tA, &tB, 0, -1, &tC, &st) == TSG_OK) {
        int keep_sym = matrixC->isSymmetric;
        if (tsg_tile_download(&tC, matrixC) == TSG_OK) {
            matrixC->isSymmetric = keep_sym;
            matrixC->m = matrixA->m; matrixC->n = matrixB->n;           // reference :2768-2771
            matrixC->tilem = matrixA->tilem; matrixC->tilen = matrixB->tilen;
            if (nnzC_computed) *nnzC_computed = (unsigned long long)st.nnzC;
            if (compression_rate) *compression_rate = st.nnzC ? (double)nnzCub / (double)st.nnzC : 0.0;
            if (time_tile) *time_tile = st.ms_total;
            if (gflops_tile) *gflops_tile = st.ms_total > 0 ? 2.0 * (double)nnzCub / (st.ms_total * 1e6) : 0.0;
            if (time_step1) *time_step1 = st.ms_step1;
            if (time_step2) *time_step2 = st.ms_step2;
            if (time_step3) *time_step3 = st.ms_step3;
            if (time_malloc) *time_malloc = st.ms_alloc;
            // same stdout lines as the reference (:2810-2812)
            printf("Non-empty tiles of C = %i\n", (int)st.numblkC);
            printf("nnzC = %i\n", (int)st.nnzC);
            printf("CUDA  TileSpGEMM runtime is %4.2f ms, gflops = %4.2f\n", st.ms_total,
                   st.ms_total > 0 ? 2.0 * (double)nnzCub / (st.ms_total * 1e6) : 0.0);
        }
    }
    tsg_tile_free(&tA); tsg_tile_free(&tB); tsg_tile_free(&tC);
}

void tile2csr(SMatrix *matrix, int tile_size_m, int tile_size_n)
{
    if (ensure_init()) return;
    const bool general = !tiles_16(tile_size_m, tile_size_n);
    if (general && !tiles_general(tile_size_m, tile_size_n)) return;
    tsg_dtile t;
    tsg_gtile g;
    tsg_dcsr c;
    memset(&t, 0, sizeof(t)); memset(&c, 0, sizeof(c)); memset(&g, 0, sizeof(g));
    // general: tiles of tile_size_m rows x tile_size_n columns (the driver passes (tile_size_m, tile_size_m) for C, src/main.cu:327)
    if (general ? (gtile_upload(matrix, 0, tile_size_m, tile_size_n, &g) == TSG_OK && gtile_tile2csr_device(&g, &c) == TSG_OK)
                : (tsg_tile_upload(matrix, 0, &t) == TSG_OK && tsg_tile2csr(&t, &c) == TSG_OK)) {
        matrix->rowpointer = (int *)malloc(((size_t)c.m + 1) * 4);
        matrix->columnindex = (int *)malloc((size_t)(c.nnz > 0 ? c.nnz : 1) * 4);
        matrix->value = (double *)malloc((size_t)(c.nnz > 0 ? c.nnz : 1) * 8);
        if (!matrix->rowpointer || !matrix->columnindex || !matrix->value) set_error(TSG_ERR_NOMEM, "tile2csr: host allocation failed");
        else if (tsg_csr_download(&c, matrix->rowpointer, matrix->columnindex, matrix->value) == TSG_OK)
            matrix->nnz = (int)c.nnz;  // reference src/tile2csr.h:103-104
    }
    tsg_tile_free(&t);
    gtile_free(&g);
    tsg_csr_free(&c);
}

/* ------------------------------- general tile sizes (Part 3) ------------------------------- */
int tsg_gtile_size_ok(int tile_rows, int tile_cols) { return gtile_size_ok(tile_rows, tile_cols) ? 1 : 0; }

int tsg_gtile_csr2tile(const tsg_dcsr *a, int col_major, int tile_rows, int tile_cols, tsg_gtile *out)
{
    if (ensure_init()) return g_err;
    return gtile_csr2tile_device(a, col_major, tile_rows, tile_cols, out);
}

int tsg_gtile_upload(const SMatrix *host, int col_major, int tile_rows, int tile_cols, tsg_gtile *out)
{
    if (ensure_init()) return g_err;
    return gtile_upload(host, col_major, tile_rows, tile_cols, out);
}

int tsg_gtile_download(const tsg_gtile *t, SMatrix *host)
{
    if (ensure_init()) return g_err;
    return gtile_download(t, host);
}

void tsg_gtile_free(tsg_gtile *t)
{
    if (!g_ready) { memset(t, 0, sizeof(*t)); return; }
    gtile_free(t);
}

int tsg_gtile_spgemm(const tsg_gtile *a, const tsg_gtile *b, tsg_gtile *c, tsg_stats *stats)
{
    if (ensure_init()) return g_err;
    return gtile_spgemm_device(a, b, c, stats);
}

int tsg_gtile_tile2csr(const tsg_gtile *t, tsg_dcsr *out)
{
    if (ensure_init()) return g_err;
    return gtile_tile2csr_device(t, out);
}

void matrix_destroy(SMatrix *matrix)
{
    free(matrix->tile_ptr);
    free(matrix->tile_columnidx);
    free(matrix->tile_nnz);
    free(matrix->tile_csr_Value);
    free(matrix->tile_csr_Col);
    free(matrix->tile_csr_Ptr);
    free(matrix->mask);
    matrix->tile_ptr = nullptr; matrix->tile_columnidx = nullptr; matrix->tile_nnz = nullptr;
    matrix->tile_csr_Value = nullptr; matrix->tile_csr_Col = nullptr; matrix->tile_csr_Ptr = nullptr; matrix->mask = nullptr;
}

void matrix_transposition(const int m, const int n, const MAT_PTR_TYPE nnz, const MAT_PTR_TYPE *csrRowPtr, const int *csrColIdx,
                          const MAT_VAL_TYPE *csrVal, int *cscRowIdx, MAT_PTR_TYPE *cscColPtr, MAT_VAL_TYPE *cscVal)
{
    (void)nnz;
    if (ensure_init()) return;
    tsg_dcsr A, AT;
    memset(&AT, 0, sizeof(AT));
    if (tsg_csr_upload(m, n, csrRowPtr, csrColIdx, csrVal, &A)) return;
    if (tsg_transpose(&A, &AT) == TSG_OK) tsg_csr_download(&AT, cscColPtr, cscRowIdx, cscVal);
    tsg_csr_free(&A);
    tsg_csr_free(&AT);
}

}  // extern "C"
