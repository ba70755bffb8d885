// sdfb_api.cu -- host side of the C ABI declared in include/sdfb.h.
//
// Mirrors the host orchestration the reference does in gpu_lib/makelevelset3_gpu.cu:595-777
// (allocate, H2D mesh, kernels, D2H phi, free) but: state lives in a reusable plan, everything is
// enqueued asynchronously on one stream, errors come back as codes instead of exit(EXIT_FAILURE)
// (gpu_lib/makelevelset3_gpu.cu:14-20), linear indices are 64-bit, and there is no debug D2H copy.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <algorithm>
#include <atomic>
#include <new>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#include <unistd.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <cuda_runtime.h>
#include "../../include/sdfb.h"
#include "sdfb_kernels.cuh"

using namespace sdfb;

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? SDFB_ERR_OOM : SDFB_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int env_int(const char *name, int dflt) { const char *v = getenv(name); return v && *v ? atoi(v) : dflt; }

}  // namespace

namespace sdfb {
// the development knobs of DESIGN.md section 4.4, read once per plan
Tuning tuning_from_env()
{
    Tuning t;
    t.relax_from = env_int("SDFB_RELAX_FROM", -1);
    t.fuse_pass = env_int("SDFB_FUSE_PASS", -1);
    t.minb = env_int("SDFB_MINB", 0);
    t.max_occ = env_int("SDFB_MAX_OCC", 0);
    t.cta_queue = env_int("SDFB_CTA_QUEUE", -1);
    t.cta_queue_until = env_int("SDFB_CTA_QUEUE_UNTIL", 8);
    t.relax_list_cap = env_int("SDFB_RELAX_LIST_CAP", 0);
    if (getenv("SDFB_RELAX_HEAVY_LIMIT")) t.relax_heavy_limit = (long long)strtoull(getenv("SDFB_RELAX_HEAVY_LIMIT"), nullptr, 10);
    t.relax_scan_from = env_int("SDFB_RELAX_SCAN_FROM", 13);
    t.relax_debug = env_int("SDFB_RELAX_DEBUG", 0);
    t.lookahead = env_int("SDFB_LOOKAHEAD", 1);
    t.look_cap = env_int("SDFB_LOOK_CAP", 0);
    t.early_copy = env_int("SDFB_EARLY_COPY", 1);
    t.col_shape = env_int("SDFB_COL_SHAPE", 0);
    t.link_timeout_s = env_int("SDFB_LINK_TIMEOUT_S", 20);
    t.order_w = env_int("SDFB_ORDER_W", -1);
    t.link_debug = env_int("SDFB_LINK_DEBUG", 0);
    t.link_trace = env_int("SDFB_LINK_TRACE", 0);
    return t;
}
}  // namespace sdfb

namespace {

bool device_is_sm100(int dev)
{
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    return major == 10;    // the only code in libsdfb.so is sm_100a SASS
}

// ---- host-side helpers for large PAGEABLE outputs (the drop-in call hands us plain malloc'ed memory) ------------
// cudaMemcpy into untouched pageable memory runs at 4-5 GB/s (first-touch page faults + a single-threaded bounce
// copy): 110-130 ms for the 537 MB phi of a 512^3 grid, more than the kernels.  So (1) the destination pages are
// touched by a few threads while the GPU is still computing, and (2) the copy is staged through a pinned ring and
// spread over the same threads.
int host_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return (int)(n == 0 ? 4 : (n > 12 ? 12 : n));
}

// A few persistent host threads for the two memory-bound jobs around a large download (touching the destination pages,
// copying from the pinned ring into pageable memory).  Spawning and joining a set of std::threads per 32 MB chunk cost
// 0.1-0.2 ms each time, a sixth of the chunk's copy.  The pool is created on first use and never destroyed (its threads
// sleep on a condition variable; joining them from a static destructor of a shared library at exit is asking for trouble).
// One job at a time: callers on several host threads (batch workers, the slabs of a multi-GPU call) take turns chunk by chunk.
class HostPool {
public:
    static HostPool &get() { static HostPool *p = new HostPool(); return *p; }
    void run(char *dst, size_t bytes, void (*fn)(char *, size_t, const char *), const char *src)
    {
        if (bytes == 0) return;
        std::lock_guard<std::mutex> job(job_mtx_);
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = dst; bytes_ = bytes; fn_ = fn; src_ = src;
            per_ = ((bytes + nt_ - 1) / nt_ + 4095) & ~(size_t)4095;
            remaining_ = nt_;
            ++gen_;
        }
        cv_work_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return remaining_ == 0; });
    }
private:
    HostPool() : nt_(host_threads())
    {
        for (int t = 0; t < nt_; ++t) std::thread([this, t] { worker(t); }).detach();
    }
    void worker(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            char *dst; size_t bytes, per; void (*fn)(char *, size_t, const char *); const char *src;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                dst = dst_; bytes = bytes_; per = per_; fn = fn_; src = src_;
            }
            const size_t lo = (size_t)t * per;
            if (lo < bytes) fn(dst + lo, bytes - lo < per ? bytes - lo : per, src ? src + lo : nullptr);
            bool last;
            { std::lock_guard<std::mutex> lk(m_); last = --remaining_ == 0; }
            if (last) cv_done_.notify_all();
        }
    }
    const int nt_;
    std::mutex job_mtx_, m_;
    std::condition_variable cv_work_, cv_done_;
    unsigned long long gen_ = 0;
    int remaining_ = 0;
    char *dst_ = nullptr; size_t bytes_ = 0, per_ = 0; void (*fn_)(char *, size_t, const char *) = nullptr; const char *src_ = nullptr;
};

void parallel_for_bytes(char *dst, size_t bytes, void (*fn)(char *, size_t, const char *), const char *src)
{
    HostPool::get().run(dst, bytes, fn, src);
}

void touch_pages(char *dst, size_t n, const char *) { for (size_t o = 0; o < n; o += 4096) reinterpret_cast<volatile char *>(dst)[o] = 0; }
// pinned ring -> the caller's pageable array.  The destination is written once and not read here: streaming stores skip the
// read-for-ownership of every destination line (a third of the DRAM traffic of this copy, which is what bounds it).
void copy_bytes(char *dst, size_t n, const char *src)
{
#if defined(__x86_64__) || defined(_M_X64)
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (n >= 4096 && ((reinterpret_cast<uintptr_t>(src) + head) & 15) == 0) {
        memcpy(dst, src, head);
        dst += head; src += head; n -= head;
        const size_t blocks = n / 64;
        const __m128i *s = reinterpret_cast<const __m128i *>(src);
        __m128i *d = reinterpret_cast<__m128i *>(dst);
        for (size_t b = 0; b < blocks; ++b, s += 4, d += 4) {
            const __m128i x0 = _mm_load_si128(s), x1 = _mm_load_si128(s + 1), x2 = _mm_load_si128(s + 2), x3 = _mm_load_si128(s + 3);
            _mm_stream_si128(d, x0); _mm_stream_si128(d + 1, x1); _mm_stream_si128(d + 2, x2); _mm_stream_si128(d + 3, x3);
        }
        _mm_sfence();
        memcpy(dst + blocks * 64, src + blocks * 64, n - blocks * 64);
        return;
    }
#endif
    memcpy(dst, src, n);
}

bool is_pageable(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Device -> host through two pinned staging buffers: the D2H copy of chunk i+1 overlaps what the sink does with chunk i
// (a threaded memcpy into pageable memory, or a write to a file).  sink(data, n, offset) returns false to stop.
// Returns a CUDA error code; *sink_ok tells whether every sink call succeeded.
// a staging ring is allocated once per DEVICE (pinning 64 MB costs several ms) and shared by that device's plans under
// a lock: its events belong to the device's context (recording an event on another device's stream is
// cudaErrorInvalidResourceHandle), and slabs on different GPUs download concurrently
constexpr int STAGE_BUFS = 4;                               // pinned buffers in the ring: up to three copies in flight while one is drained
constexpr size_t STAGE_CHUNK = (size_t)16 << 20;
struct StagingRing {
    std::mutex mtx;
    char *ring[STAGE_BUFS] = {nullptr};
    cudaEvent_t ev[STAGE_BUFS] = {nullptr};
} g_staging_dev[64];

template <class Sink>
cudaError_t staged_d2h_sink(const void *src, size_t bytes, cudaStream_t st, Sink sink, bool *sink_ok)
{
    const size_t CH = STAGE_CHUNK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);                     // the caller holds a DeviceGuard for the plan's device
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    StagingRing &g_staging = g_staging_dev[dev];
    std::lock_guard<std::mutex> lock(g_staging.mtx);
    char **ring = g_staging.ring;
    cudaEvent_t *ev = g_staging.ev;
    bool ok = true;
    for (int b = 0; b < STAGE_BUFS && e == cudaSuccess; ++b) {
        if (!ring[b]) e = cudaHostAlloc(reinterpret_cast<void **>(&ring[b]), CH, cudaHostAllocPortable);
        if (e == cudaSuccess && !ev[b]) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
    }
    const size_t nch = (bytes + CH - 1) / CH;
    auto chunk = [&](size_t i) { return bytes - i * CH < CH ? bytes - i * CH : CH; };
    auto issue = [&](size_t i) {                            // device -> ring[i % STAGE_BUFS], then the event that says so
        cudaError_t r = cudaMemcpyAsync(ring[i % STAGE_BUFS], static_cast<const char *>(src) + i * CH, chunk(i), cudaMemcpyDeviceToHost, st);
        if (r == cudaSuccess) r = cudaEventRecord(ev[i % STAGE_BUFS], st);
        return r;
    };
    for (size_t i = 0; i + 1 < (size_t)STAGE_BUFS && i < nch && e == cudaSuccess; ++i) e = issue(i);
    for (size_t i = 0; i < nch && e == cudaSuccess && ok; ++i) {
        // the buffer of chunk i + STAGE_BUFS - 1 held chunk i - 1, which the previous iteration drained
        if (i + STAGE_BUFS - 1 < nch) e = issue(i + STAGE_BUFS - 1);
        if (e == cudaSuccess) e = cudaEventSynchronize(ev[i % STAGE_BUFS]);
        if (e == cudaSuccess) ok = sink(ring[i % STAGE_BUFS], chunk(i), i * CH);
    }
    if (e != cudaSuccess || !ok) cudaStreamSynchronize(st); // nothing may still be writing into the ring when the lock goes
    if (sink_ok) *sink_ok = ok;
    return e;
}

// Host -> device for a pageable source, through the same ring: the worker threads copy chunk i into a pinned buffer while
// the DMA of chunk i-1 runs.  cudaMemcpyAsync from pageable memory lets the driver do this with one thread (24 MB of mesh:
// 2.3 ms; this way 0.8 ms).
cudaError_t staged_h2d(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    const size_t CH = STAGE_CHUNK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    StagingRing &g_staging = g_staging_dev[dev];
    std::lock_guard<std::mutex> lock(g_staging.mtx);
    char **ring = g_staging.ring;
    cudaEvent_t *ev = g_staging.ev;
    for (int b = 0; b < STAGE_BUFS && e == cudaSuccess; ++b) {
        if (!ring[b]) e = cudaHostAlloc(reinterpret_cast<void **>(&ring[b]), CH, cudaHostAllocPortable);
        if (e == cudaSuccess && !ev[b]) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
    }
    const size_t nch = (bytes + CH - 1) / CH;
    for (size_t i = 0; i < nch && e == cudaSuccess; ++i) {
        const size_t n = bytes - i * CH < CH ? bytes - i * CH : CH;
        const int b = (int)(i % STAGE_BUFS);
        if (i >= (size_t)STAGE_BUFS) e = cudaEventSynchronize(ev[b]);          // the DMA that last read this buffer
        if (e != cudaSuccess) break;
        parallel_for_bytes(ring[b], n, copy_bytes, static_cast<const char *>(src) + i * CH);
        e = cudaMemcpyAsync(static_cast<char *>(dst) + i * CH, ring[b], n, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
    }
    // the ring is shared: nothing may still be reading it when the lock goes (the copies are short; the kernels behind them are not waited for)
    for (int b = 0; b < STAGE_BUFS && b < (int)nch; ++b) { const cudaError_t w = cudaEventSynchronize(ev[b]); if (e == cudaSuccess) e = w; }
    return e;
}

cudaError_t staged_d2h(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    return staged_d2h_sink(src, bytes, st, [dst](const char *data, size_t n, size_t off) {
        parallel_for_bytes(static_cast<char *>(dst) + off, n, copy_bytes, data);
        return true;
    }, nullptr);
}

// ---- device memory: a library-private stream-ordered pool per device ---------------------------------------------
// cudaMalloc/cudaFree cost 30-900 ms per one-shot call on the B200 boxes (a dozen buffers mapped and unmapped every
// call; 128^3: 40-100 ms per call around 10 ms of kernels, with stalls up to 0.9 s).  All plan memory therefore comes
// from a cudaMemPool of our own that keeps freed blocks (up to SDFB_POOL_RETAIN_MB, default 16 GiB, above which the
// driver releases at the next synchronisation); sdfb_trim_memory() hands everything back.  Allocations are made on
// the pool's own stream and that stream is synchronised at once (it never carries work), so the memory may be used
// on any stream; frees are enqueued on it after the plan's work has completed (plan_quiesce).
struct DevicePool {
    cudaMemPool_t pool = nullptr;
    cudaStream_t st = nullptr;
};
DevicePool g_pools[64];
std::mutex g_pool_mtx;

cudaError_t pool_for_current_device(DevicePool **out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_pool_mtx);
    DevicePool &dp = g_pools[dev];
    if (!dp.pool) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if ((e = cudaMemPoolCreate(&pool, &props)) != cudaSuccess) return e;
        unsigned long long retain = (unsigned long long)16 << 30;
        if (getenv("SDFB_POOL_RETAIN_MB")) retain = strtoull(getenv("SDFB_POOL_RETAIN_MB"), nullptr, 10) << 20;
        if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &retain)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&dp.st, cudaStreamNonBlocking)) != cudaSuccess) {
            cudaMemPoolDestroy(pool);
            return e;
        }
        dp.pool = pool;
    }
    *out = &dp;
    return cudaSuccess;
}

template <class T>
cudaError_t dev_alloc(T **ptr, size_t bytes)
{
    DevicePool *dp = nullptr;
    cudaError_t e = pool_for_current_device(&dp);
    if (e != cudaSuccess) return e;
    void *v = nullptr;
    if ((e = cudaMallocFromPoolAsync(&v, bytes ? bytes : 1, dp->pool, dp->st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(dp->st)) != cudaSuccess) return e;
    *ptr = static_cast<T *>(v);
    return cudaSuccess;
}

// the caller has made sure nothing on the device still uses the block (plan_quiesce)
void dev_free(void *ptr)
{
    if (!ptr) return;
    DevicePool *dp = nullptr;
    if (pool_for_current_device(&dp) == cudaSuccess) cudaFreeAsync(ptr, dp->st);
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct sdfb_plan {
    int device = 0;
    uint32_t flags = 0;
    Grid g{};
    float init_phi = 0.f;
    // device state
    uint64_t *cells = nullptr;       // (nkl+2) planes
    int32_t *counts = nullptr;       // slab voxels
    float *phi = nullptr;            // slab voxels, i fastest
    float *phi_k = nullptr;          // slab voxels, k fastest (only with SDFB_OUT_KFASTEST)
    int32_t *scratch = nullptr;      // slab voxels, lazily allocated for tri / count downloads
    unsigned long long *changed = nullptr;
    uint32_t *progress = nullptr;    // column-schedule flags
    size_t progress_words = 0;
    uint32_t epoch = 0;              // column-schedule launch counter since the flags were last zeroed
    void *relax = nullptr;           // scratch of the relaxation schedule (lazily allocated, zeroed once)
    int last_sweep = -1;             // highest sweep index run since the last band: the relaxation schedule tells the
                                     // cells it changed by this sweep's stamp, so an index must not repeat
    int look_next = -1, look_hi = -1;   // lookahead window of the relaxation schedule: the next sweep it is valid for, and its end
    int look_lo = -1;                   // ... and its first sweep
    // mesh
    uint64_t ntri = 0, nvert = 0;
    uint32_t *tri_own = nullptr;     // owned copies when the mesh came from the host
    float *xyz_own = nullptr;
    uint64_t tri_own_cap = 0, xyz_own_cap = 0;
    TriRec *rec = nullptr;
    uint32_t *units = nullptr;
    TriExt *ext = nullptr;
    uint64_t *prefix = nullptr, *block_sums = nullptr;
    uint64_t rec_cap = 0;
    bool have_mesh = false, have_band = false, have_sign = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start, after band, after sweeps, after sign
    cudaEvent_t ev_copy = nullptr;   // end of the last asynchronous phi download
    bool copy_pending = false;
    bool timed = false;
    bool own_stream = false;         // the plan is only ever used on `stream` (one-shot and batch calls): quiescing waits for
    cudaStream_t stream = nullptr;   // that stream instead of the whole device
    int max_ctas = 0;                // batch mode: cap on the persistent sweep grids so that several plans share the SMs (0 = no cap)
    Tuning tun;                      // development knobs, read from the environment when the plan is created
    LinkState link;                  // exact multi-GPU mode: hand-over buffers shared with the neighbouring slabs' plans
};

namespace {

// cap on a persistent sweep grid when `plans` plans share the device: its share of the 3 CTAs an SM holds, but
// never less than half a CTA per SM (more plans than that simply queue)
int sm_share_ctas(int sms, int plans) { return (3 * sms) / plans > sms / 2 ? (3 * sms) / plans : sms / 2; }

// nothing on the device may still use the plan's buffers when they go back to the pool
void plan_quiesce(sdfb_plan *p)
{
    if (p->own_stream) cudaStreamSynchronize(p->stream); else cudaDeviceSynchronize();
    if (p->copy_pending && p->ev_copy) { cudaEventSynchronize(p->ev_copy); p->copy_pending = false; }
}

void free_mesh(sdfb_plan *p)
{
    plan_quiesce(p);
    dev_free(p->tri_own); dev_free(p->xyz_own); dev_free(p->rec); dev_free(p->units); dev_free(p->ext);
    dev_free(p->prefix); dev_free(p->block_sums);
    p->tri_own = nullptr; p->xyz_own = nullptr; p->rec = nullptr; p->units = nullptr; p->ext = nullptr;
    p->tri_own_cap = 0; p->xyz_own_cap = 0;
    p->prefix = nullptr; p->block_sums = nullptr; p->rec_cap = 0; p->have_mesh = false;
}

int ensure_mesh_capacity(sdfb_plan *p, uint64_t ntri)
{
    if (ntri <= p->rec_cap && p->rec) return SDFB_OK;
    if (p->rec) plan_quiesce(p);
    dev_free(p->rec); dev_free(p->units); dev_free(p->ext); dev_free(p->prefix); dev_free(p->block_sums);
    p->rec = nullptr; p->units = nullptr; p->ext = nullptr; p->prefix = nullptr; p->block_sums = nullptr; p->rec_cap = 0;
    uint64_t cap = ntri ? ntri : 1;
    CU(dev_alloc(&p->rec, cap * sizeof(TriRec)));
    CU(dev_alloc(&p->units, cap * sizeof(uint32_t)));
    CU(dev_alloc(&p->ext, cap * sizeof(TriExt)));
    CU(dev_alloc(&p->prefix, (cap + 1) * sizeof(uint64_t)));
    CU(dev_alloc(&p->block_sums, (cap / 2048 + 2) * sizeof(uint64_t)));
    p->rec_cap = cap;
    return SDFB_OK;
}

// The mesh named a vertex that does not exist.  The reference indexes x[] unchecked there (undefined behaviour,
// cpu_lib/makelevelset3.cpp:205; python/tests/test_sdfgen.py:826-847 accepts a crash or an exception); the device
// replaced the index by 0 instead of faulting, and the calls that deliver results refuse to.
int bad_index_error(const sdfb_plan *p, unsigned long long t)
{
    return fail(SDFB_ERR_INVALID, "triangle %llu names a vertex index >= %llu (the vertex count); no result delivered",
                t, (unsigned long long)p->nvert);
}

int build_records(sdfb_plan *p, const uint32_t *d_tri, const float *d_xyz, uint64_t ntri, uint64_t nvert, cudaStream_t st)
{
    if (ntri > SDFB_MAX_TRIANGLES) return fail(SDFB_ERR_LIMIT, "%llu triangles exceed the limit of %u", (unsigned long long)ntri, SDFB_MAX_TRIANGLES);
    int rc = ensure_mesh_capacity(p, ntri);
    if (rc) return rc;
    if (ntri && !nvert) return fail(SDFB_ERR_INVALID, "%llu triangles but no vertices", (unsigned long long)ntri);
    CU(cudaMemsetAsync(p->changed + 3, 0xff, sizeof(unsigned long long), st));     // lowest triangle with a bad vertex index
    g_launches += launch_tri_prep(d_tri, d_xyz, ntri, nvert, p->rec, p->changed + 3, st);
    CU(cudaGetLastError());
    // (An L2 persistence window over the records was measured: no effect at C2, and cudaDeviceSetLimit is a
    // device-wide, synchronising setting a library should not touch -- removed.)
    p->ntri = ntri; p->nvert = nvert; p->have_mesh = true; p->have_band = false; p->have_sign = false;
    return SDFB_OK;
}

// exact multi-GPU mode: unmap the neighbours' buffers and free our own (the caller made sure no neighbour still writes)
// drop the mappings of the neighbours' buffers (importer side); this plan's own inbound buffers stay allocated
void link_close_peers(sdfb_plan *p)
{
    for (int side = 0; side < 2; ++side) {
        if (p->link.peer_base[side]) cudaIpcCloseMemHandle(p->link.peer_base[side]);
        p->link.peer_base[side] = nullptr; p->link.peer_halo[side] = nullptr; p->link.peer_flags[side] = nullptr;
    }
    p->link.active = false;
    cudaGetLastError();
}

void link_release(sdfb_plan *p)
{
    link_close_peers(p);
    if (p->link.in_halo) cudaFree(p->link.in_halo);
    if (p->link.trace) cudaFree(p->link.trace);
    p->link = LinkState{};
    cudaGetLastError();
}

// what sdfb_plan_link_export hands to the neighbours (SDFB_LINK_HANDLE_BYTES)
struct LinkHandle {
    cudaIpcMemHandle_t mem;          // 64 bytes
    int32_t ni, nj, nk, k_lo, k_hi, device;
    uint64_t bytes;
    uint64_t local_ptr;              // the exporter's own pointer: valid for same-process links only
    int64_t pid;
};
static_assert(sizeof(LinkHandle) <= SDFB_LINK_HANDLE_BYTES, "LinkHandle must fit the public handle size");

int link_ensure_buffers(sdfb_plan *p)
{
    if (p->link.in_halo) return SDFB_OK;
    if (p->g.nkl() < 2 && (p->g.k_lo == 0 || p->g.k_hi == p->g.nk))
        return fail(SDFB_ERR_INVALID, "a linked slab on a grid face needs at least 2 planes (got [%d,%d))", p->g.k_lo, p->g.k_hi);
    const size_t plane = (size_t)p->g.plane();
    p->link.NJ = (int)link_flag_words_per_sweep(p->g);
    const size_t halo_bytes = (size_t)LINK_SWEEPS * plane * sizeof(uint64_t);
    const size_t bytes = halo_bytes + (size_t)LINK_SWEEPS * p->link.NJ * sizeof(unsigned long long);
    void *base = nullptr;
    // plain cudaMalloc, not the pool: the block is exported with cudaIpcGetMemHandle / opened by peer devices
    cudaError_t e = cudaMalloc(&base, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SDFB_ERR_OOM, "allocating %zu bytes of link buffers failed: %s", bytes, cudaGetErrorString(e)); }
    // flags start at 0 (below any run << 32 with run >= 1); zeroed and complete BEFORE any neighbour learns the address
    CU(cudaMemset(base, 0, bytes));
    CU(cudaDeviceSynchronize());
    if (p->tun.link_trace) {
        CU(cudaMalloc(reinterpret_cast<void **>(&p->link.trace), 2 * LINK_SWEEPS * sizeof(unsigned long long)));
        CU(cudaMemset(p->link.trace, 0, 2 * LINK_SWEEPS * sizeof(unsigned long long)));
    }
    p->link.in_halo = static_cast<uint64_t *>(base);
    p->link.in_flags = reinterpret_cast<unsigned long long *>(static_cast<char *>(base) + halo_bytes);
    p->link.bytes = bytes;
    return SDFB_OK;
}

}  // namespace

extern "C" {

const char *sdfb_version(void) { return "sdfgen-b200 0.2 (sm_100a)"; }
const char *sdfb_last_error(void) { return g_err; }
uint64_t sdfb_launch_count(void) { return g_launches.load(); }

int sdfb_trim_memory(void)
{
    std::lock_guard<std::mutex> lock(g_pool_mtx);
    for (auto &dp : g_pools) {
        if (!dp.pool) continue;
        CU(cudaStreamSynchronize(dp.st));                         // pending frees
        CU(cudaMemPoolTrimTo(dp.pool, 0));
    }
    return SDFB_OK;
}

int sdfb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int usable = 0;
    for (int d = 0; d < n; ++d) if (device_is_sm100(d)) ++usable;
    return usable;
}

int sdfb_plan_create(sdfb_plan **out, int device, int32_t ni, int32_t nj, int32_t nk,
                     int32_t k_lo, int32_t k_hi, uint32_t flags)
{
    if (!out) return fail(SDFB_ERR_INVALID, "plan pointer is null");
    *out = nullptr;
    if (ni <= 0 || nj <= 0 || nk <= 0) return fail(SDFB_ERR_INVALID, "grid dimensions must be positive (got %d x %d x %d)", ni, nj, nk);
    if (ni > 32767 || nj > 32767 || nk > 32767) return fail(SDFB_ERR_INVALID, "grid dimensions above 32767 are not supported");
    if (k_lo < 0 || k_hi > nk || k_lo >= k_hi) return fail(SDFB_ERR_INVALID, "bad slab [%d,%d) of %d planes", k_lo, k_hi, nk);
    if (flags & SDFB_SWEEP_STRIPS) return fail(SDFB_ERR_INVALID, "the strips schedule (flag 0x8) was an experiment of round 1 and is no longer built");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SDFB_ERR_NO_DEVICE, "no CUDA device is visible; libsdfb has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(SDFB_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    if (!device_is_sm100(device)) return fail(SDFB_ERR_NO_DEVICE, "device %d is not an sm_100 (B200) part; libsdfb ships sm_100a code only", device);
    DeviceGuard dg(device);
    if (!dg.ok) return fail(SDFB_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    sdfb_plan *p = new (std::nothrow) sdfb_plan();
    if (!p) return fail(SDFB_ERR_OOM, "host allocation failed");
    p->device = device; p->flags = flags;
    p->g.ni = ni; p->g.nj = nj; p->g.nk = nk; p->g.k_lo = k_lo; p->g.k_hi = k_hi;
    p->g.dx = 1.f; p->g.ox = p->g.oy = p->g.oz = 0.f; p->g.band = 1;
    const size_t V = (size_t)p->g.slab_voxels();
    cudaError_t e;
    if ((e = dev_alloc(&p->cells, ((size_t)p->g.cell_count() + 8) * sizeof(uint64_t))) != cudaSuccess ||   // +8: bulk prefetches round up to 16 B
        (e = dev_alloc(&p->counts, V * sizeof(int32_t))) != cudaSuccess ||
        (e = dev_alloc(&p->phi, V * sizeof(float))) != cudaSuccess ||
        ((flags & SDFB_OUT_KFASTEST) && (e = dev_alloc(&p->phi_k, V * sizeof(float))) != cudaSuccess) ||
        (e = dev_alloc(&p->changed, 4 * sizeof(unsigned long long))) != cudaSuccess) {   // changed, evaluations, inside count, bad triangle
        sdfb_plan_destroy(p);
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? SDFB_ERR_OOM : SDFB_ERR_CUDA, "device allocation of %zu voxels failed: %s", V, cudaGetErrorString(e));
    }
    p->tun = tuning_from_env();
    p->progress_words = std::max(sweep_columns_progress_words(p->g), sweep_columns_progress_words_ek12(p->g));   // either build may run
    if (p->progress_words && (e = dev_alloc(&p->progress, p->progress_words * sizeof(uint32_t))) != cudaSuccess) {
        sdfb_plan_destroy(p);
        cudaGetLastError();
        return fail(SDFB_ERR_OOM, "device allocation failed: %s", cudaGetErrorString(e));
    }
    for (auto &ev : p->ev) {
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) { sdfb_plan_destroy(p); return fail(SDFB_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)); }
    }
    if ((e = cudaEventCreateWithFlags(&p->ev_copy, cudaEventDisableTiming)) != cudaSuccess) { sdfb_plan_destroy(p); return fail(SDFB_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)); }
    cudaMemset(p->changed, 0, 2 * sizeof(unsigned long long));
    cudaMemset(p->changed + 3, 0xff, sizeof(unsigned long long));
    *out = p;
    return SDFB_OK;
}

int sdfb_plan_destroy(sdfb_plan *p)
{
    if (!p) return SDFB_OK;
    DeviceGuard dg(p->device);
    plan_quiesce(p);
    free_mesh(p);
    dev_free(p->cells); dev_free(p->counts); dev_free(p->phi); dev_free(p->phi_k); dev_free(p->scratch);
    dev_free(p->changed); dev_free(p->progress); dev_free(p->relax);
    link_release(p);
    for (auto &ev : p->ev) if (ev) cudaEventDestroy(ev);
    if (p->ev_copy) cudaEventDestroy(p->ev_copy);
    delete p;
    cudaGetLastError();
    return SDFB_OK;
}

int sdfb_plan_set_concurrency(sdfb_plan *p, int32_t plans_in_flight)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if (plans_in_flight < 1) return fail(SDFB_ERR_INVALID, "plans_in_flight must be at least 1");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    p->max_ctas = plans_in_flight == 1 ? 0 : sm_share_ctas(sms, plans_in_flight);
    return SDFB_OK;
}

int sdfb_plan_set_mesh_host(sdfb_plan *p, const uint32_t *tri, uint64_t ntri, const float *xyz, uint64_t nvert, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if ((ntri && !tri) || (nvert && !xyz)) return fail(SDFB_ERR_INVALID, "mesh pointer is null");
    if (ntri > SDFB_MAX_TRIANGLES) return fail(SDFB_ERR_LIMIT, "%llu triangles exceed the limit of %u", (unsigned long long)ntri, SDFB_MAX_TRIANGLES);
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (ntri > p->tri_own_cap || !p->tri_own) {      // grow-only staging buffers: repeated calls do not reallocate
        if (p->tri_own) { plan_quiesce(p); dev_free(p->tri_own); }
        p->tri_own = nullptr; p->tri_own_cap = 0;
        CU(dev_alloc(&p->tri_own, (ntri ? ntri : 1) * 3 * sizeof(uint32_t)));
        p->tri_own_cap = ntri ? ntri : 1;
    }
    if (nvert > p->xyz_own_cap || !p->xyz_own) {
        if (p->xyz_own) { plan_quiesce(p); dev_free(p->xyz_own); }
        p->xyz_own = nullptr; p->xyz_own_cap = 0;
        CU(dev_alloc(&p->xyz_own, (nvert ? nvert : 1) * 3 * sizeof(float)));
        p->xyz_own_cap = nvert ? nvert : 1;
    }
    auto upload = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (bytes >= ((size_t)4 << 20) && is_pageable(src)) return staged_h2d(dst, src, bytes, st);
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    };
    if (ntri) CU(upload(p->tri_own, tri, ntri * 3 * sizeof(uint32_t)));
    if (nvert) CU(upload(p->xyz_own, xyz, nvert * 3 * sizeof(float)));
    return build_records(p, p->tri_own, p->xyz_own, ntri, nvert, st);
}

int sdfb_plan_set_mesh_device(sdfb_plan *p, const uint32_t *d_tri, uint64_t ntri, const float *d_xyz, uint64_t nvert, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if ((ntri && !d_tri) || (nvert && !d_xyz)) return fail(SDFB_ERR_INVALID, "mesh pointer is null");
    DeviceGuard dg(p->device);
    return build_records(p, d_tri, d_xyz, ntri, nvert, (cudaStream_t)stream);
}

int sdfb_plan_band(sdfb_plan *p, const float origin[3], float dx, int32_t exact_band, void *stream)
{
    if (!p || !origin) return fail(SDFB_ERR_INVALID, "null argument");
    if (!p->have_mesh) return fail(SDFB_ERR_STATE, "sdfb_plan_band called before a mesh was set");
    if (!(dx > 0.f)) return fail(SDFB_ERR_INVALID, "cell spacing dx must be positive");
    if (exact_band < 0 || exact_band > 1024) return fail(SDFB_ERR_INVALID, "exact_band %d out of range [0,1024]", exact_band);
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    p->g.dx = dx; p->g.ox = origin[0]; p->g.oy = origin[1]; p->g.oz = origin[2]; p->g.band = exact_band;
    // (ni+nj+nk)*dx: int sum converted to float, one float multiply (cpu_lib/makelevelset3.cpp:197)
    volatile float nsum = (float)(p->g.ni + p->g.nj + p->g.nk);
    p->init_phi = nsum * dx;
    CU(cudaEventRecord(p->ev[0], st));
    g_launches += launch_init(p->cells, p->g.cell_count(), p->init_phi, st);
    CU(cudaMemsetAsync(p->counts, 0, (size_t)p->g.slab_voxels() * sizeof(int32_t), st));
    CU(cudaMemsetAsync(p->changed, 0, 2 * sizeof(unsigned long long), st));
    if (p->progress) { CU(cudaMemsetAsync(p->progress, 0, p->progress_words * sizeof(uint32_t), st)); p->epoch = 0; }
    g_launches += launch_band(p->rec, p->ntri, p->g, p->units, p->ext, p->prefix, p->block_sums, p->cells, p->counts, p->init_phi, st);
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev[1], st));
    p->have_band = true; p->have_sign = false; p->timed = false; p->last_sweep = -1;
    p->look_next = p->look_hi = -1;
    if (p->link.active) ++p->link.run;       // linked slabs run band() in lockstep: flag words are run << 32 | steps
    return SDFB_OK;
}

int sdfb_plan_sweep(sdfb_plan *p, int32_t first, int32_t count, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if (!p->have_band) return fail(SDFB_ERR_STATE, "sdfb_plan_sweep called before sdfb_plan_band");
    if (first < 0 || count < 0) return fail(SDFB_ERR_INVALID, "bad sweep range");
    // the stamp memo assumes that every sweep below `first` has run on these cells since the last band (a neighbour whose
    // stamp is not newer than the last sweep that looked at it is skipped): sweeps may be repeated, never skipped
    if (first > p->last_sweep + 1)
        return fail(SDFB_ERR_STATE, "sweep %d requested but the last sweep run since sdfb_plan_band is %d: sweeps must be run in order", first, p->last_sweep);
    if (count == 0) return SDFB_OK;
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Tuning &tun = p->tun;
    auto reset_epoch_if_needed = [&](uint32_t more) -> cudaError_t {       // progress words are epoch<<16 | steps: start over before it wraps
        if (p->epoch + more < 65000u) return cudaSuccess;
        p->epoch = 0;
        return cudaMemsetAsync(p->progress, 0, p->progress_words * sizeof(uint32_t), st);
    };
    if (p->link.active) {
        // exact multi-GPU mode: every sweep runs in the fused column launch, which hands the slab's boundary plane to the
        // downstream neighbour column by column (LinkSweep, sdfb_kernels.cuh).  No host synchronisation in here: the
        // neighbour's launch may be enqueued by this very thread right after this one.
        if (p->flags & (SDFB_SWEEP_LEVELS | SDFB_SWEEP_RELAX)) return fail(SDFB_ERR_STATE, "linked plans use the column schedule; SDFB_SWEEP_LEVELS / SDFB_SWEEP_RELAX cannot be combined with links");
        if (first + count > LINK_SWEEPS) return fail(SDFB_ERR_INVALID, "linked plans run the reference's %d sweeps only (asked for %d..%d)", LINK_SWEEPS, first, first + count - 1);
        // a sweep's hand-over buffers and flags are written once per run: repeating an index would find the flags already raised
        if (first != p->last_sweep + 1) return fail(SDFB_ERR_STATE, "linked plans run every sweep exactly once per sdfb_plan_band, in order (asked for %d after %d)", first, p->last_sweep);
        for (int side = 0; side < 2; ++side) {
            const bool has = side == 0 ? p->g.k_lo > 0 : p->g.k_hi < p->g.nk;
            if (has && !p->link.peer_halo[side]) return fail(SDFB_ERR_STATE, "linked plan: the slab %s this one has not been linked (sdfb_plan_link_import / _local)", side == 0 ? "below" : "above");
        }
        CU(reset_epoch_if_needed((uint32_t)count));
        // (8 x 16 build unless SDFB_COL_SHAPE=12 is set -- on EVERY slab of the run: the hand-over is per J column and EJ is 8 in
        // both builds, but a slab's lag behind its upstream neighbour is counted in that neighbour's EK)
        const int l = tun.col_shape == 12
            ? launch_sweep_columns_fused_ek12(p->cells, p->rec, p->g, first, count, p->changed, p->progress, p->progress_words,
                                              &p->epoch, st, tun, p->max_ctas > 0 ? (p->max_ctas * 4 + 2) / 3 : 0, &p->link)
            : launch_sweep_columns_fused(p->cells, p->rec, p->g, first, count, p->changed, p->progress, p->progress_words,
                                         &p->epoch, st, tun, p->max_ctas, &p->link);
        if (!l) return fail(SDFB_ERR_STATE, "linked plan: the fused column launch declined sweeps %d..%d on slab [%d,%d)", first, first + count - 1, p->g.k_lo, p->g.k_hi);
        g_launches += l;
        if (first + count - 1 > p->last_sweep) p->last_sweep = first + count - 1;
        CU(cudaGetLastError());
        CU(cudaEventRecord(p->ev[2], st));
        p->have_sign = false;
        return SDFB_OK;
    }
    // first sweep index handled by the relaxation schedule: the reference's second pass and everything after it
    int relax_from = (p->flags & SDFB_SWEEP_RELAX) ? 0 : ((p->flags & SDFB_SWEEP_COLUMNS) ? 1 << 30 : 8);
    if (tun.relax_from >= 0 && !(p->flags & (SDFB_SWEEP_RELAX | SDFB_SWEEP_COLUMNS))) relax_from = tun.relax_from;
    // The column sweeps of the first pass go out as ONE launch in which consecutive sweeps overlap where their directions
    // allow it (sdfb_sweep_columns.cu: k_sweep_columns_fused).  SDFB_FUSE_PASS=0 turns it off (one launch per sweep),
    // =1 also fuses launches of >= 300 M voxels, which otherwise stay on the per-sweep path.
    int fused_until = first;
    const int fuse_mode = tun.fuse_pass;                                        // -1 default, 0 off, 1 forced
    const bool big_launch = (int64_t)p->g.ni * (p->g.nj - 1) * p->g.nkl() >= ((int64_t)300 << 20);
    // which build of the column schedule: 8 x 12 columns (160-thread CTAs, four per SM) below 300 M voxels, 8 x 16 above
    // (sdfb_sweep_columns.cu); linked plans always run 8 x 16 (their hand-over buffers are laid out for it)
    const bool ek12 = tun.col_shape ? tun.col_shape == 12 : !big_launch;
    // max_ctas (several plans share the device) is this plan's share of THREE CTAs per SM; the 8 x 12 build fits four
    const int max_ctas12 = p->max_ctas > 0 ? (p->max_ctas * 4 + 2) / 3 : 0;
    auto columns = [&](int s, const unsigned int *run_if) {
        return ek12 ? launch_sweep_columns_ek12(p->cells, p->rec, p->g, s, p->changed, p->progress, ++p->epoch, st, tun, run_if, max_ctas12)
                    : launch_sweep_columns(p->cells, p->rec, p->g, s, p->changed, p->progress, ++p->epoch, st, tun, run_if, p->max_ctas);
    };
    if (fuse_mode != 0 && (fuse_mode == 1 || !big_launch) &&
        !(p->flags & (SDFB_SWEEP_LEVELS | SDFB_SWEEP_RELAX)) && first < 8 && first < relax_from) {
        int n = (first + count < 8 ? first + count : 8);
        if (n > relax_from) n = relax_from;
        n -= first;
        if (n >= 2) {
            CU(reset_epoch_if_needed((uint32_t)n));
            const int l = ek12 ? launch_sweep_columns_fused_ek12(p->cells, p->rec, p->g, first, n, p->changed, p->progress, p->progress_words,
                                                                  &p->epoch, st, tun, max_ctas12)
                               : launch_sweep_columns_fused(p->cells, p->rec, p->g, first, n, p->changed, p->progress, p->progress_words,
                                                            &p->epoch, st, tun, p->max_ctas);
            if (l) { g_launches += l; fused_until = first + n; p->look_next = p->look_hi = -1; }
        }
    }
    for (int s = fused_until; s < first + count; ++s) {
        if (p->flags & SDFB_SWEEP_LEVELS) {
            p->look_next = p->look_hi = -1;
            g_launches += launch_sweep_levels(p->cells, p->rec, p->g, s, p->changed, st);
        } else if (s >= relax_from && s + 1 < 31 && s > p->last_sweep && sweep_relax_supported(p->g)) {
            if (!p->relax) {
                const size_t bytes = sweep_relax_scratch_bytes(p->g);
                CU(dev_alloc(&p->relax, bytes));
                CU(cudaMemsetAsync(p->relax, 0, bytes, st));
            }
            // Lookahead (sdfb_sweep_relax.cu: k_look_scan): the sweeps of the second pass and later that this call still has
            // to run are scanned for in ONE pass over the cells; each then starts from the window's lists.  The window only
            // holds while its sweeps follow one another through this branch.
            bool look = false;
            if (tun.lookahead && s >= 8 && s + 1 < 31) {
                if (!(p->look_next == s && s < p->look_hi)) {
                    int hi = first + count < s + 8 ? first + count : s + 8;
                    if (hi > 30) hi = 30;
                    p->look_next = p->look_hi = -1;
                    if (hi - s >= 2) {
                        const int l = launch_look_scan(p->cells, p->rec, p->g, s, hi, p->changed, p->relax, st, tun, p->max_ctas);
                        if (l) { g_launches += l; p->look_next = s; p->look_hi = hi; p->look_lo = s; }
                    }
                }
                look = p->look_next == s && s < p->look_hi;
                if (look) ++p->look_next;
            }
            g_launches += launch_sweep_relax(p->cells, p->rec, p->g, s, p->changed, p->relax, st, tun, p->max_ctas, look);
            // a sweep that turns out to change a large part of the grid is handed back (cells restored, flag set):
            // this launch then runs it with the column schedule, and exits at once otherwise
            CU(reset_epoch_if_needed(1));
            g_launches += columns(s, sweep_relax_fallback_flag(p->relax));
        } else {
            p->look_next = p->look_hi = -1;
            CU(reset_epoch_if_needed(1));
            g_launches += columns(s, nullptr);
        }
    }
    if (first + count - 1 > p->last_sweep) p->last_sweep = first + count - 1;
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev[2], st));
    p->have_sign = false;
    return SDFB_OK;
}

int sdfb_plan_sign(sdfb_plan *p, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if (!p->have_band) return fail(SDFB_ERR_STATE, "sdfb_plan_sign called before sdfb_plan_band");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (p->copy_pending) { CU(cudaStreamWaitEvent(st, p->ev_copy, 0)); p->copy_pending = false; }   // phi is still being read
    g_launches += launch_sign(p->cells, p->counts, p->g, !(p->flags & SDFB_NO_SIGN), false, p->phi, st);
    if (p->flags & SDFB_OUT_KFASTEST)
        g_launches += launch_relayout_i32(reinterpret_cast<const int32_t *>(p->phi), p->g, reinterpret_cast<int32_t *>(p->phi_k), st);
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev[3], st));
    p->have_sign = true; p->timed = true;
    return SDFB_OK;
}

int sdfb_plan_run(sdfb_plan *p, const float origin[3], float dx, int32_t exact_band, void *stream)
{
    int rc = sdfb_plan_band(p, origin, dx, exact_band, stream);
    if (rc) return rc;
    rc = sdfb_plan_sweep(p, 0, 16, stream);
    if (rc) return rc;
    return sdfb_plan_sign(p, stream);
}

int sdfb_plan_device_ptrs(sdfb_plan *p, void **cells, void **counts, void **phi)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if (cells) *cells = p->cells;
    if (counts) *counts = p->counts;
    if (phi) *phi = (p->flags & SDFB_OUT_KFASTEST) ? p->phi_k : p->phi;
    return SDFB_OK;
}

int sdfb_plan_halo_refresh(sdfb_plan *p, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    DeviceGuard dg(p->device);
    g_launches += launch_halo_refresh(p->cells, p->g, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return SDFB_OK;
}

int sdfb_plan_counters(sdfb_plan *p, void *stream, uint64_t out[2])
{
    if (!p || !out) return fail(SDFB_ERR_INVALID, "null argument");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, p->changed, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(p->changed, 0, sizeof(h), st));
    CU(cudaStreamSynchronize(st));
    out[0] = h[0]; out[1] = h[1];
    return SDFB_OK;
}

int sdfb_plan_verify(sdfb_plan *p, void *stream, uint64_t out[4])
{
    if (!p || !out) return fail(SDFB_ERR_INVALID, "null argument");
    if (!p->have_band) return fail(SDFB_ERR_STATE, "nothing to verify: run the plan first");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d = nullptr;
    CU(dev_alloc(&d, 4 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), st));
    g_launches += launch_verify_cells(p->cells, p->rec, p->g, p->init_phi, d, st);
    unsigned long long h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(d);
    CU(e);
    for (int q = 0; q < 4; ++q) out[q] = h[q];
    return SDFB_OK;
}

int sdfb_plan_changed(sdfb_plan *p, void *stream, uint64_t *changed)
{
    if (!changed) return fail(SDFB_ERR_INVALID, "null argument");
    uint64_t c[2];
    int rc = sdfb_plan_counters(p, stream, c);
    if (!rc) *changed = c[0];
    return rc;
}

// Copies the slab's results to host memory.  global == false: the three arrays hold the slab only (slab_voxels values).
// global == true: they are arrays of the WHOLE ni x nj x nk grid and the slab lands at its place -- a contiguous run of
// planes in the i-fastest layout, a strided copy (nkl values every nk) in the k-fastest one (python/sdfgen_py.cpp:80-86,
// common/sdf_io.cpp:49-57).
static int download_impl(sdfb_plan *p, float *phi_out, int32_t *tri_out, int32_t *count_out, cudaStream_t st, bool global)
{
    const size_t V = (size_t)p->g.slab_voxels();
    const bool kf = (p->flags & SDFB_OUT_KFASTEST) != 0;
    const size_t nkl = (size_t)p->g.nkl(), nk = (size_t)p->g.nk, rows = (size_t)p->g.plane();
    const bool strided = global && kf && nkl != nk;
    const size_t off = !global ? 0 : (kf ? (size_t)p->g.k_lo : (size_t)p->g.k_lo * rows);
    auto copy_out = [&](void *dst, const void *src) -> cudaError_t {     // 4-byte elements
        char *d = static_cast<char *>(dst) + off * 4;
        if (strided) return cudaMemcpy2DAsync(d, nk * 4, src, nkl * 4, nkl * 4, rows, cudaMemcpyDeviceToHost, st);
        if (V * 4 >= ((size_t)64 << 20) && is_pageable(d)) return staged_d2h(d, src, V * 4, st);
        return cudaMemcpyAsync(d, src, V * 4, cudaMemcpyDeviceToHost, st);
    };
    if (phi_out) {
        if (!p->have_sign) {   // unsigned phi straight from the cells
            if (p->copy_pending) { CU(cudaStreamWaitEvent(st, p->ev_copy, 0)); p->copy_pending = false; }
            g_launches += launch_sign(p->cells, p->counts, p->g, false, false, p->phi, st);
            if (kf) g_launches += launch_relayout_i32(reinterpret_cast<const int32_t *>(p->phi), p->g, reinterpret_cast<int32_t *>(p->phi_k), st);
        }
        CU(copy_out(phi_out, kf ? p->phi_k : p->phi));
    }
    if (tri_out || (count_out && kf)) {
        if (!p->scratch) CU(dev_alloc(&p->scratch, V * sizeof(int32_t) * 2));
    }
    if (tri_out) {
        g_launches += launch_unpack_tri(p->cells, p->g, false, p->scratch, st);
        if (kf) g_launches += launch_relayout_i32(p->scratch, p->g, p->scratch + V, st);
        CU(copy_out(tri_out, kf ? p->scratch + V : p->scratch));
    }
    if (count_out) {
        if (kf) {
            g_launches += launch_relayout_i32(p->counts, p->g, p->scratch, st);
            CU(copy_out(count_out, p->scratch));
        } else {
            CU(copy_out(count_out, p->counts));
        }
    }
    CU(cudaGetLastError());
    unsigned long long bad = ~0ull;
    CU(cudaMemcpyAsync(&bad, p->changed + 3, sizeof(bad), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (bad != ~0ull) return bad_index_error(p, bad);
    return SDFB_OK;
}

int sdfb_plan_download(sdfb_plan *p, float *phi_out, int32_t *tri_out, int32_t *count_out, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if (!p->have_band) return fail(SDFB_ERR_STATE, "nothing to download: run the plan first");
    DeviceGuard dg(p->device);
    return download_impl(p, phi_out, tri_out, count_out, (cudaStream_t)stream, false);
}

int sdfb_plan_download_global(sdfb_plan *p, float *phi_grid, int32_t *tri_grid, int32_t *count_grid, void *stream)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    if (!p->have_band) return fail(SDFB_ERR_STATE, "nothing to download: run the plan first");
    DeviceGuard dg(p->device);
    return download_impl(p, phi_grid, tri_grid, count_grid, (cudaStream_t)stream, true);
}

int sdfb_plan_download_phi_async(sdfb_plan *p, float *phi_out, void *copy_stream)
{
    if (!p || !phi_out) return fail(SDFB_ERR_INVALID, "null argument");
    if (!p->have_sign) return fail(SDFB_ERR_STATE, "sdfb_plan_download_phi_async needs a completed sdfb_plan_sign");
    DeviceGuard dg(p->device);
    cudaStream_t cs = (cudaStream_t)copy_stream;
    const size_t V = (size_t)p->g.slab_voxels();
    CU(cudaStreamWaitEvent(cs, p->ev[3], 0));                     // the sign pass that produced phi
    CU(cudaMemcpyAsync(phi_out, (p->flags & SDFB_OUT_KFASTEST) ? p->phi_k : p->phi, V * sizeof(float), cudaMemcpyDeviceToHost, cs));
    CU(cudaEventRecord(p->ev_copy, cs));
    p->copy_pending = true;
    return SDFB_OK;
}

int sdfb_plan_write_sdf(sdfb_plan *p, const char *path, const float min_box[3], float dx, int64_t *inside_count_out, void *stream)
{
    if (!p || !path || !min_box) return fail(SDFB_ERR_INVALID, "null argument");
    if (!p->have_sign) return fail(SDFB_ERR_STATE, "sdfb_plan_write_sdf needs a completed sdfb_plan_sign");
    if (p->g.k_lo != 0 || p->g.k_hi != p->g.nk) return fail(SDFB_ERR_STATE, "sdfb_plan_write_sdf needs a whole-grid plan (this one holds the slab [%d,%d) of %d planes)", p->g.k_lo, p->g.k_hi, p->g.nk);
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t V = (size_t)p->g.slab_voxels();
    // the file stores k fastest (common/sdf_io.cpp:49-57): take the plan's own k-fastest copy or make one in the scratch
    const float *src = p->phi_k;
    if (!src) {
        if (!p->scratch) CU(dev_alloc(&p->scratch, V * sizeof(int32_t) * 2));      // same size as the download path allocates
        g_launches += launch_relayout_i32(reinterpret_cast<const int32_t *>(p->phi), p->g, p->scratch, st);
        src = reinterpret_cast<const float *>(p->scratch);
    }
    // inside count = number of values < 0 (-0.0f does not count, common/sdf_io.cpp:53), reduced on the device
    CU(cudaMemsetAsync(p->changed + 2, 0, sizeof(unsigned long long), st));
    g_launches += launch_count_negative(src, (int64_t)V, p->changed + 2, st);
    CU(cudaGetLastError());

    FILE *f = fopen(path, "wb");
    if (!f) return fail(SDFB_ERR_IO, "Failed to open file for writing: %s", path);
    // header: 3 x int32 dims, 3 x float32 bounds_min, 3 x float32 bounds_max = min + n * dx in float (common/sdf_io.cpp:22-45)
    int32_t dims[3] = {p->g.ni, p->g.nj, p->g.nk};
    float bounds[6];
    for (int a = 0; a < 3; ++a) {
        bounds[a] = min_box[a];
        volatile float ext = (float)dims[a] * dx;             // product rounded to float before the sum, as in the reference
        bounds[3 + a] = min_box[a] + ext;
    }
    bool ok = fwrite(dims, sizeof(dims), 1, f) == 1 && fwrite(bounds, sizeof(bounds), 1, f) == 1;
    cudaError_t e = cudaSuccess;
    if (ok) e = staged_d2h_sink(src, V * sizeof(float), st, [f](const char *data, size_t n, size_t) { return fwrite(data, 1, n, f) == n; }, &ok);
    if (fclose(f) != 0) ok = false;
    if (e != cudaSuccess) return fail(SDFB_ERR_CUDA, "copying phi to the host failed: %s", cudaGetErrorString(e));
    if (!ok) return fail(SDFB_ERR_IO, "Failed to write SDF data to file: %s", path);
    unsigned long long tail[2] = {0, ~0ull};                         // inside count, lowest triangle with a bad vertex index
    CU(cudaMemcpyAsync(tail, p->changed + 2, sizeof(tail), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (inside_count_out) *inside_count_out = (int64_t)tail[0];
    if (tail[1] != ~0ull) return bad_index_error(p, tail[1]);
    return SDFB_OK;
}

int sdfb_plan_phase_ms(sdfb_plan *p, float out[4])
{
    if (!p || !out) return fail(SDFB_ERR_INVALID, "null argument");
    if (!p->timed) return fail(SDFB_ERR_STATE, "no completed run to time (band, sweep and sign must all have been enqueued)");
    DeviceGuard dg(p->device);
    CU(cudaEventSynchronize(p->ev[3]));
    CU(cudaEventElapsedTime(&out[0], p->ev[0], p->ev[1]));
    CU(cudaEventElapsedTime(&out[1], p->ev[1], p->ev[2]));
    CU(cudaEventElapsedTime(&out[2], p->ev[2], p->ev[3]));
    CU(cudaEventElapsedTime(&out[3], p->ev[0], p->ev[3]));
    return SDFB_OK;
}

// One-shot call, large grid, phi only: the download is as long as the second pass (C2: 13.7 ms for 537 MB into pageable
// memory against 14.3 ms of kernels), and the second pass changes a few ten thousand of the 134 M values.  So the output
// is produced EARLY -- sign and layout from the cells as they are after the first pass -- and copied to the host on a second
// stream while the second pass runs; what the second pass changes is exactly the lookahead window's c_list
// (sdfb_sweep_relax.cu), which comes back as {index, value} patches that the host applies.  Exact whenever it is used: if
// the window was not complete (a list overflowed, a sweep was handed back, no window at all), the plain path below runs
// instead.  Returns SDFB_OK with *done = true when phi_out holds the result; *done = false (and SDFB_OK) = not applicable or
// not usable this time, nothing was written that the plain path will not overwrite.  SDFB_EARLY_COPY=0 turns it off.
namespace {
constexpr uint32_t PATCH_CAP = 1u << 20;      // patches per call; more changed cells than that -> plain path
cudaStream_t g_copy_stream[64] = {nullptr};
std::mutex g_copy_stream_mtx;

int oneshot_early_copy(sdfb_plan *p, const float origin[3], float dx, int32_t exact_band, float *phi_out, bool *done, bool *ran)
{
    *done = false; *ran = false;
    const Grid &g = p->g;
    const size_t V = (size_t)g.slab_voxels(), out_bytes = V * sizeof(float);
    if (!p->tun.early_copy || !p->tun.lookahead || (p->flags & (SDFB_SWEEP_LEVELS | SDFB_SWEEP_COLUMNS)) || p->link.active ||
        g.k_lo != 0 || g.k_hi != g.nk || out_bytes < ((size_t)64 << 20) || !sweep_relax_supported(g) || V >= ((size_t)1 << 32))
        return SDFB_OK;
    if (p->tun.relax_from >= 0 && p->tun.relax_from != 8) return SDFB_OK;
    const int dev = p->device;
    cudaStream_t cs = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_copy_stream_mtx);
        if (dev < 0 || dev >= 64) return SDFB_OK;
        if (!g_copy_stream[dev]) { CU(cudaStreamCreateWithFlags(&g_copy_stream[dev], cudaStreamNonBlocking)); }
        cs = g_copy_stream[dev];
    }
    cudaStream_t st = nullptr;                                   // the one-shot call's work runs on the default stream
    const bool kf = (p->flags & SDFB_OUT_KFASTEST) != 0;
    *ran = true;
    int rc = sdfb_plan_band(p, origin, dx, exact_band, st);
    if (!rc) rc = sdfb_plan_sweep(p, 0, 8, st);
    if (rc) return rc;
    // the early output, and the event after which it may be copied
    g_launches += launch_sign(p->cells, p->counts, g, !(p->flags & SDFB_NO_SIGN), false, p->phi, st);
    if (kf) g_launches += launch_relayout_i32(reinterpret_cast<const int32_t *>(p->phi), g, reinterpret_cast<int32_t *>(p->phi_k), st);
    CU(cudaGetLastError());
    cudaEvent_t early = nullptr;
    CU(cudaEventCreateWithFlags(&early, cudaEventDisableTiming));
    struct EventGuard { cudaEvent_t e; ~EventGuard() { if (e) cudaEventDestroy(e); } } eg{early};
    CU(cudaEventRecord(early, st));
    rc = sdfb_plan_sweep(p, 8, 8, st);
    if (rc) return rc;
    const bool window = p->look_lo == 8 && p->look_hi == 16 && p->look_next == 16;
    uint32_t *d_patch = nullptr;
    unsigned int *d_head = nullptr;
    struct DevGuard { void *a, *b; ~DevGuard() { if (a) dev_free(a); if (b) dev_free(b); } } dg{nullptr, nullptr};
    if (window) {
        CU(dev_alloc(&d_patch, (size_t)PATCH_CAP * 8));
        dg.a = d_patch;
        CU(dev_alloc(&d_head, 16));
        dg.b = d_head;
        if (!launch_look_patches(p->cells, p->phi, g, kf, p->relax, p->tun, d_patch, reinterpret_cast<float *>(d_patch + PATCH_CAP), PATCH_CAP, d_head, st))
            return SDFB_OK;                                      // (cannot happen for a grid that had a window)
        ++g_launches;
        CU(cudaGetLastError());
    }
    // all of the above is asynchronous: touch the (usually fresh, pageable) output pages while the first pass runs
    const bool pageable = is_pageable(phi_out);
    if (pageable) parallel_for_bytes(reinterpret_cast<char *>(phi_out), out_bytes, touch_pages, nullptr);
    if (window) {
        CU(cudaStreamWaitEvent(cs, early, 0));
        const float *src = kf ? p->phi_k : p->phi;
        if (pageable) CU(staged_d2h(phi_out, src, out_bytes, cs));
        else CU(cudaMemcpyAsync(phi_out, src, out_bytes, cudaMemcpyDeviceToHost, cs));
        CU(cudaStreamSynchronize(cs));
    }
    unsigned int head[4] = {0, 0, 1, 0};
    unsigned long long bad = ~0ull;
    if (window) CU(cudaMemcpyAsync(head, d_head, 12, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&bad, p->changed + 3, sizeof(bad), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (bad != ~0ull) return bad_index_error(p, bad);
    if (!window || head[2] != 0u || head[1] > PATCH_CAP || head[0] != head[1]) return SDFB_OK;     // plain path (sweeps are done)
    const unsigned n = head[0];
    if (n) {
        std::vector<uint32_t> idx(n);
        std::vector<float> val(n);
        CU(cudaMemcpy(idx.data(), d_patch, (size_t)n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(val.data(), d_patch + PATCH_CAP, (size_t)n * 4, cudaMemcpyDeviceToHost));
        for (unsigned e = 0; e < n; ++e) phi_out[idx[e]] = val[e];
    }
    *done = true;
    return SDFB_OK;
}
}  // namespace

int sdfb_make_level_set3(const uint32_t *tri, uint64_t ntri, const float *xyz, uint64_t nvert,
                         const float origin[3], float dx, int32_t ni, int32_t nj, int32_t nk,
                         int32_t exact_band, float *phi_out, int32_t *closest_tri_out,
                         int32_t *intersection_count_out, uint32_t flags)
{
    if (!phi_out || !origin) return fail(SDFB_ERR_INVALID, "null argument");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(SDFB_ERR_NO_DEVICE, "no CUDA device is visible; libsdfb has no CPU fallback"); }
    sdfb_plan *p = nullptr;
    int rc = sdfb_plan_create(&p, dev, ni, nj, nk, 0, nk, flags);
    if (rc) return rc;
    p->own_stream = true; p->stream = nullptr;                   // everything below runs on the default stream
    rc = sdfb_plan_set_mesh_host(p, tri, ntri, xyz, nvert, nullptr);
    bool done = false, ran = false;
    if (!rc && !closest_tri_out && !intersection_count_out) {
        DeviceGuard dg(p->device);
        rc = oneshot_early_copy(p, origin, dx, exact_band, phi_out, &done, &ran);
    }
    if (!rc && !done) {
        // plain path: run (unless the early-copy attempt already ran band and sweeps and only its patches were unusable),
        // sign, blocking download
        if (!ran) rc = sdfb_plan_run(p, origin, dx, exact_band, nullptr);
        else rc = sdfb_plan_sign(p, nullptr);
        // everything above is asynchronous: touch the (usually fresh, pageable) output pages while the GPU computes
        const size_t out_bytes = (size_t)ni * nj * nk * sizeof(float);
        if (!rc && !ran && out_bytes >= ((size_t)64 << 20) && is_pageable(phi_out)) parallel_for_bytes(reinterpret_cast<char *>(phi_out), out_bytes, touch_pages, nullptr);
        if (!rc) rc = sdfb_plan_download(p, phi_out, closest_tri_out, intersection_count_out, nullptr);
    }
    sdfb_plan_destroy(p);
    return rc;
}

// Many meshes / grids per call (the reference lists a batch API as wanted, README.md:216-221; small grids are
// launch- and latency-bound, so one grid at a time leaves most SMs idle).  `concurrency` worker threads each own a
// non-blocking stream and a plan that is reused while consecutive items have the same dimensions; the persistent
// sweep grids are capped to the worker's share of the SMs so that the kernels of different items really overlap.
int sdfb_make_level_set3_batch(sdfb_batch_item *items, int32_t n, int32_t concurrency, uint32_t flags)
{
    if (n < 0 || (n > 0 && !items)) return fail(SDFB_ERR_INVALID, "bad batch");
    if (n == 0) return SDFB_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(SDFB_ERR_NO_DEVICE, "no CUDA device is visible; libsdfb has no CPU fallback"); }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || !device_is_sm100(dev)) {
        cudaGetLastError();
        return fail(SDFB_ERR_NO_DEVICE, "device %d is not an sm_100 (B200) part; libsdfb ships sm_100a code only", dev);
    }
    int W = concurrency <= 0 ? 4 : concurrency;
    if (W > 16) W = 16;
    if (W > n) W = n;
    std::atomic<int32_t> next{0};
    std::mutex err_mtx;
    int first_rc = SDFB_OK;
    char first_msg[sizeof(g_err)] = "";
    auto worker = [&]() {
        cudaStream_t st = nullptr;
        sdfb_plan *p = nullptr;
        int rc0 = cudaSetDevice(dev) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess
                      ? SDFB_OK : fail(SDFB_ERR_CUDA, "could not create a stream on device %d", dev);
        for (;;) {
            const int32_t i = next.fetch_add(1);
            if (i >= n) break;
            sdfb_batch_item &it = items[i];
            int rc = rc0;
            if (!rc && !it.phi_out) rc = fail(SDFB_ERR_INVALID, "batch item %d: phi_out is null", i);
            if (!rc && (!p || p->g.ni != it.ni || p->g.nj != it.nj || p->g.nk != it.nk)) {
                sdfb_plan_destroy(p); p = nullptr;
                rc = sdfb_plan_create(&p, dev, it.ni, it.nj, it.nk, 0, it.nk, flags);
                if (!rc) { p->own_stream = true; p->stream = st; }
                if (!rc && W > 1) p->max_ctas = sm_share_ctas(sms, W);
            }
            if (!rc) rc = sdfb_plan_set_mesh_host(p, it.tri, it.ntri, it.xyz, it.nvert, st);
            if (!rc) rc = sdfb_plan_run(p, it.origin, it.dx, it.exact_band, st);
            if (!rc) rc = sdfb_plan_download(p, it.phi_out, nullptr, nullptr, st);
            it.status = rc;
            if (rc) {
                std::lock_guard<std::mutex> lock(err_mtx);
                if (!first_rc) { first_rc = rc; snprintf(first_msg, sizeof(first_msg), "batch item %d: %.480s", i, g_err); }
            }
        }
        sdfb_plan_destroy(p);
        if (st) cudaStreamDestroy(st);
    };
    std::vector<std::thread> th;
    for (int w = 1; w < W; ++w) th.emplace_back(worker);
    {
        DeviceGuard dg(dev);
        worker();                                                 // the calling thread is worker 0
    }
    for (auto &t : th) t.join();
    if (first_rc) return fail(first_rc, "%s", first_msg);
    return SDFB_OK;
}

// ---- exact multi-GPU mode: linked k-slabs ---------------------------------------------------------------------------

int sdfb_plan_link_export(sdfb_plan *p, void *handle_out)
{
    if (!p || !handle_out) return fail(SDFB_ERR_INVALID, "null argument");
    DeviceGuard dg(p->device);
    int rc = link_ensure_buffers(p);
    if (rc) return rc;
    LinkHandle h{};
    cudaError_t e = cudaIpcGetMemHandle(&h.mem, p->link.in_halo);
    if (e != cudaSuccess) { cudaGetLastError(); memset(&h.mem, 0, sizeof(h.mem)); }      // same-process links still work
    h.ni = p->g.ni; h.nj = p->g.nj; h.nk = p->g.nk; h.k_lo = p->g.k_lo; h.k_hi = p->g.k_hi; h.device = p->device;
    h.bytes = p->link.bytes; h.local_ptr = (uint64_t)(uintptr_t)p->link.in_halo; h.pid = (int64_t)getpid();
    memset(handle_out, 0, SDFB_LINK_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    p->link.active = true;
    return SDFB_OK;
}

int sdfb_plan_link_import(sdfb_plan *p, int32_t side, const void *handle)
{
    if (!p || !handle) return fail(SDFB_ERR_INVALID, "null argument");
    if (side != 0 && side != 1) return fail(SDFB_ERR_INVALID, "side must be 0 (the slab below) or 1 (the slab above)");
    LinkHandle h;
    memcpy(&h, handle, sizeof(h));
    if (h.ni != p->g.ni || h.nj != p->g.nj || h.nk != p->g.nk) return fail(SDFB_ERR_INVALID, "link: the neighbour's grid is %d x %d x %d, this plan's %d x %d x %d", h.ni, h.nj, h.nk, p->g.ni, p->g.nj, p->g.nk);
    if (side == 0 ? h.k_hi != p->g.k_lo : h.k_lo != p->g.k_hi)
        return fail(SDFB_ERR_INVALID, "link: slab [%d,%d) is not the slab %s [%d,%d)", h.k_lo, h.k_hi, side == 0 ? "below" : "above", p->g.k_lo, p->g.k_hi);
    DeviceGuard dg(p->device);
    int rc = link_ensure_buffers(p);
    if (rc) return rc;
    if (p->link.peer_base[side]) { cudaIpcCloseMemHandle(p->link.peer_base[side]); p->link.peer_base[side] = nullptr; }
    void *base = nullptr;
    if (h.pid == (int64_t)getpid()) {
        // same process: the exporter's pointer is valid here (unified addressing); another device needs peer access
        base = (void *)(uintptr_t)h.local_ptr;
        if (h.device != p->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, p->device, h.device));
            if (!can) return fail(SDFB_ERR_NO_DEVICE, "link: device %d cannot access device %d's memory (no P2P)", p->device, h.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(SDFB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
    } else {
        cudaError_t e = cudaIpcOpenMemHandle(&base, h.mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(SDFB_ERR_CUDA, "cudaIpcOpenMemHandle failed: %s (slabs in different processes need CUDA IPC between their devices)", cudaGetErrorString(e)); }
        p->link.peer_base[side] = base;
    }
    const size_t halo_bytes = (size_t)LINK_SWEEPS * (size_t)p->g.plane() * sizeof(uint64_t);
    p->link.peer_halo[side] = static_cast<uint64_t *>(base);
    p->link.peer_flags[side] = reinterpret_cast<unsigned long long *>(static_cast<char *>(base) + halo_bytes);
    p->link.active = true;
    return SDFB_OK;
}

int sdfb_plan_link_trace(sdfb_plan *p, void *stream, uint64_t out[32])
{
    if (!p || !out) return fail(SDFB_ERR_INVALID, "null argument");
    if (!p->link.trace) return fail(SDFB_ERR_STATE, "no trace: create the plan with SDFB_LINK_TRACE=1 in the environment and link it");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long h[2 * LINK_SWEEPS];
    CU(cudaMemcpyAsync(h, p->link.trace, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(p->link.trace, 0, sizeof(h), st));
    CU(cudaStreamSynchronize(st));
    for (int s = 0; s < LINK_SWEEPS; ++s) { out[2 * s] = h[2 * s] ? ~h[2 * s] : 0; out[2 * s + 1] = h[2 * s + 1]; }
    return SDFB_OK;
}

int sdfb_plan_unlink(sdfb_plan *p)
{
    if (!p) return fail(SDFB_ERR_INVALID, "plan is null");
    DeviceGuard dg(p->device);
    plan_quiesce(p);
    // importer side only: the exported block itself is freed by sdfb_plan_destroy (or reused by the next export), i.e. after
    // every neighbour had the chance to close its mapping of it
    link_close_peers(p);
    return SDFB_OK;
}

int sdfb_slab_bounds(int32_t nk, int32_t slabs, int32_t index, int32_t *k_lo, int32_t *k_hi)
{
    if (nk <= 0 || slabs <= 0 || slabs > nk || index < 0 || index >= slabs || !k_lo || !k_hi) return fail(SDFB_ERR_INVALID, "cannot cut %d planes into %d slabs (index %d)", nk, slabs, index);
    const int base = nk / slabs, rem = nk % slabs;
    *k_lo = index * base + (index < rem ? index : rem);
    *k_hi = *k_lo + base + (index < rem ? 1 : 0);
    return SDFB_OK;
}

// One grid over several GPUs of this process: k-slabs, one plan per device, linked so that the 16 sweeps keep the
// reference's serial order across the slab faces (bit-identical to the single-GPU result).  Everything is enqueued on
// every device before the first blocking call: a slab's sweep kernel waits, on the device, for its neighbour's.
int sdfb_make_level_set3_multi(const uint32_t *tri, uint64_t ntri, const float *xyz, uint64_t nvert,
                               const float origin[3], float dx, int32_t ni, int32_t nj, int32_t nk,
                               int32_t exact_band, float *phi_out, int32_t *closest_tri_out,
                               int32_t *intersection_count_out, int32_t num_gpus, uint32_t flags)
{
    if (!phi_out || !origin) return fail(SDFB_ERR_INVALID, "null argument");
    const int avail = sdfb_device_count();
    if (avail == 0) return fail(SDFB_ERR_NO_DEVICE, "no CUDA device is visible; libsdfb has no CPU fallback");
    if (num_gpus <= 0) num_gpus = avail;                                   // 0 = all usable devices
    if (num_gpus > avail) return fail(SDFB_ERR_NO_DEVICE, "%d GPUs requested, %d usable", num_gpus, avail);
    // a slab needs two planes to be linked on a grid face; small grids simply use fewer devices
    while (num_gpus > 1 && nk / num_gpus < 2) --num_gpus;
    if (num_gpus == 1 || (flags & (SDFB_SWEEP_LEVELS | SDFB_SWEEP_RELAX)))
        return sdfb_make_level_set3(tri, ntri, xyz, nvert, origin, dx, ni, nj, nk, exact_band, phi_out, closest_tri_out, intersection_count_out, flags);
    const int N = num_gpus;
    std::vector<sdfb_plan *> plans(N, nullptr);
    std::vector<cudaStream_t> streams(N, nullptr);
    int rc = SDFB_OK;
    int dev0 = 0;
    cudaGetDevice(&dev0);
    auto cleanup = [&]() {
        for (int r = 0; r < N; ++r) if (plans[r]) { DeviceGuard dg(r); cudaStreamSynchronize(streams[r]); }
        for (int r = 0; r < N; ++r) {
            DeviceGuard dg(r);
            sdfb_plan_destroy(plans[r]);
            if (streams[r]) cudaStreamDestroy(streams[r]);
        }
        cudaSetDevice(dev0);
    };
    std::vector<char> handles((size_t)N * SDFB_LINK_HANDLE_BYTES);
    for (int r = 0; r < N && !rc; ++r) {
        int32_t lo, hi;
        sdfb_slab_bounds(nk, N, r, &lo, &hi);
        DeviceGuard dg(r);
        if (cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking) != cudaSuccess) { rc = fail(SDFB_ERR_CUDA, "could not create a stream on device %d", r); break; }
        rc = sdfb_plan_create(&plans[r], r, ni, nj, nk, lo, hi, flags);
        if (!rc) { plans[r]->own_stream = true; plans[r]->stream = streams[r]; }
        if (!rc) rc = sdfb_plan_link_export(plans[r], handles.data() + (size_t)r * SDFB_LINK_HANDLE_BYTES);
    }
    for (int r = 0; r < N && !rc; ++r) {
        if (r > 0) rc = sdfb_plan_link_import(plans[r], 0, handles.data() + (size_t)(r - 1) * SDFB_LINK_HANDLE_BYTES);
        if (!rc && r + 1 < N) rc = sdfb_plan_link_import(plans[r], 1, handles.data() + (size_t)(r + 1) * SDFB_LINK_HANDLE_BYTES);
    }
    // uploads and record building first (they may allocate, which synchronises), then the phases on every device
    for (int r = 0; r < N && !rc; ++r) rc = sdfb_plan_set_mesh_host(plans[r], tri, ntri, xyz, nvert, streams[r]);
    for (int r = 0; r < N && !rc; ++r) rc = sdfb_plan_band(plans[r], origin, dx, exact_band, streams[r]);
    int enq = 0;                                                            // sweeps enqueued on devices [0, enq)
    for (int r = 0; r < N && !rc; ++r) { rc = sdfb_plan_sweep(plans[r], 0, LINK_SWEEPS, streams[r]); if (!rc) enq = r + 1; }
    if (rc && enq > 0 && enq < N) {
        // a neighbour's launch failed while ours are already waiting for it: nothing can complete them but the watchdog;
        // this cannot happen for well-formed arguments (every check precedes the first launch), report it as fatal
        char keep[sizeof(g_err)]; snprintf(keep, sizeof(keep), "%.400s (after %d of %d slab launches; the device waits will time out)", g_err, enq, N);
        cleanup();
        return fail(rc, "%s", keep);
    }
    for (int r = 0; r < N && !rc; ++r) rc = sdfb_plan_sign(plans[r], streams[r]);
    const size_t out_bytes = (size_t)ni * nj * nk * sizeof(float);
    if (!rc && out_bytes >= ((size_t)64 << 20) && is_pageable(phi_out)) parallel_for_bytes(reinterpret_cast<char *>(phi_out), out_bytes, touch_pages, nullptr);
    if (!rc) {
        // each slab downloads into its place of the caller's arrays, one host thread per device
        std::vector<int> rcs(N, SDFB_OK);
        std::vector<std::string> msgs(N);
        std::vector<std::thread> th;
        for (int r = 0; r < N; ++r) th.emplace_back([&, r]() {
            rcs[r] = sdfb_plan_download_global(plans[r], phi_out, closest_tri_out, intersection_count_out, streams[r]);
            if (rcs[r]) msgs[r] = g_err;
        });
        for (auto &t : th) t.join();
        for (int r = 0; r < N && !rc; ++r) if (rcs[r]) { rc = rcs[r]; snprintf(g_err, sizeof(g_err), "%s", msgs[r].c_str()); }
    }
    char keep[sizeof(g_err)]; snprintf(keep, sizeof(keep), "%s", g_err);
    cleanup();
    if (rc) return fail(rc, "%s", keep);
    return SDFB_OK;
}

}  // extern "C"
