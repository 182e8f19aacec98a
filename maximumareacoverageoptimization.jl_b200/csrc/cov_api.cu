// cov_api.cu -- the C ABI of libcoverage_cuda (include/coverage_cuda.h): handle lifecycle, the
// device-resident cell store, parameter upload, and the batched evaluation pipelines.
//
// The reference (/root/reference, pure Julia) has no FFI; each entry point replaces the Julia
// lines cited beside its declaration in include/coverage_cuda.h. Nothing here computes coverage
// on the CPU: without a CUDA device every entry point that needs one fails with COV_ERR_CUDA.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>
#include "cov_device.cuh"
#include "cov_kernels.cuh"
#include "../../include/coverage_cuda.h"

using namespace cov;


// ------------------------------------------------------------------------------------------
// a few host threads for the staging copies of pageable buffers (a Julia Array is pageable: one
// thread's memcpy into the pinned staging buffer would cap the host path at ~10 GB/s)
// ------------------------------------------------------------------------------------------
class CopyPool {
public:
    explicit CopyPool(int n)
    {
        for (int k = 0; k < n; ++k) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_.store(true);
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    // memcpy split over the pool and the calling thread; returns when all of it is done
    void copy(void *dst, const void *src, size_t bytes)
    {
        const size_t parts = workers_.size() + 1;
        if (bytes < (1u << 20) || parts == 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t per = ((bytes + parts - 1) / parts + 4095) / 4096 * 4096;
        int posted = 0;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (size_t k = 1; k < parts; ++k) {
                const size_t off = k * per;
                if (off >= bytes) break;
                const size_t n = std::min(per, bytes - off);
                tasks_.push_back([=] { memcpy((char *)dst + off, (const char *)src + off, n); });
                ++posted;
            }
            pending_.fetch_add(posted);
            avail_.fetch_add(posted);
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(per, bytes));
        // the parts are short (a few hundred microseconds at most): wait for them without sleeping
        for (unsigned spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins) {
            if ((spins & 1023u) == 1023u) std::this_thread::yield();
            else relax();
        }
    }

private:
    static void relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // Workers stay awake for about a millisecond after their last job: the slices of one pipelined call arrive
    // every few hundred microseconds, and a futex wake-up per slice cost more than the copy itself.
    void run()
    {
        for (;;) {
            const auto idle_since = std::chrono::steady_clock::now();
            unsigned spins = 0;
            while (avail_.load(std::memory_order_acquire) == 0 && !stop_.load(std::memory_order_relaxed)) {
                relax();
                if ((++spins & 255u) == 0 &&
                    std::chrono::steady_clock::now() - idle_since > std::chrono::microseconds(1000)) {
                    std::unique_lock<std::mutex> lk(m_);
                    cv_.wait(lk, [this] { return stop_.load() || avail_.load() > 0; });
                    break;
                }
            }
            std::function<void()> job;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (tasks_.empty()) {
                    if (stop_.load()) return;
                    continue;
                }
                job = std::move(tasks_.back());
                tasks_.pop_back();
                avail_.fetch_sub(1);
            }
            job();
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> workers_;
    std::vector<std::function<void()>> tasks_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<int> pending_{0}, avail_{0};
    std::atomic<bool> stop_{false};
};

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct cov_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr; // stream = own_stream or an adopted one
    cudaStream_t s_in = nullptr, s_out = nullptr;        // copy engines of the host pipeline
    std::string err;

    // cell store
    bool have_grid = false;
    DevBuf mult, cls, planes, planes_q;
    GridDesc g{};
    int64_t n_entries = 0, n_cells = 0;
    int area_exact = 0;
    // The point list in the reference's order, for the ordered kernel (area_exact == 0 or COV_KERNEL_ORDERED).
    // ent_known: the host mirror follows the order cov_set_points / cov_add_points gave (removals pending in
    // ent_removed_pending are applied lazily: every entry of a covered cell goes, the others keep their order);
    // otherwise the order is createPOI's (i outer, j inner, duplicates together), built from the cell store on demand.
    std::vector<int> ent_cell_h;
    std::vector<unsigned char> ent_cls_h;
    bool ent_known = false, ent_removed_pending = false, ent_dev_valid = false;
    DevBuf ent_cell, ent_cls;

    // closure parameters
    bool have_params = false;
    ObjParams o{};
    DevBuf params; // 5N doubles: r_max, prev_x, prev_y, prev_z, cons3_G

    LaunchCfg cfg{};
    int64_t chunk = 0;

    // scratch
    DevBuf counter, stats, xyT, small_in, argmin_obj, argmin_idx, removed, overflow;
    DevBuf backup; // mult + cls before an append, restored when the append fails (overflow, mixed weights)
    // forest-fire automaton state (ping-pong) and its direction probabilities
    DevBuf fire[2], fire_p;
    int fire_cur = 0;
    bool have_fire = false;
    // evaluation window (device) and staging (pinned host)
    DevBuf dX, d_obj, d_count, d_feas, d_clscnt, d_prog;
    DevBuf d_raw[2]; // packed slices as they arrive (cov_eval_batch_packed), widened into dX
    void *h_in[2] = {nullptr, nullptr};
    size_t h_in_cap = 0;
    void *h_out = nullptr;
    size_t h_out_cap = 0;
    void *h_small = nullptr; // pinned scratch: 4 KiB of scalars, then one candidate
    void *h_poll = nullptr;  // pinned scratch of the small-batch path

    std::vector<cudaEvent_t> ev_pool; // recycled
    // (start, stop) events around every coverage-kernel launch since the last drain
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_spans;
    size_t last_call_begin = 0;   // first span of the last eval call
    double drained_ms = 0;        // spans already folded into the running total
    int64_t drained_launches = 0;
    double last_call_ms_cache = -1;
    int64_t launches = 0;
    LaunchInfo last_info{};
    int counter_next = 0; // next unused slot of the zeroed counter ring
    CopyPool *pool = nullptr; // created on the first large pageable transfer
    // COV_OPT_TRACE: per-slice timeline of the last host-path call (ms since its first copy was queued)
    int trace = 0;
    int zero_copy_out = 1; // COV_OPT_ZEROCOPY_OUT: kernels write results straight into pinned host memory
    std::vector<cudaEvent_t> trace_ev; // start, then per slice: h2d done, kernel start, kernel end, d2h done
    std::vector<double> trace_ms;
};

static thread_local std::string g_err_nohandle;
constexpr int kCounterSlots = 4096; // work-dispenser counters, zeroed in bulk (one fresh slot per launch)

static int fail(cov_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    else g_err_nohandle = msg;
    return code;
}
static int fail_cuda(cov_handle *h, cudaError_t e, const char *what)
{
    (void)cudaGetLastError(); // clear the sticky-free error state
    return fail(h, COV_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                 \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return fail_cuda(h, e_, #call);   \
    } while (0)

static int ensure(cov_handle *h, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap && b.p) return COV_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    bytes = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        b.p = nullptr;
        return fail(h, e == cudaErrorMemoryAllocation ? COV_ERR_NOMEM : COV_ERR_CUDA,
                    std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    b.cap = bytes;
    return COV_OK;
}
#define OK(call)                   \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != COV_OK) return rc_; \
    } while (0)

static int ensure_pinned(cov_handle *h, void **p, size_t *cap, size_t bytes)
{
    if (*p && bytes <= *cap) return COV_OK;
    if (*p) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaFreeHost(*p));
        *p = nullptr;
        *cap = 0;
    }
    cudaError_t e = cudaMallocHost(p, std::max<size_t>(bytes, 4096));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *p = nullptr;
        return fail(h, COV_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    }
    *cap = std::max<size_t>(bytes, 4096);
    return COV_OK;
}

static cudaEvent_t get_event(cov_handle *h)
{
    if (!h->ev_pool.empty()) {
        cudaEvent_t e = h->ev_pool.back();
        h->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
// Fold every recorded span into the running totals (synchronises the main stream).
static cudaError_t drain_spans(cov_handle *h)
{
    if (h->kernel_spans.empty()) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return e;
    double last = 0;
    for (size_t k = 0; k < h->kernel_spans.size(); ++k) {
        float t = 0;
        e = cudaEventElapsedTime(&t, h->kernel_spans[k].first, h->kernel_spans[k].second);
        if (e != cudaSuccess) return e;
        h->drained_ms += t;
        h->drained_launches += 1;
        if (k >= h->last_call_begin) last += t;
        h->ev_pool.push_back(h->kernel_spans[k].first);
        h->ev_pool.push_back(h->kernel_spans[k].second);
    }
    h->last_call_ms_cache = last;
    h->kernel_spans.clear();
    h->last_call_begin = 0;
    return cudaSuccess;
}
// A new eval call starts: remember where its spans begin; keep the event pool bounded.
static void recycle_spans(cov_handle *h)
{
    if (h->kernel_spans.size() > 2048) (void)drain_spans(h);
    h->last_call_begin = h->kernel_spans.size();
    h->last_call_ms_cache = -1;
}

// Device-side address of pinned host memory (differs from the host address for some registered buffers);
// nullptr when the buffer cannot be written from the device.
static void *device_view(void *host)
{
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, host, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return d;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
extern "C" int cov_abi_version(void) { return COV_ABI_VERSION; }

extern "C" int cov_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int cov_create(int device, cov_handle **out)
{
    cov_handle *h = nullptr; // for the macros
    if (!out) return fail(nullptr, COV_ERR_INVALID, "cov_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, COV_ERR_CUDA,
                    std::string("cov_create: no usable CUDA device (") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                        "); libcoverage_cuda has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, COV_ERR_INVALID, "cov_create: device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, COV_ERR_CUDA,
                    std::string("cov_create: device ") + prop.name + " is sm_" + std::to_string(prop.major) +
                        std::to_string(prop.minor) + "; this library is built for sm_100a only");
    cov_handle *nh = new cov_handle();
    nh->device = device;
    h = nh;
    auto bail = [&](cudaError_t ce, const char *what) {
        int rc = fail_cuda(nullptr, ce, what);
        delete nh;
        return rc;
    };
    if ((e = cudaStreamCreateWithFlags(&nh->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&nh->s_in, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&nh->s_out, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    nh->stream = nh->own_stream;
    nh->cfg.kernel = COV_KERNEL_AUTO;
    nh->cfg.plane_mode = -1;
    nh->cfg.num_sms = prop.multiProcessorCount;
    nh->cfg.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if ((e = cudaMallocHost(&nh->h_small, 4096 + 3 * kMaxUavs * 8)) != cudaSuccess) return bail(e, "cudaMallocHost");
    int rc = ensure(nh, nh->counter, kCounterSlots * sizeof(unsigned long long));
    if (rc == COV_OK && cudaMemset(nh->counter.p, 0, kCounterSlots * sizeof(unsigned long long)) != cudaSuccess)
        rc = fail_cuda(nh, cudaGetLastError(), "cudaMemset");
    if (rc == COV_OK) rc = ensure(nh, nh->stats, 256);
    if (rc == COV_OK) rc = ensure(nh, nh->removed, 256);
    if (rc == COV_OK) rc = ensure(nh, nh->overflow, 256);
    if (rc == COV_OK) rc = ensure(nh, nh->argmin_obj, 1024 * sizeof(double));
    if (rc == COV_OK) rc = ensure(nh, nh->argmin_idx, 1024 * sizeof(long long));
    if (rc != COV_OK) {
        g_err_nohandle = nh->err;
        delete nh;
        return rc;
    }
    *out = nh;
    return COV_OK;
}

extern "C" void cov_destroy(cov_handle *h)
{
    if (!h) return;
    DeviceGuard dg(h->device);
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->s_in);
    cudaStreamSynchronize(h->s_out);
    DevBuf *bufs[] = {&h->mult, &h->cls, &h->planes, &h->planes_q, &h->params, &h->counter, &h->stats, &h->xyT,
                      &h->small_in, &h->argmin_obj, &h->argmin_idx, &h->removed, &h->overflow, &h->dX,
                      &h->d_obj, &h->d_count, &h->d_feas, &h->d_clscnt, &h->d_prog, &h->fire[0], &h->fire[1], &h->fire_p,
                      &h->backup, &h->ent_cell, &h->ent_cls, &h->d_raw[0], &h->d_raw[1]};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (int k = 0; k < 2; ++k)
        if (h->h_in[k]) cudaFreeHost(h->h_in[k]);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->h_poll) cudaFreeHost(h->h_poll);
    delete h->pool;
    (void)drain_spans(h);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->trace_ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->s_in);
    cudaStreamDestroy(h->s_out);
    cudaStreamDestroy(h->own_stream);
    (void)cudaGetLastError();
    delete h;
}

extern "C" const char *cov_last_error(const cov_handle *h) { return h ? h->err.c_str() : g_err_nohandle.c_str(); }

extern "C" int cov_set_option(cov_handle *h, int option, int64_t value)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "cov_set_option: NULL handle");
    switch (option) {
    case COV_OPT_KERNEL:
        if (value < COV_KERNEL_AUTO || value > COV_KERNEL_ORDERED) return fail(h, COV_ERR_INVALID, "unknown kernel id");
        h->cfg.kernel = (int)value;
        return COV_OK;
    case COV_OPT_WARPS_PER_CTA:
        if (value < 0 || value > 32) return fail(h, COV_ERR_INVALID, "warps per CTA must be 0..32");
        h->cfg.warps_per_cta = (int)value;
        return COV_OK;
    case COV_OPT_CTAS_PER_SM:
        if (value < 0 || value > 32) return fail(h, COV_ERR_INVALID, "CTAs per SM must be 0..32");
        h->cfg.ctas_per_sm = (int)value;
        return COV_OK;
    case COV_OPT_BAND_ROWS:
        if (value < 0 || value > kMaxDim) return fail(h, COV_ERR_INVALID, "band rows out of range");
        h->cfg.band_rows = (int)value;
        return COV_OK;
    case COV_OPT_FORCE_EXACT:
        h->cfg.force_exact = value != 0;
        return COV_OK;
    case COV_OPT_CHUNK:
        if (value < 0) return fail(h, COV_ERR_INVALID, "chunk must be >= 0");
        h->chunk = value;
        return COV_OK;
    case COV_OPT_TRACE:
        h->trace = value != 0;
        return COV_OK;
    case COV_OPT_ZEROCOPY_OUT:
        h->zero_copy_out = value != 0;
        return COV_OK;
    case COV_OPT_PLANE_MODE:
        if (value < -1 || value > 4) return fail(h, COV_ERR_INVALID, "plane mode must be -1..4");
        h->cfg.plane_mode = (int)value;
        return COV_OK;
    case COV_OPT_PROGRESSIVE_INDEX:
        if (value < 0 || value > kMaxUavs) return fail(h, COV_ERR_INVALID, "progressive index must be 0..1024");
        h->o.prog_which = (int)value;
        return COV_OK;
    }
    return fail(h, COV_ERR_INVALID, "unknown option");
}

extern "C" int cov_get_option(const cov_handle *h, int option, int64_t *value)
{
    if (!h || !value) return COV_ERR_INVALID;
    switch (option) {
    case COV_OPT_KERNEL: *value = h->cfg.kernel; return COV_OK;
    case COV_OPT_WARPS_PER_CTA: *value = h->cfg.warps_per_cta; return COV_OK;
    case COV_OPT_CTAS_PER_SM: *value = h->cfg.ctas_per_sm; return COV_OK;
    case COV_OPT_BAND_ROWS: *value = h->cfg.band_rows; return COV_OK;
    case COV_OPT_FORCE_EXACT: *value = h->cfg.force_exact; return COV_OK;
    case COV_OPT_CHUNK: *value = h->chunk; return COV_OK;
    case COV_OPT_TRACE: *value = h->trace; return COV_OK;
    case COV_OPT_ZEROCOPY_OUT: *value = h->zero_copy_out; return COV_OK;
    case COV_OPT_PLANE_MODE: *value = h->cfg.plane_mode; return COV_OK;
    case COV_OPT_PROGRESSIVE_INDEX: *value = h->o.prog_which; return COV_OK;
    }
    return COV_ERR_INVALID;
}

extern "C" void cov_get_limits(cov_limits *out)
{
    if (!out) return;
    out->max_uavs = kMaxUavs;
    out->max_nx = kMaxDim;
    out->max_ny = kMaxDim;
    out->max_planes = kMaxPlanes;
    out->max_classes = kMaxClasses;
}

extern "C" double cov_threshold(double R) { return threshold(R); }

// ------------------------------------------------------------------------------------------
// cell store
// ------------------------------------------------------------------------------------------
static bool f32_exact(double v) { return (double)(float)v == v; }

static int check_lattice(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy)
{
    if (nx < 1 || ny < 1) return fail(h, COV_ERR_INVALID, "grid: nx, ny must be >= 1");
    if (nx > kMaxDim || ny > kMaxDim) return fail(h, COV_ERR_LIMIT, "grid: nx, ny must be <= 32768");
    if (!(dx > 0) || !(dy > 0) || std::isinf(dx) || std::isinf(dy))
        return fail(h, COV_ERR_INVALID, "grid: dx, dy must be finite and > 0");
    return COV_OK;
}

static void describe_lattice(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy)
{
    GridDesc &g = h->g;
    g.nx = (int)nx;
    g.ny = (int)ny;
    g.wpr = (int)((nx + 31) / 32);
    g.stride = g.wpr | 1; // odd: 32 lanes on 32 consecutive rows hit 32 different banks
    g.plane_words = (int)(((int64_t)g.ny * g.stride + 3) / 4 * 4);
    g.qstride = (g.wpr + 1) & ~1; // even, and an odd number of word pairs (see GridDesc::planes_q)
    if (((g.qstride >> 1) & 1) == 0) g.qstride += 2;
    g.planes_q = nullptr;
    g.dx = dx;
    g.dy = dy;
    g.hdx = dx / 2;
    g.hdy = dy / 2;
    g.inv_dx = 1.0 / dx;
    g.inv_dy = 1.0 / dy;
    g.dxf = (float)dx;
    g.dyf = (float)dy;
    g.hdxf = (float)g.hdx;
    g.hdyf = (float)g.hdy;
    g.inv_dxf = (float)g.inv_dx;
    g.inv_dyf = (float)g.inv_dy;
    g.extent = (float)std::max((double)nx * dx, (double)ny * dy);
    g.extent = std::nextafter(g.extent, INFINITY);
    g.k_ca = std::nextafter((float)(0.75 * g.inv_dx * 1.000001), INFINITY);
    g.lattice_f32_exact = f32_exact(dx) && f32_exact(dy) && f32_exact(g.hdx) && f32_exact(g.hdy) &&
                          f32_exact((double)nx * dx) && f32_exact((double)ny * dy);
}

// v = m * 2^e with m odd (v finite, > 0)
static void odd_decompose(double v, uint64_t &m, int &e)
{
    int ex;
    double fr = std::frexp(v, &ex); // v = fr * 2^ex, fr in [0.5, 1)
    m = (uint64_t)std::ldexp(fr, 53);
    e = ex - 53;
    while (m && !(m & 1)) {
        m >>= 1;
        ++e;
    }
}

// Derive the bit planes from the device-resident mult/cls bytes and refresh the statistics.
static int rebuild_planes(cov_handle *h, int n_classes, const double *class_weight)
{
    GridDesc &g = h->g;
    const long long ncell = (long long)g.nx * g.ny;
    unsigned long long *d_stats = (unsigned long long *)h->stats.p;
    CK(launch_grid_stats((const unsigned char *)h->mult.p, (const unsigned char *)h->cls.p, ncell, d_stats,
                         h->stream));
    h->launches += 1;
    unsigned long long *hs = (unsigned long long *)h->h_small;
    CK(cudaMemcpyAsync(hs, d_stats, (2 + kMaxClasses) * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_entries = (int64_t)hs[0];
    h->n_cells = (int64_t)hs[1];
    g.n_classes = std::max(1, n_classes);
    for (int k = 0; k < kMaxClasses; ++k) g.class_weight[k] = (k < n_classes) ? class_weight[k] : 0.0;
    int np = 0;
    for (int k = 0; k < g.n_classes; ++k) {
        const unsigned orm = (unsigned)hs[2 + k];
        for (int b = 0; b < 8; ++b)
            if (orm & (1u << b)) {
                if (np >= kMaxPlanes)
                    return fail(h, COV_ERR_LIMIT,
                                "cell store needs more than 8 bit planes (weight classes x multiplicity bits)");
                g.plane_class[np] = k;
                g.plane_mult[np] = 1 << b;
                ++np;
            }
    }
    if (np == 0) { // empty store: one all-zero plane keeps the kernels uniform
        g.plane_class[0] = 0;
        g.plane_mult[0] = 1;
        np = 1;
    }
    for (int l = np; l < kMaxPlanes; ++l) {
        g.plane_class[l] = 0;
        g.plane_mult[l] = 0;
    }
    g.n_planes = np;
    OK(ensure(h, h->planes, (size_t)np * g.plane_words * 4));
    g.planes = (const uint32_t *)h->planes.p;
    CK(launch_pack_planes((const unsigned char *)h->mult.p, (const unsigned char *)h->cls.p, g,
                          (uint32_t *)h->planes.p, h->stream));
    h->launches += 1;
    g.planes_q = nullptr;
    if (np == 1) { // the sweep mode of the CTA kernel reads plane 0 in its own aligned, swizzled layout
        OK(ensure(h, h->planes_q, ((size_t)g.ny * g.qstride + 8) * 4));
        CK(cudaMemsetAsync(h->planes_q.p, 0, ((size_t)g.ny * g.qstride + 8) * 4, h->stream));
        CK(launch_requad_plane(g, (uint32_t *)h->planes_q.p, h->stream));
        h->launches += 1;
        g.planes_q = (const uint32_t *)h->planes_q.p;
    }
    // area_exact: every partial sum of the reference's list-order Float64 accumulation is an
    // integer multiple of a common power of two and stays below 2^53 of them, hence exact in any
    // order, and so is sum_k fl(w_k * count_k).
    h->area_exact = 1;
    {
        int qmin = 0;
        bool first = true, ok = true;
        uint64_t ms[kMaxClasses];
        int es[kMaxClasses];
        for (int k = 0; k < g.n_classes; ++k) {
            const double w = g.class_weight[k];
            if (w == 0.0) {
                ms[k] = 0;
                es[k] = 0;
                continue;
            }
            if (!(w > 0) || std::isinf(w)) {
                ok = false;
                break;
            }
            odd_decompose(w, ms[k], es[k]);
            if (first || es[k] < qmin) qmin = es[k];
            first = false;
        }
        if (ok) {
            long double total = 0;
            for (int k = 0; k < g.n_classes; ++k)
                if (ms[k]) total += std::ldexp((long double)ms[k], es[k] - qmin) * (long double)h->n_entries;
            ok = total < 9007199254740992.0L;
        }
        h->area_exact = ok ? 1 : 0;
    }
    h->have_grid = true;
    return COV_OK;
}

static int alloc_cells(cov_handle *h)
{
    const size_t ncell = (size_t)h->g.nx * h->g.ny;
    OK(ensure(h, h->mult, ncell + 4)); // add_points works on aligned 4-byte words
    OK(ensure(h, h->cls, ncell + 4));
    return COV_OK;
}

// ---- the ordered point list ------------------------------------------------------------------------------
static void entries_forget(cov_handle *h) // the store was re-created without a list: order = createPOI's, on demand
{
    h->ent_known = false;
    h->ent_removed_pending = false;
    h->ent_dev_valid = false;
    h->ent_cell_h.clear();
    h->ent_cls_h.clear();
    h->g.ent_cell = nullptr;
    h->g.ent_cls = nullptr;
    h->g.n_ent = 0;
}
// rmvCoveredPOI deletes every entry of a covered cell and keeps the order of the rest
// (src/CellFunctions.jl:81-108): drop the mirror's entries whose cell is now empty.
static int entries_apply_removals(cov_handle *h)
{
    if (!h->ent_known || !h->ent_removed_pending) return COV_OK;
    const size_t ncell = (size_t)h->g.nx * h->g.ny;
    std::vector<unsigned char> m(ncell);
    CK(cudaMemcpyAsync(m.data(), h->mult.p, ncell, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    size_t keep = 0;
    for (size_t p = 0; p < h->ent_cell_h.size(); ++p)
        if (m[(size_t)h->ent_cell_h[p]]) {
            h->ent_cell_h[keep] = h->ent_cell_h[p];
            h->ent_cls_h[keep] = h->ent_cls_h[p];
            ++keep;
        }
    h->ent_cell_h.resize(keep);
    h->ent_cls_h.resize(keep);
    h->ent_removed_pending = false;
    h->ent_dev_valid = false;
    return COV_OK;
}
// Make g.ent_cell / g.ent_cls / g.n_ent describe the current list on the device.
static int entries_on_device(cov_handle *h)
{
    if (h->ent_dev_valid) return COV_OK;
    OK(entries_apply_removals(h));
    if (!h->ent_known) {
        // no list was ever given (bit grid, cell bytes, createPOI, automaton): createPOI's order,
        // i outer and j inner (src/AreaCoverageCalculation.jl:14-15), the entries of a cell together
        const size_t ncell = (size_t)h->g.nx * h->g.ny;
        std::vector<unsigned char> m(ncell), k(ncell);
        CK(cudaMemcpyAsync(m.data(), h->mult.p, ncell, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(k.data(), h->cls.p, ncell, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->ent_cell_h.clear();
        h->ent_cls_h.clear();
        for (int i = 0; i < h->g.nx; ++i)
            for (int j = 0; j < h->g.ny; ++j) {
                const size_t cell = (size_t)i + (size_t)h->g.nx * j;
                for (unsigned r = 0; r < m[cell]; ++r) {
                    h->ent_cell_h.push_back((int)cell);
                    h->ent_cls_h.push_back((unsigned char)(k[cell] & (kMaxClasses - 1)));
                }
            }
    }
    const size_t P = h->ent_cell_h.size();
    if ((int64_t)P != h->n_entries)
        return fail(h, COV_ERR_STATE, "internal: the ordered point list disagrees with the cell store (" +
                                          std::to_string(P) + " vs " + std::to_string(h->n_entries) + " entries)");
    OK(ensure(h, h->ent_cell, std::max<size_t>(P, 1) * sizeof(int)));
    OK(ensure(h, h->ent_cls, std::max<size_t>(P, 1)));
    if (P) {
        CK(cudaMemcpyAsync(h->ent_cell.p, h->ent_cell_h.data(), P * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->ent_cls.p, h->ent_cls_h.data(), P, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream)); // pageable sources
    }
    if (!h->ent_known) { // the canonical list is cheap to rebuild and can be large: do not keep it
        h->ent_cell_h.clear();
        h->ent_cell_h.shrink_to_fit();
        h->ent_cls_h.clear();
        h->ent_cls_h.shrink_to_fit();
    }
    h->g.ent_cell = (const int *)h->ent_cell.p;
    h->g.ent_cls = (const unsigned char *)h->ent_cls.p;
    h->g.n_ent = (long long)P;
    h->ent_dev_valid = true;
    return COV_OK;
}

// An append (cov_add_points, cov_fire_step) can fail half-way: a multiplicity overflowing 255 or entries of
// different weights on one cell are found by the kernel that is already writing.  The store is snapshot
// before and put back on failure, so that mult/cls, the bit planes and n_entries never disagree.
static int snapshot_cells(cov_handle *h)
{
    const size_t ncell = (size_t)h->g.nx * h->g.ny + 4;
    OK(ensure(h, h->backup, 2 * ncell));
    CK(cudaMemcpyAsync(h->backup.p, h->mult.p, ncell, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync((char *)h->backup.p + ncell, h->cls.p, ncell, cudaMemcpyDeviceToDevice, h->stream));
    return COV_OK;
}
static void restore_cells(cov_handle *h)
{
    const size_t ncell = (size_t)h->g.nx * h->g.ny + 4;
    cudaError_t e = cudaMemcpyAsync(h->mult.p, h->backup.p, ncell, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->cls.p, (char *)h->backup.p + ncell, ncell, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { // cannot vouch for the store any more: make the next evaluation fail loudly
        (void)cudaGetLastError();
        h->have_grid = false;
    }
}

extern "C" int cov_set_grid_bits(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy,
                                 const uint32_t *bits, double weight)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!bits) return fail(h, COV_ERR_INVALID, "cov_set_grid_bits: bits is NULL");
    OK(check_lattice(h, nx, ny, dx, dy));
    h->have_grid = false;
    h->have_fire = false; // the automaton's state belongs to the lattice it was initialised on
    entries_forget(h);
    describe_lattice(h, nx, ny, dx, dy);
    OK(alloc_cells(h));
    const size_t words = (size_t)ny * ((nx + 31) / 32);
    DevBuf tmp;
    OK(ensure(h, tmp, words * 4));
    cudaError_t e = cudaMemcpyAsync(tmp.p, bits, words * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
        e = launch_bits_to_cells((const uint32_t *)tmp.p, (int)nx, (int)ny, (unsigned char *)h->mult.p,
                                 (unsigned char *)h->cls.p, h->stream);
    h->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp.p);
    if (e != cudaSuccess) return fail_cuda(h, e, "cov_set_grid_bits");
    return rebuild_planes(h, 1, &weight);
}

extern "C" int cov_set_grid_cells(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy,
                                  const uint8_t *mult, const uint8_t *cls, int64_t n_classes,
                                  const double *class_weight)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!mult || !class_weight) return fail(h, COV_ERR_INVALID, "cov_set_grid_cells: NULL argument");
    if (n_classes < 1 || n_classes > kMaxClasses)
        return fail(h, COV_ERR_LIMIT, "cov_set_grid_cells: n_classes must be 1..4");
    OK(check_lattice(h, nx, ny, dx, dy));
    const size_t ncell = (size_t)nx * ny;
    if (cls)
        for (size_t t = 0; t < ncell; ++t)
            if (cls[t] >= n_classes) return fail(h, COV_ERR_INVALID, "cov_set_grid_cells: class index out of range");
    h->have_grid = false;
    h->have_fire = false; // the automaton's state belongs to the lattice it was initialised on
    entries_forget(h);
    describe_lattice(h, nx, ny, dx, dy);
    OK(alloc_cells(h));
    CK(cudaMemcpyAsync(h->mult.p, mult, ncell, cudaMemcpyHostToDevice, h->stream));
    if (cls) CK(cudaMemcpyAsync(h->cls.p, cls, ncell, cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemsetAsync(h->cls.p, 0, ncell, h->stream));
    CK(launch_normalize_cls((const unsigned char *)h->mult.p, (unsigned char *)h->cls.p, (long long)ncell, h->stream));
    h->launches += 1;
    CK(cudaStreamSynchronize(h->stream));
    return rebuild_planes(h, (int)n_classes, class_weight);
}

extern "C" int cov_set_grid_full(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    OK(check_lattice(h, nx, ny, dx, dy));
    h->have_grid = false;
    h->have_fire = false; // the automaton's state belongs to the lattice it was initialised on
    entries_forget(h);
    describe_lattice(h, nx, ny, dx, dy);
    OK(alloc_cells(h));
    CK(launch_fill_full((unsigned char *)h->mult.p, (unsigned char *)h->cls.p, (long long)nx * ny, h->stream));
    h->launches += 1;
    const double w = dx * dy; // createPOI: area = weight = dx*dy
    return rebuild_planes(h, 1, &w);
}

// Map list entries onto the lattice (bit-exact centre check) and weight classes.
static int map_points(cov_handle *h, const double *pts5, int64_t P, std::vector<int> &cell,
                      std::vector<unsigned char> &pcls, int &n_classes, double *class_weight)
{
    const GridDesc &g = h->g;
    cell.resize((size_t)P);
    pcls.resize((size_t)P);
    for (int64_t p = 0; p < P; ++p) {
        const double x = pts5[5 * p], y = pts5[5 * p + 1], w = pts5[5 * p + 3];
        const double fi = std::nearbyint((x + g.hdx) / g.dx), fj = std::nearbyint((y + g.hdy) / g.dy);
        bool ok = fi >= 1 && fi <= g.nx && fj >= 1 && fj <= g.ny;
        if (ok) {
            volatile double cx = fi * g.dx; // individually rounded, as the reference forms them
            volatile double cy = fj * g.dy;
            ok = (cx - g.hdx == x) && (cy - g.hdy == y);
        }
        if (!ok) {
            char buf[160];
            snprintf(buf, sizeof buf, "point %lld (%.17g, %.17g) is not a cell centre of the %dx%d lattice",
                     (long long)p, x, y, g.nx, g.ny);
            return fail(h, COV_ERR_OFF_LATTICE, buf);
        }
        int k = -1;
        for (int q = 0; q < n_classes; ++q)
            if (class_weight[q] == w) {
                k = q;
                break;
            }
        if (k < 0) {
            if (w != w) return fail(h, COV_ERR_INVALID, "point weight is NaN");
            if (n_classes >= kMaxClasses)
                return fail(h, COV_ERR_LIMIT, "more than 4 distinct point weights");
            k = n_classes;
            class_weight[n_classes++] = w;
        }
        cell[(size_t)p] = ((int)fi - 1) + g.nx * ((int)fj - 1);
        pcls[(size_t)p] = (unsigned char)k;
    }
    return COV_OK;
}

static int upload_points(cov_handle *h, const std::vector<int> &cell, const std::vector<unsigned char> &pcls)
{
    const size_t P = cell.size();
    if (P == 0) return COV_OK;
    DevBuf dcell, dcls;
    int rc = ensure(h, dcell, P * sizeof(int));
    if (rc == COV_OK) rc = ensure(h, dcls, P);
    cudaError_t e = cudaSuccess;
    int *hov = (int *)h->h_small;
    if (rc == COV_OK) {
        e = cudaMemcpyAsync(dcell.p, cell.data(), P * sizeof(int), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dcls.p, pcls.data(), P, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(h->overflow.p, 0, sizeof(int), h->stream);
        if (e == cudaSuccess)
            e = launch_add_points((unsigned char *)h->mult.p, (unsigned char *)h->cls.p, (const int *)dcell.p,
                                  (const unsigned char *)dcls.p, (long long)P, (int *)h->overflow.p, h->stream);
        h->launches += 1;
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(hov, h->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    if (dcell.p) cudaFree(dcell.p);
    if (dcls.p) cudaFree(dcls.p);
    if (rc != COV_OK) return rc;
    if (e != cudaSuccess) return fail_cuda(h, e, "upload_points");
    if (*hov == 1) return fail(h, COV_ERR_LIMIT, "more than 255 list entries on one cell");
    if (*hov == 2) return fail(h, COV_ERR_INVALID, "entries on one cell carry different weights");
    return COV_OK;
}

extern "C" int cov_set_points(cov_handle *h, const double *pts5, int64_t P, int64_t nx, int64_t ny, double dx,
                              double dy)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (P < 0 || (P > 0 && !pts5)) return fail(h, COV_ERR_INVALID, "cov_set_points: bad list");
    OK(check_lattice(h, nx, ny, dx, dy));
    h->have_grid = false;
    h->have_fire = false; // the automaton's state belongs to the lattice it was initialised on
    entries_forget(h);
    describe_lattice(h, nx, ny, dx, dy);
    std::vector<int> cell;
    std::vector<unsigned char> pcls;
    int n_classes = 0;
    double cw[kMaxClasses] = {0, 0, 0, 0};
    OK(map_points(h, pts5, P, cell, pcls, n_classes, cw));
    if (n_classes == 0) {
        n_classes = 1;
        cw[0] = dx * dy;
    }
    OK(alloc_cells(h));
    const size_t ncell = (size_t)nx * ny;
    CK(cudaMemsetAsync(h->mult.p, 0, ncell + 4, h->stream));
    CK(cudaMemsetAsync(h->cls.p, 0xff, ncell + 4, h->stream));
    OK(upload_points(h, cell, pcls));
    h->ent_cell_h = cell; // the list order of the reference's points_of_interest
    h->ent_cls_h = pcls;
    h->ent_known = true;
    return rebuild_planes(h, n_classes, cw);
}

extern "C" int cov_add_points(cov_handle *h, const double *pts5, int64_t P)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!h->have_grid) return fail(h, COV_ERR_STATE, "cov_add_points: no grid set");
    if (P < 0 || (P > 0 && !pts5)) return fail(h, COV_ERR_INVALID, "cov_add_points: bad list");
    std::vector<int> cell;
    std::vector<unsigned char> pcls;
    int n_classes = h->g.n_classes;
    double cw[kMaxClasses];
    for (int k = 0; k < kMaxClasses; ++k) cw[k] = h->g.class_weight[k];
    if (h->n_entries == 0) n_classes = 0; // an empty store has no weight yet
    OK(map_points(h, pts5, P, cell, pcls, n_classes, cw));
    if (n_classes == 0) {
        n_classes = 1;
        cw[0] = h->g.class_weight[0];
    }
    OK(entries_apply_removals(h)); // before new entries can land on cells that a removal emptied
    OK(snapshot_cells(h));
    const int rc = upload_points(h, cell, pcls);
    if (rc != COV_OK) { // nothing was appended: the store, its planes and its counts stay as they were
        const std::string msg = h->err;
        restore_cells(h);
        h->err = msg;
        return rc;
    }
    if (h->ent_known) { // push! appends at the end (src/CellFunctions.jl:74)
        h->ent_cell_h.insert(h->ent_cell_h.end(), cell.begin(), cell.end());
        h->ent_cls_h.insert(h->ent_cls_h.end(), pcls.begin(), pcls.end());
    }
    h->ent_dev_valid = false;
    return rebuild_planes(h, n_classes, cw);
}

extern "C" int cov_get_grid_info(const cov_handle *h, cov_grid_info *info)
{
    if (!h || !info) return COV_ERR_INVALID;
    if (!h->have_grid) return COV_ERR_STATE;
    info->nx = h->g.nx;
    info->ny = h->g.ny;
    info->dx = h->g.dx;
    info->dy = h->g.dy;
    info->n_entries = h->n_entries;
    info->n_cells = h->n_cells;
    info->n_planes = h->g.n_planes;
    info->n_classes = h->g.n_classes;
    info->area_exact = h->area_exact;
    info->planes_in_smem = span_small_applies(h->g, h->have_params ? h->o.N : 1, h->cfg, 0, nullptr, nullptr) ? 1 : 0;
    return COV_OK;
}

extern "C" int cov_get_class_weights(const cov_handle *h, double *class_weight, int64_t cap)
{
    if (!h || !class_weight) return COV_ERR_INVALID;
    if (!h->have_grid) return COV_ERR_STATE;
    for (int64_t k = 0; k < cap && k < h->g.n_classes; ++k) class_weight[k] = h->g.class_weight[k];
    return COV_OK;
}

extern "C" int cov_get_grid_cells(cov_handle *h, uint8_t *mult)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!h->have_grid) return fail(h, COV_ERR_STATE, "no grid set");
    if (!mult) return fail(h, COV_ERR_INVALID, "NULL output");
    CK(cudaMemcpyAsync(mult, h->mult.p, (size_t)h->g.nx * h->g.ny, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return COV_OK;
}

// discs [x;y;R] -> device [x;y;T(R)]
static int upload_discs_T(cov_handle *h, const double *xyR, int64_t N)
{
    if (N < 1 || N > kMaxUavs) return fail(h, COV_ERR_LIMIT, "N must be 1..1024");
    if (!xyR) return fail(h, COV_ERR_INVALID, "NULL disc vector");
    OK(ensure(h, h->xyT, (size_t)3 * N * 8));
    OK(ensure(h, h->small_in, (size_t)3 * N * 8));
    CK(cudaMemcpyAsync(h->small_in.p, xyR, (size_t)3 * N * 8, cudaMemcpyHostToDevice, h->stream));
    CK(launch_thresholds((const double *)h->small_in.p, (int)N, (double *)h->xyT.p, h->stream));
    h->launches += 1;
    return COV_OK;
}

extern "C" int cov_remove_covered(cov_handle *h, const double *xyR, int64_t N, int64_t *removed)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!h->have_grid) return fail(h, COV_ERR_STATE, "cov_remove_covered: no grid set");
    OK(upload_discs_T(h, xyR, N));
    CK(launch_remove_covered((unsigned char *)h->mult.p, (unsigned char *)h->cls.p, h->g, (const double *)h->xyT.p, (int)N,
                             (unsigned long long *)h->removed.p, h->stream));
    h->launches += 1;
    unsigned long long *hr = (unsigned long long *)h->h_small + 16;
    CK(cudaMemcpyAsync(hr, h->removed.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (removed) *removed = (int64_t)*hr;
    if (*hr) {
        h->ent_removed_pending = h->ent_known;
        h->ent_dev_valid = false;
    }
    double cw[kMaxClasses];
    for (int k = 0; k < kMaxClasses; ++k) cw[k] = h->g.class_weight[k];
    return rebuild_planes(h, h->g.n_classes, cw);
}

extern "C" int cov_covered_mask(cov_handle *h, const double *x, uint8_t *mask)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!h->have_grid || !h->have_params) return fail(h, COV_ERR_STATE, "cov_covered_mask: grid/params not set");
    if (!mask) return fail(h, COV_ERR_INVALID, "NULL output");
    OK(upload_discs_T(h, x, h->o.N));
    const size_t ncell = (size_t)h->g.nx * h->g.ny;
    DevBuf dm;
    OK(ensure(h, dm, ncell));
    cudaError_t e = launch_covered_mask((unsigned char *)dm.p, h->g, (const double *)h->xyT.p, h->o.N, h->stream);
    h->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(mask, dm.p, ncell, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dm.p);
    if (e != cudaSuccess) return fail_cuda(h, e, "cov_covered_mask");
    return COV_OK;
}


// ------------------------------------------------------------------------------------------
// forest-fire automaton (src/DynamicArea.jl of the reference) feeding the cell store directly
// ------------------------------------------------------------------------------------------
extern "C" int cov_fire_init(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy, const uint8_t *state,
                             int32_t push_initial)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!state) return fail(h, COV_ERR_INVALID, "cov_fire_init: state is NULL");
    OK(check_lattice(h, nx, ny, dx, dy));
    const size_t ncell = (size_t)nx * ny;
    for (size_t t = 0; t < ncell; ++t)
        if (state[t] > 2) return fail(h, COV_ERR_INVALID, "cov_fire_init: cell states must be 0 (empty), 1 (tree), 2 (fire)");
    h->have_grid = false;
    h->have_fire = false;
    entries_forget(h);
    describe_lattice(h, nx, ny, dx, dy);
    OK(alloc_cells(h));
    OK(ensure(h, h->fire[0], ncell));
    OK(ensure(h, h->fire[1], ncell));
    OK(ensure(h, h->fire_p, 9 * sizeof(double)));
    CK(cudaMemcpyAsync(h->fire[0].p, state, ncell, cudaMemcpyHostToDevice, h->stream));
    CK(launch_fire_seed((const unsigned char *)h->fire[0].p, (unsigned char *)h->mult.p, (unsigned char *)h->cls.p,
                        (long long)ncell, push_initial, h->stream));
    h->launches += 1;
    CK(cudaStreamSynchronize(h->stream));
    h->fire_cur = 0;
    h->have_fire = true;
    const double w = dx * dy; // DynamicArea.jl:65: area = weight = dx*dy
    return rebuild_planes(h, 1, &w);
}

extern "C" int cov_fire_step(cov_handle *h, double wind_speed, double wind_direction, double prob_spread,
                             uint64_t seed, int64_t step, int32_t append, int64_t *n_pushed)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!h->have_fire) return fail(h, COV_ERR_STATE, "cov_fire_step: cov_fire_init has not been called");
    if (step < 0 || step > 0xffffffffll) return fail(h, COV_ERR_INVALID, "cov_fire_step: step out of range");
    if (h->g.n_classes != 1)
        return fail(h, COV_ERR_STATE, "cov_fire_step: the cell store holds several weight classes");
    // the probability of ignition from the neighbour at window index (a, b), a, b = 1..3, exactly as the
    // reference forms it: wind_speed * cos(wind_direction - atan(2-b, 2-a)) * prob_spread
    double *hp = (double *)h->h_small + 64;
    for (int b = 1; b <= 3; ++b)
        for (int a = 1; a <= 3; ++a) {
            volatile double ang = atan2((double)(2 - b), (double)(2 - a));
            volatile double c = cos(wind_direction - ang);
            volatile double p = wind_speed * c;
            hp[(a - 1) + 3 * (b - 1)] = p * prob_spread;
        }
    CK(cudaMemcpyAsync(h->fire_p.p, hp, 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->removed.p, 0, sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->overflow.p, 0, sizeof(int), h->stream));
    if (append) OK(snapshot_cells(h));
    const int cur = h->fire_cur;
    CK(launch_fire_step((const unsigned char *)h->fire[cur].p, (unsigned char *)h->fire[cur ^ 1].p,
                        (unsigned char *)h->mult.p, (unsigned char *)h->cls.p, h->g.nx, h->g.ny, seed,
                        (unsigned int)step, (const double *)h->fire_p.p, append, (unsigned long long *)h->removed.p,
                        (int *)h->overflow.p, h->stream));
    h->launches += 1;
    unsigned long long *hr = (unsigned long long *)h->h_small + 16;
    int *hov = (int *)h->h_small;
    CK(cudaMemcpyAsync(hr, h->removed.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hov, h->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (*hov) { // the step did not happen: automaton state and cell store stay as they were
        if (append) restore_cells(h);
        if (n_pushed) *n_pushed = 0;
        return fail(h, COV_ERR_LIMIT, "more than 255 list entries on one cell");
    }
    h->fire_cur = cur ^ 1;
    if (n_pushed) *n_pushed = (int64_t)*hr;
    if (!append || *hr == 0) return COV_OK;
    entries_forget(h); // the automaton pushes in cell order; from here on the list order is createPOI's
    double cw[kMaxClasses];
    for (int k = 0; k < kMaxClasses; ++k) cw[k] = h->g.class_weight[k];
    return rebuild_planes(h, 1, cw);
}

extern "C" int cov_fire_get_state(cov_handle *h, uint8_t *state)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!h->have_fire) return fail(h, COV_ERR_STATE, "cov_fire_get_state: cov_fire_init has not been called");
    if (!state) return fail(h, COV_ERR_INVALID, "NULL output");
    CK(cudaMemcpyAsync(state, h->fire[h->fire_cur].p, (size_t)h->g.nx * h->g.ny, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return COV_OK;
}

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
extern "C" int cov_set_params(cov_handle *h, int64_t N, const double *r_max, double penalty_scale,
                              const double *prev_xyR, const double *d_lim, double tan_half_fov, double sep_min,
                              int32_t use_cons7)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (N < 1 || N > kMaxUavs) return fail(h, COV_ERR_LIMIT, "cov_set_params: N must be 1..1024");
    if (!r_max) return fail(h, COV_ERR_INVALID, "cov_set_params: r_max is NULL");
    if (prev_xyR && !d_lim) return fail(h, COV_ERR_INVALID, "cov_set_params: d_lim is NULL while prev_xyR is set");
    std::vector<double> hp((size_t)5 * N, 0.0);
    for (int64_t i = 0; i < N; ++i) hp[i] = r_max[i];
    if (prev_xyR)
        for (int64_t i = 0; i < N; ++i) {
            hp[N + i] = prev_xyR[i];
            hp[2 * N + i] = prev_xyR[N + i];
            volatile double z = prev_xyR[2 * N + i] / tan_half_fov; // z1 = pre.R / tan(FOV/2)
            hp[3 * N + i] = z;
            hp[4 * N + i] = threshold_ge(d_lim[i]);
        }
    OK(ensure(h, h->params, (size_t)5 * N * 8));
    // pageable source: the copy is staged before the call returns, hp may go out of scope
    CK(cudaMemcpyAsync(h->params.p, hp.data(), (size_t)5 * N * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    ObjParams &o = h->o;
    o.N = (int)N;
    o.use_cons3 = prev_xyR != nullptr;
    o.use_cons7 = use_cons7 != 0;
    o.use_cons8 = sep_min > 0;
    o.penalty_scale = penalty_scale;
    o.tan_half_fov = tan_half_fov;
    {
        volatile double c7 = 19 * tan_half_fov;
        o.cons7_R = c7;
    }
    o.sep_T = o.use_cons8 ? threshold(sep_min) : 0.0;
    const double *base = (const double *)h->params.p;
    o.r_max = base;
    o.prev_x = base + N;
    o.prev_y = base + 2 * N;
    o.prev_z = base + 3 * N;
    o.cons3_G = base + 4 * N;
    h->have_params = true;
    return COV_OK;
}

// ------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------
static int launch_on_main(cov_handle *h, const double *dX, int64_t B, const EvalOut &out, bool timed)
{
    h->cfg.ordered = h->area_exact ? 0 : 1; // non-dyadic weights: the sum is replayed in list order
    if (h->cfg.ordered || h->cfg.kernel == COV_KERNEL_ORDERED) OK(entries_on_device(h));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        e0 = get_event(h);
        e1 = get_event(h);
        CK(cudaEventRecord(e0, h->stream));
    }
    if (h->counter_next >= kCounterSlots) { // ring used up: zero it again (stream-ordered after its last user)
        CK(cudaMemsetAsync(h->counter.p, 0, kCounterSlots * sizeof(unsigned long long), h->stream));
        h->counter_next = 0;
    }
    cudaError_t e = launch_eval(h->g, h->o, h->cfg, dX, (long long)B, out,
                                (unsigned long long *)h->counter.p + h->counter_next++,
                                h->stream, &h->last_info);
    if (e != cudaSuccess) {
        if (timed) {
            h->ev_pool.push_back(e0);
            h->ev_pool.push_back(e1);
        }
        if (e == cudaErrorInvalidConfiguration) {
            (void)cudaGetLastError();
            return fail(h, COV_ERR_LIMIT, "grid row / N too large for the shared-memory plan of this kernel");
        }
        return fail_cuda(h, e, "coverage kernel launch");
    }
    h->launches += 1;
    if (timed) {
        CK(cudaEventRecord(e1, h->stream));
        h->kernel_spans.emplace_back(e0, e1);
    }
    return COV_OK;
}

static int check_ready(cov_handle *h, const char *who)
{
    if (!h->have_grid) return fail(h, COV_ERR_STATE, std::string(who) + ": no grid set");
    if (!h->have_params) return fail(h, COV_ERR_STATE, std::string(who) + ": no parameters set");
    return COV_OK;
}

extern "C" int cov_eval_batch_device(cov_handle *h, const double *dX, int64_t B, double *d_obj, int64_t *d_count,
                                     uint8_t *d_feasible)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    OK(check_ready(h, "cov_eval_batch_device"));
    if (B < 0 || (B > 0 && (!dX || !d_obj))) return fail(h, COV_ERR_INVALID, "cov_eval_batch_device: bad arguments");
    if (B == 0) return COV_OK;
    recycle_spans(h);
    EvalOut out{};
    out.obj = d_obj;
    out.count = (long long *)d_count;
    out.feasible = d_feasible;
    return launch_on_main(h, dX, B, out, true);
}

// The host pipeline: candidates are cut into slices; slice k's H2D copy (stream s_in), the kernel
// of slice k-1 (main stream) and the D2H copy of slice k-2's results (stream s_out) overlap.
// Pageable host buffers go through two pinned staging buffers; pinned ones are DMA'd in place.
static void host_copy(cov_handle *h, void *dst, const void *src, size_t bytes)
{
    static const size_t pool_min = [] {
        const char *e = getenv("COV_POOL_MIN_BYTES"); // experiments; default: from 1 MiB on
        return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)(1u << 20);
    }();
    if (bytes >= pool_min) {
        if (!h->pool) {
            const unsigned hc = std::thread::hardware_concurrency();
            const char *e = getenv("COV_POOL_THREADS"); // experiments; default: up to 7 workers + the calling thread
            const unsigned want = e ? (unsigned)atoi(e) : 7u;
            h->pool = new CopyPool((int)std::min(std::max(want, 1u), std::max(2u, hc) - 1u));
        }
        h->pool->copy(dst, src, bytes);
    } else {
        memcpy(dst, src, bytes);
    }
}

// Small batches (a MADS poll set): one stream, pinned scratch, one synchronisation.
#ifndef COV_ZC_IN_LIMIT
#define COV_ZC_IN_LIMIT (256u << 10) // the whole small-batch path; measured against 32 KiB: 512 candidates 53 -> 42 us, 2184 candidates 74 -> 62 us
#endif
#ifndef COV_ZC_MID_LIMIT
// measured on B200 (5 UAVs, pinned buffers, tools/batch_size_sweep.py): 4096 candidates 70 -> 55 us, 32 768 168 -> 133 us,
// 131 072 465 -> 392 us, 262 144 761 -> 715 us; level with the copy-engine pipeline at 524 288 (SM reads reach
// ~49 GB/s over PCIe, the DMA engine 55 GB/s), behind it at 1 M
#define COV_ZC_MID_LIMIT (48u << 20)
#endif
static int eval_host_small(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count, uint8_t *feasible,
                           int64_t *class_count, double *progressive)
{
    const int N = h->o.N;
    const int ncls = h->g.n_classes;
    const size_t in_bytes = (size_t)B * 3 * N * 8;
    const size_t o_obj = 0, o_cnt = o_obj + (size_t)B * 8, o_cls = o_cnt + (count ? (size_t)B * 8 : 0),
                 o_prg = o_cls + (class_count ? (size_t)B * 8 * ncls : 0), o_fea = o_prg + (progressive ? (size_t)B * 8 : 0),
                 out_bytes = o_fea + (feasible ? (size_t)B : 0);
    if (!h->h_poll) { // 2 MiB of pinned scratch: candidates at 0 (<= 256 KiB), results from 512 KiB on
        cudaError_t e = cudaMallocHost(&h->h_poll, 2u << 20);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            h->h_poll = nullptr;
            return fail(h, COV_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        }
    }
    char *hin = (char *)h->h_poll, *hout = (char *)h->h_poll + (512u << 10);
    memcpy(hin, X, in_bytes);
    // poll-sized calls (a MADS poll set, the scalar closure): no copy engines at all -- the kernel reads the
    // candidates from the pinned scratch and writes the results into it over PCIe; the stream synchronise is
    // the only wait.  Saves two DMA set-ups per call, which is most of what a poll costs.
    char *vin = nullptr, *vout = nullptr;
    if (h->zero_copy_out && in_bytes <= (COV_ZC_IN_LIMIT)) {
        vin = (char *)device_view(hin);
        vout = (char *)device_view(hout);
    }
    if (vin && vout) {
        EvalOut out{};
        out.obj = (double *)(vout + o_obj);
        out.count = count ? (long long *)(vout + o_cnt) : nullptr;
        out.class_count = class_count ? (long long *)(vout + o_cls) : nullptr;
        out.progressive = progressive ? (double *)(vout + o_prg) : nullptr;
        out.feasible = feasible ? (unsigned char *)(vout + o_fea) : nullptr;
        OK(launch_on_main(h, (const double *)vin, B, out, true));
        CK(cudaStreamSynchronize(h->stream));
        const char *so = hout;
        memcpy(obj, so + o_obj, (size_t)B * 8);
        if (count) memcpy(count, so + o_cnt, (size_t)B * 8);
        if (class_count) memcpy(class_count, so + o_cls, (size_t)B * 8 * ncls);
        if (progressive) memcpy(progressive, so + o_prg, (size_t)B * 8);
        if (feasible) memcpy(feasible, so + o_fea, (size_t)B);
        return COV_OK;
    }
    OK(ensure(h, h->dX, std::max<size_t>(in_bytes, 1 << 20)));
    OK(ensure(h, h->d_obj, std::max<size_t>(out_bytes, 1 << 20))); // one device block for every output
    CK(cudaMemcpyAsync(h->dX.p, hin, in_bytes, cudaMemcpyHostToDevice, h->stream));
    char *dbase = (char *)h->d_obj.p;
    EvalOut out{};
    out.obj = (double *)(dbase + o_obj);
    out.count = count ? (long long *)(dbase + o_cnt) : nullptr;
    out.class_count = class_count ? (long long *)(dbase + o_cls) : nullptr;
    out.progressive = progressive ? (double *)(dbase + o_prg) : nullptr;
    out.feasible = feasible ? (unsigned char *)(dbase + o_fea) : nullptr;
    OK(launch_on_main(h, (const double *)h->dX.p, B, out, true));
    CK(cudaMemcpyAsync(hout, dbase, out_bytes, cudaMemcpyDeviceToHost, h->stream)); // ONE copy back
    CK(cudaStreamSynchronize(h->stream));
    const char *so = hout;
    memcpy(obj, so + o_obj, (size_t)B * 8);
    if (count) memcpy(count, so + o_cnt, (size_t)B * 8);
    if (class_count) memcpy(class_count, so + o_cls, (size_t)B * 8 * ncls);
    if (progressive) memcpy(progressive, so + o_prg, (size_t)B * 8);
    if (feasible) memcpy(feasible, so + o_fea, (size_t)B);
    return COV_OK;
}

// The poll winner of a host-path call (cov_argmin, cov_eval_batch_best): reduced on the device from the
// device-resident copy of the objectives and flags, 16 bytes come back per window.
struct BestReq {
    int barrier;
    double best = INFINITY;
    int64_t idx = -1;
};

// Packed candidates (cov_eval_batch_packed): `elem`-byte values that the device widens to doubles, value * g.
struct PackedIn {
    const void *Q;
    int pack;
    double g;
    size_t elem;
};
static void unpack_host(const PackedIn &pk, size_t n, double *out) // the same arithmetic as unpack_kernel, for poll sets
{
    switch (pk.pack) {
    case COV_PACK_F32:
        for (size_t k = 0; k < n; ++k) out[k] = (double)((const float *)pk.Q)[k];
        break;
    case COV_PACK_I32:
        for (size_t k = 0; k < n; ++k) out[k] = (double)((const int32_t *)pk.Q)[k] * pk.g;
        break;
    default:
        for (size_t k = 0; k < n; ++k) out[k] = (double)((const int16_t *)pk.Q)[k] * pk.g;
        break;
    }
}

static int eval_host(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count, uint8_t *feasible,
                     int64_t *class_count, double *progressive, BestReq *best = nullptr, const PackedIn *pk = nullptr)
{
    OK(check_ready(h, "cov_eval_batch"));
    if (B < 0 || (B > 0 && ((pk ? !pk->Q : !X) || (!obj && !best))))
        return fail(h, COV_ERR_INVALID, "cov_eval_batch: bad arguments");
    if (progressive && h->o.prog_which > h->o.N)
        return fail(h, COV_ERR_INVALID, "cov_eval_batch_ex: COV_OPT_PROGRESSIVE_INDEX names UAV " +
                                            std::to_string(h->o.prog_which) + " but N = " + std::to_string(h->o.N));
    if (B == 0) return COV_OK;
    recycle_spans(h);
    const int N = h->o.N;
    const int ncls = h->g.n_classes;
    const size_t row_bytes = (size_t)3 * N * 8;
    if (!best && (size_t)B * row_bytes <= (256u << 10) && (size_t)B * (25 + 8 * (size_t)ncls) <= (1536u << 10) && h->chunk == 0) {
        if (!pk) return eval_host_small(h, X, B, obj, count, feasible, class_count, progressive);
        std::vector<double> wide((size_t)B * 3 * N); // a poll set: widened on the host, then the poll path as it is
        unpack_host(*pk, wide.size(), wide.data());
        return eval_host_small(h, wide.data(), B, obj, count, feasible, class_count, progressive);
    }
#if COV_ZC_MID_LIMIT > 0
    // mid-size batches in pinned buffers: one launch that reads the candidates straight over PCIe (the
    // kernel's own unit prefetch overlaps transfer and compute) and writes the results back the same way
    if (!pk && !best && h->zero_copy_out && h->chunk == 0 && N <= 8 && (size_t)B * row_bytes <= (size_t)(COV_ZC_MID_LIMIT) && is_pinned_host(X) &&
        is_pinned_host(obj) && (!count || is_pinned_host(count)) && (!feasible || is_pinned_host(feasible)) &&
        (!class_count || is_pinned_host(class_count)) && (!progressive || is_pinned_host(progressive))) {
        const double *vx = (const double *)device_view((void *)X);
        EvalOut out{};
        out.obj = (double *)device_view(obj);
        out.count = count ? (long long *)device_view(count) : nullptr;
        out.feasible = feasible ? (unsigned char *)device_view(feasible) : nullptr;
        out.class_count = class_count ? (long long *)device_view(class_count) : nullptr;
        out.progressive = progressive ? (double *)device_view(progressive) : nullptr;
        if (vx && out.obj && (!count || out.count) && (!feasible || out.feasible) && (!class_count || out.class_count) &&
            (!progressive || out.progressive)) {
            OK(launch_on_main(h, vx, B, out, true));
            CK(cudaStreamSynchronize(h->stream));
            return COV_OK;
        }
    }
#endif
    // device window: at most ~1 GiB of candidates at a time
    const int64_t window = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)((1ull << 30) / row_bytes)));
    // slice size: ~16 MiB of candidates, a multiple of 32 candidates (keeps device slices 16-byte aligned).
    // Packed input is compute-bound, not PCIe-bound (5 UAVs, int16: 0.54 ms of copies beside 0.7 ms of kernels per
    // 1 M candidates), so fewer, larger slices win: ~32 MiB of widened candidates (measured on B200, 1 M x 5 UAVs:
    // int16 1.126 -> 1.098 ms per call, int32 / float32 1.51 -> 1.37 ms; tools/packed_exp.py)
    const size_t slice_bytes = pk ? (32ull << 20) : (16ull << 20);
    int64_t chunk = h->chunk > 0 ? h->chunk : std::max<int64_t>(1024, (int64_t)(slice_bytes / row_bytes) / 32 * 32);
    chunk = std::min(chunk, window);
    // (two experiments lost here, DESIGN.md 5: a quarter- and a half-size first slice so that the kernels start sooner,
    // 1.107 -> 1.125 ms; and no copy engine at all, the widening kernel reading the pinned buffer over PCIe beside the
    // previous slice's objective kernel, 1.107 -> 1.21 ms)
    OK(ensure(h, h->dX, (size_t)window * row_bytes));
    OK(ensure(h, h->d_obj, (size_t)window * 8));
    if (count) OK(ensure(h, h->d_count, (size_t)window * 8));
    if (feasible || best) OK(ensure(h, h->d_feas, (size_t)window));
    if (class_count) OK(ensure(h, h->d_clscnt, (size_t)window * 8 * ncls));
    if (progressive) OK(ensure(h, h->d_prog, (size_t)window * 8));
    // what crosses PCIe per candidate: the doubles themselves, or the packed values (widened on the device)
    const size_t in_row_bytes = pk ? (size_t)3 * N * pk->elem : row_bytes;
    const char *in_base = pk ? (const char *)pk->Q : (const char *)X;
    const bool in_pinned = is_pinned_host(in_base);
    if (!in_pinned)
        for (int k = 0; k < 2; ++k) {
            size_t cap = h->h_in_cap;
            OK(ensure_pinned(h, &h->h_in[k], &cap, (size_t)chunk * in_row_bytes));
            if (k == 1) h->h_in_cap = cap;
        }
    if (pk)
        for (int k = 0; k < 2; ++k) OK(ensure(h, h->d_raw[k], (size_t)chunk * in_row_bytes));
    // outputs that are not pinned are staged per window and copied out at the window's end
    const bool obj_p = !obj || is_pinned_host(obj), cnt_p = !count || is_pinned_host(count),
               fea_p = !feasible || is_pinned_host(feasible), cls_p = !class_count || is_pinned_host(class_count),
               prg_p = !progressive || is_pinned_host(progressive);
    size_t stage_out = 0;
    size_t off_obj = 0, off_cnt = 0, off_fea = 0, off_cls = 0, off_prg = 0;
    if (!obj_p) { off_obj = stage_out; stage_out += (size_t)window * 8; }
    if (!cnt_p) { off_cnt = stage_out; stage_out += (size_t)window * 8; }
    if (!cls_p) { off_cls = stage_out; stage_out += (size_t)window * 8 * ncls; }
    if (!prg_p) { off_prg = stage_out; stage_out += (size_t)window * 8; }
    if (!fea_p) { off_fea = stage_out; stage_out += (size_t)window; }
    if (stage_out) OK(ensure_pinned(h, &h->h_out, &h->h_out_cap, stage_out));
    char *so = (char *)h->h_out;

    cudaEvent_t ev_in = get_event(h), ev_k = get_event(h), ev_free[2] = {get_event(h), get_event(h)};
    cudaEvent_t ev_raw[2] = {get_event(h), get_event(h)}; // packed input: slot k of d_raw has been widened
    bool free_armed[2] = {false, false}, raw_armed[2] = {false, false};
    int rc = COV_OK;
    auto done = [&](int code) {
        h->ev_pool.push_back(ev_in);
        h->ev_pool.push_back(ev_k);
        h->ev_pool.push_back(ev_free[0]);
        h->ev_pool.push_back(ev_free[1]);
        h->ev_pool.push_back(ev_raw[0]);
        h->ev_pool.push_back(ev_raw[1]);
        return code;
    };
#define CKD(call)                                                         \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return done(fail_cuda(h, e_, #call));      \
    } while (0)
    // the copy streams must not start before earlier work on the main stream (grid/params) is done
    CKD(cudaEventRecord(ev_k, h->stream));
    CKD(cudaStreamWaitEvent(h->s_in, ev_k, 0));
    auto tr = [&](cudaStream_t s) {
        if (!h->trace) return;
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        h->trace_ev.push_back(e);
    };
    for (cudaEvent_t e : h->trace_ev) cudaEventDestroy(e);
    h->trace_ev.clear();
    h->trace_ms.clear();
    tr(h->s_in);
    for (int64_t w0 = 0; w0 < B; w0 += window) {
        const int64_t wn = std::min(window, B - w0);
        int slot = 0;
        int64_t cn = 0;
        for (int64_t c0 = 0; c0 < wn; c0 += cn, slot ^= 1) {
            const int64_t left = wn - c0;
            cn = std::min(chunk, left);
            const char *src = in_base + (size_t)(w0 + c0) * in_row_bytes;
            double *dst = (double *)h->dX.p + (size_t)c0 * 3 * N;
            if (!in_pinned) {
                if (free_armed[slot]) CKD(cudaEventSynchronize(ev_free[slot]));
                host_copy(h, h->h_in[slot], src, (size_t)cn * in_row_bytes);
                src = (const char *)h->h_in[slot];
            }
            if (pk && raw_armed[slot]) CKD(cudaStreamWaitEvent(h->s_in, ev_raw[slot], 0)); // slot widened: reusable
            CKD(cudaMemcpyAsync(pk ? h->d_raw[slot].p : (void *)dst, src, (size_t)cn * in_row_bytes, cudaMemcpyHostToDevice,
                                h->s_in));
            if (!in_pinned) {
                CKD(cudaEventRecord(ev_free[slot], h->s_in));
                free_armed[slot] = true;
            }
            CKD(cudaEventRecord(ev_in, h->s_in));
            tr(h->s_in);
            CKD(cudaStreamWaitEvent(h->stream, ev_in, 0));
            tr(h->stream);
            if (pk) {
                CKD(launch_unpack(h->d_raw[slot].p, pk->pack, pk->g, dst, (long long)cn * 3 * N, h->stream));
                h->launches += 1;
                CKD(cudaEventRecord(ev_raw[slot], h->stream));
                raw_armed[slot] = true;
            }
            EvalOut out{};
            const int64_t g0 = w0 + c0;
            bool zc = h->zero_copy_out != 0 && (obj != nullptr || !best); // (cov_argmin has no per-candidate outputs)
            if (zc) {
                // results go straight to pinned host memory (the caller's buffer, or the staging block): posted
                // PCIe writes of 17 B per candidate instead of three D2H copies per slice beside the H2D stream
                char *v_obj = (char *)device_view(obj_p ? (void *)obj : (void *)(so + off_obj));
                char *v_cnt = !count ? nullptr : (char *)device_view(cnt_p ? (void *)count : (void *)(so + off_cnt));
                char *v_fea = !feasible ? nullptr : (char *)device_view(fea_p ? (void *)feasible : (void *)(so + off_fea));
                char *v_cls = !class_count ? nullptr : (char *)device_view(cls_p ? (void *)class_count : (void *)(so + off_cls));
                char *v_prg = !progressive ? nullptr : (char *)device_view(prg_p ? (void *)progressive : (void *)(so + off_prg));
                if (!v_obj || (count && !v_cnt) || (feasible && !v_fea) || (class_count && !v_cls) || (progressive && !v_prg)) {
                    zc = false; // a buffer the device cannot address: copy back instead
                } else {
                    out.obj = (double *)v_obj + (obj_p ? g0 : c0);
                    out.count = !count ? nullptr : (long long *)v_cnt + (cnt_p ? g0 : c0);
                    out.feasible = !feasible ? nullptr : (unsigned char *)v_fea + (fea_p ? g0 : c0);
                    out.class_count = !class_count ? nullptr : (long long *)v_cls + (size_t)(cls_p ? g0 : c0) * ncls;
                    out.progressive = !progressive ? nullptr : (double *)v_prg + (prg_p ? g0 : c0);
                    if (best) { // results straight to the host as usual; the winner is reduced from device mirrors
                        out.obj_mirror = (double *)h->d_obj.p + c0;
                        out.feasible_mirror = (unsigned char *)h->d_feas.p + c0;
                    }
                }
            }
            if (!zc) {
                out.obj = (double *)h->d_obj.p + c0;
                out.count = count ? (long long *)h->d_count.p + c0 : nullptr;
                out.feasible = (feasible || best) ? (unsigned char *)h->d_feas.p + c0 : nullptr;
                out.class_count = class_count ? (long long *)h->d_clscnt.p + (size_t)c0 * ncls : nullptr;
                out.progressive = progressive ? (double *)h->d_prog.p + c0 : nullptr;
            }
            rc = launch_on_main(h, dst, cn, out, true);
            if (rc != COV_OK) return done(rc);
            CKD(cudaEventRecord(ev_k, h->stream));
            tr(h->stream);
            if (!zc) {
                CKD(cudaStreamWaitEvent(h->s_out, ev_k, 0));
                if (obj)
                    CKD(cudaMemcpyAsync(obj_p ? (void *)(obj + g0) : (void *)(so + off_obj + (size_t)c0 * 8), out.obj,
                                        (size_t)cn * 8, cudaMemcpyDeviceToHost, h->s_out));
                if (count)
                    CKD(cudaMemcpyAsync(cnt_p ? (void *)(count + g0) : (void *)(so + off_cnt + (size_t)c0 * 8),
                                        out.count, (size_t)cn * 8, cudaMemcpyDeviceToHost, h->s_out));
                if (feasible)
                    CKD(cudaMemcpyAsync(fea_p ? (void *)(feasible + g0) : (void *)(so + off_fea + (size_t)c0),
                                        out.feasible, (size_t)cn, cudaMemcpyDeviceToHost, h->s_out));
                if (class_count)
                    CKD(cudaMemcpyAsync(cls_p ? (void *)(class_count + (size_t)g0 * ncls)
                                              : (void *)(so + off_cls + (size_t)c0 * 8 * ncls),
                                        out.class_count, (size_t)cn * 8 * ncls, cudaMemcpyDeviceToHost, h->s_out));
                if (progressive)
                    CKD(cudaMemcpyAsync(prg_p ? (void *)(progressive + g0) : (void *)(so + off_prg + (size_t)c0 * 8),
                                        out.progressive, (size_t)cn * 8, cudaMemcpyDeviceToHost, h->s_out));
            }
            tr(zc ? h->stream : h->s_out);
        }
        // end of window: results home, device window reusable
        if (best) {
            CKD(launch_argmin((const double *)h->d_obj.p, (const unsigned char *)h->d_feas.p, wn, best->barrier,
                              (double *)h->argmin_obj.p, (long long *)h->argmin_idx.p, 1024, h->stream));
            h->launches += 2;
            double *ho = (double *)h->h_small + 40;
            long long *hi = (long long *)h->h_small + 41;
            CKD(cudaMemcpyAsync(ho, h->argmin_obj.p, 8, cudaMemcpyDeviceToHost, h->stream));
            CKD(cudaMemcpyAsync(hi, h->argmin_idx.p, 8, cudaMemcpyDeviceToHost, h->stream));
            CKD(cudaStreamSynchronize(h->stream));
            if (*hi >= 0 && (best->idx < 0 || *ho < best->best)) { // ties keep the earlier window's index
                best->best = *ho;
                best->idx = w0 + *hi;
            }
        }
        CKD(cudaStreamSynchronize(h->stream));
        CKD(cudaStreamSynchronize(h->s_out));
        if (!obj_p) host_copy(h, obj + w0, so + off_obj, (size_t)wn * 8);
        if (!cnt_p) host_copy(h, count + w0, so + off_cnt, (size_t)wn * 8);
        if (!fea_p) host_copy(h, feasible + w0, so + off_fea, (size_t)wn);
        if (!cls_p) memcpy(class_count + (size_t)w0 * ncls, so + off_cls, (size_t)wn * 8 * ncls);
        if (!prg_p) memcpy(progressive + w0, so + off_prg, (size_t)wn * 8);
    }
#undef CKD
    if (h->trace && h->trace_ev.size() > 1) {
        for (size_t k = 1; k < h->trace_ev.size(); ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, h->trace_ev[0], h->trace_ev[k]);
            h->trace_ms.push_back(t);
        }
    }
    return done(COV_OK);
}

extern "C" int cov_eval_batch(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count,
                              uint8_t *feasible)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    return eval_host(h, X, B, obj, count, feasible, nullptr, nullptr);
}

extern "C" int cov_eval_batch_ex(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count,
                                 uint8_t *feasible, int64_t *class_count, double *progressive)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    return eval_host(h, X, B, obj, count, feasible, class_count, progressive);
}

extern "C" int cov_eval_batch_best(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count,
                                   uint8_t *feasible, int32_t barrier, double *best_obj, int64_t *best_idx)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (!best_obj || !best_idx) return fail(h, COV_ERR_INVALID, "cov_eval_batch_best: NULL winner outputs");
    BestReq br{barrier};
    OK(eval_host(h, X, B, obj, count, feasible, nullptr, nullptr, &br));
    *best_obj = br.idx >= 0 ? br.best : INFINITY;
    *best_idx = br.idx;
    return COV_OK;
}

extern "C" int cov_eval_batch_packed(cov_handle *h, const void *Q, int32_t pack, double granularity, int64_t B, double *obj,
                                     int64_t *count, uint8_t *feasible, int32_t barrier, double *best_obj,
                                     int64_t *best_idx)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    const size_t elem = pack == COV_PACK_F32 || pack == COV_PACK_I32 ? 4 : pack == COV_PACK_I16 ? 2 : 0;
    if (!elem) return fail(h, COV_ERR_INVALID, "cov_eval_batch_packed: pack must be COV_PACK_F32, _I32 or _I16");
    if (pack != COV_PACK_F32 && !(granularity > 0 && std::isfinite(granularity)))
        return fail(h, COV_ERR_INVALID, "cov_eval_batch_packed: granularity must be a positive finite number");
    if ((best_obj == nullptr) != (best_idx == nullptr))
        return fail(h, COV_ERR_INVALID, "cov_eval_batch_packed: best_obj and best_idx go together");
    PackedIn pk{Q, pack, granularity, elem};
    if (!best_obj) return eval_host(h, nullptr, B, obj, count, feasible, nullptr, nullptr, nullptr, &pk);
    BestReq br{barrier};
    OK(eval_host(h, nullptr, B, obj, count, feasible, nullptr, nullptr, &br, &pk));
    *best_obj = br.idx >= 0 ? br.best : INFINITY;
    *best_idx = br.idx;
    return COV_OK;
}

// One candidate through pinned scratch: copy in, one launch, copy out, one synchronisation.
extern "C" int cov_eval_one(cov_handle *h, const double *x, double *obj)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    OK(check_ready(h, "cov_eval_one"));
    if (!x || !obj) return fail(h, COV_ERR_INVALID, "cov_eval_one: NULL argument");
    recycle_spans(h);
    const int N = h->o.N;
    const size_t bytes = (size_t)3 * N * 8;
    char *hx = (char *)h->h_small + 4096; // pinned scratch for one candidate (3 * kMaxUavs doubles)
    memcpy(hx, x, bytes);
    if (h->zero_copy_out) {
        // no copy engines: the kernel reads the candidate from, and writes the objective to, pinned host memory
        double *hres = (double *)h->h_small + 32;
        const double *vx = (const double *)device_view(hx);
        double *vres = (double *)device_view(hres);
        if (vx && vres) {
            EvalOut out{};
            out.obj = vres;
            OK(launch_on_main(h, vx, 1, out, false));
            CK(cudaStreamSynchronize(h->stream));
            *obj = *hres;
            return COV_OK;
        }
    }
    OK(ensure(h, h->dX, bytes));
    OK(ensure(h, h->d_obj, 8));
    CK(cudaMemcpyAsync(h->dX.p, hx, bytes, cudaMemcpyHostToDevice, h->stream));
    EvalOut out{};
    out.obj = (double *)h->d_obj.p;
    OK(launch_on_main(h, (const double *)h->dX.p, 1, out, false));
    double *hres = (double *)h->h_small + 32;
    CK(cudaMemcpyAsync(hres, h->d_obj.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *obj = *hres;
    return COV_OK;
}


// ------------------------------------------------------------------------------------------
// the batched MADS poll driver in native code: the caller of the objective.
// Same settings as the reference's TDM_STATIC_opt.optimize (src/TDM_STATIC_opt.jl:118-222: iteration
// limit, granularity on every variable, extreme barrier from the fused constraints, feasible incumbent
// else the start point) and the same published algorithm as mads.py (granular mesh of Audet, Le Digabel &
// Tribes 2019: poll size a * 10^b with a in {1, 2, 5}, mesh size max(10^(b - |b - b0|), granularity),
// 2n directions from a random Householder matrix rounded onto the mesh, complete polling); every poll
// set is ONE launch through the small-batch path.  DirectSearch.jl's own iterates are not reproducible
// (unseeded LTMADS), so trajectories are not a parity target; objective values are.
// ------------------------------------------------------------------------------------------
namespace {
struct Rng { // splitmix64 + Box-Muller: self-contained, seedable
    uint64_t s;
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double u01() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    double normal() { return std::sqrt(-2.0 * std::log(u01())) * std::cos(6.283185307179586 * u01()); }
};
struct MeshVar {
    double a;
    int b, b0;
};
} // namespace

extern "C" int cov_mads_solve(cov_handle *h, const double *x0, int64_t n_iter, double granularity, uint64_t seed,
                              double *x_out, double *obj_out, int64_t *stats)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    OK(check_ready(h, "cov_mads_solve"));
    if (!x0 || !x_out || n_iter < 0 || !(granularity >= 0)) return fail(h, COV_ERR_INVALID, "cov_mads_solve: bad arguments");
    const int n = 3 * h->o.N;
    const double g = granularity;
    auto snap = [g](double v) { return g > 0 ? std::nearbyint(v / g) * g : v; };
    std::unordered_map<std::string, double> cache;
    int64_t evaluations = 0, batches = 0, successes = 0, iterations = 0;
    std::vector<double> P, vals;
    std::vector<uint8_t> feas;
    auto key_of = [n](const double *p) { return std::string(reinterpret_cast<const char *>(p), (size_t)n * 8); };
    // objective with the extreme barrier for the rows of P that are not cached yet
    auto evaluate = [&](const std::vector<double> &pts, std::vector<double> &f) -> int {
        const int64_t m = (int64_t)pts.size() / n;
        f.assign((size_t)m, 0.0);
        std::vector<int64_t> todo;
        std::vector<double> Q;
        for (int64_t k = 0; k < m; ++k) {
            auto it = cache.find(key_of(&pts[(size_t)k * n]));
            if (it != cache.end()) f[(size_t)k] = it->second;
            else {
                todo.push_back(k);
                Q.insert(Q.end(), &pts[(size_t)k * n], &pts[(size_t)k * n] + n);
            }
        }
        if (!todo.empty()) {
            vals.assign(todo.size(), 0.0);
            feas.assign(todo.size(), 1);
            recycle_spans(h);
            OK(eval_host(h, Q.data(), (int64_t)todo.size(), vals.data(), nullptr, feas.data(), nullptr, nullptr));
            evaluations += (int64_t)todo.size();
            batches += 1;
            for (size_t q = 0; q < todo.size(); ++q) {
                const double v = (feas[q] && vals[q] == vals[q]) ? vals[q] : INFINITY;
                cache[key_of(&Q[q * n])] = v;
                f[(size_t)todo[q]] = v;
            }
        }
        return COV_OK;
    };
    std::vector<double> x(x0, x0 + n), f;
    OK(evaluate(x, f));
    double fx = f[0];
    bool feasible_found = std::isfinite(fx);
    std::vector<MeshVar> mesh((size_t)n);
    for (int i = 0; i < n; ++i) { // initial poll size: about |x0_i| / 10 (at least 1), on the {1, 2, 5} x 10^b ladder
        const double target = std::max(std::max(std::fabs(x[i]) / 10.0, 1.0), g);
        const int b = (int)std::floor(std::log10(target));
        const double mant = target / std::pow(10.0, b);
        mesh[i] = {mant < 2 ? 1.0 : (mant < 5 ? 2.0 : 5.0), b, b};
    }
    Rng rng{seed * 0x9E3779B97F4A7C15ull + 0x1234567ull};
    std::vector<double> v((size_t)n), H((size_t)n * n), delta((size_t)n), rho((size_t)n);
    for (int64_t it = 0; it < n_iter; ++it) {
        ++iterations;
        for (int i = 0; i < n; ++i) {
            const double Delta = std::max(mesh[i].a * std::pow(10.0, mesh[i].b), g);
            delta[i] = std::max(std::pow(10.0, mesh[i].b - std::abs(mesh[i].b - mesh[i].b0)), g);
            rho[i] = std::max(std::nearbyint(Delta / delta[i]), 1.0);
        }
        double nrm = 0;
        for (int i = 0; i < n; ++i) {
            v[i] = rng.normal();
            nrm += v[i] * v[i];
        }
        nrm = std::sqrt(nrm);
        for (int i = 0; i < n; ++i) v[i] /= nrm;
        P.clear();
        for (int sgn = 1; sgn >= -1; sgn -= 2)
            for (int j = 0; j < n; ++j) { // column j of I - 2 v v^T, scaled to the poll-to-mesh ratio, rounded
                double hmax = 0;
                for (int i = 0; i < n; ++i) {
                    H[i] = (i == j ? 1.0 : 0.0) - 2.0 * v[i] * v[j];
                    hmax = std::max(hmax, std::fabs(H[i]));
                }
                bool any = false, moved = false;
                const size_t at = P.size();
                P.resize(at + n);
                for (int i = 0; i < n; ++i) {
                    const double d = std::nearbyint(rho[i] * H[i] / hmax);
                    any |= d != 0;
                    P[at + i] = snap(x[i] + sgn * d * delta[i]);
                    moved |= P[at + i] != x[i];
                }
                if (!any || !moved) P.resize(at);
            }
        // drop exact duplicates (keep the first)
        {
            std::unordered_map<std::string, int> seen;
            std::vector<double> U;
            for (size_t k = 0; k * n < P.size(); ++k)
                if (seen.emplace(key_of(&P[k * n]), 1).second) U.insert(U.end(), &P[k * n], &P[k * n] + n);
            P.swap(U);
        }
        double best = INFINITY;
        size_t kb = 0;
        if (!P.empty()) {
            OK(evaluate(P, f));
            for (size_t k = 0; k < f.size(); ++k)
                if (f[k] < best) {
                    best = f[k];
                    kb = k;
                }
        }
        if (best < fx || (!feasible_found && std::isfinite(best))) {
            std::copy(&P[kb * n], &P[kb * n] + n, x.begin());
            fx = best;
            feasible_found = true;
            ++successes;
            for (auto &m : mesh) { // enlarge
                if (m.a == 1.0) m.a = 2.0;
                else if (m.a == 2.0) m.a = 5.0;
                else {
                    m.a = 1.0;
                    ++m.b;
                }
            }
        } else { // refine; stop when every poll size sits at its granularity
            bool moved = false;
            for (auto &m : mesh) {
                if (g > 0 && m.a * std::pow(10.0, m.b) <= g) continue;
                if (m.a == 1.0) {
                    m.a = 5.0;
                    --m.b;
                } else if (m.a == 2.0) m.a = 1.0;
                else m.a = 2.0;
                moved = true;
            }
            if (!moved) break;
        }
    }
    const double *res = feasible_found ? x.data() : x0; // p.x if a feasible incumbent exists, else the start (p.i)
    std::copy(res, res + n, x_out);
    if (obj_out) *obj_out = fx;
    if (stats) {
        stats[0] = iterations;
        stats[1] = evaluations;
        stats[2] = batches;
        stats[3] = successes;
    }
    return COV_OK;
}

extern "C" int cov_argmin(cov_handle *h, const double *X, int64_t B, int32_t barrier, double *best_obj,
                          int64_t *best_idx)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    OK(check_ready(h, "cov_argmin"));
    if (B < 0 || (B > 0 && !X) || !best_obj || !best_idx) return fail(h, COV_ERR_INVALID, "cov_argmin: bad arguments");
    double best = INFINITY;
    int64_t bidx = -1;
    const int N = h->o.N;
    const size_t row_bytes = (size_t)3 * N * 8;
    if ((size_t)B * row_bytes > (4u << 20)) { // big sets: the sliced pipeline (H2D of slice k+1 beside the kernel of slice k)
        BestReq br{barrier};
        OK(eval_host(h, X, B, nullptr, nullptr, nullptr, nullptr, nullptr, &br));
        *best_obj = br.idx >= 0 ? br.best : INFINITY;
        *best_idx = br.idx;
        return COV_OK;
    }
    recycle_spans(h);
    const int64_t window = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(B, 1), (int64_t)((1ull << 30) / row_bytes)));
    OK(ensure(h, h->dX, (size_t)window * row_bytes));
    OK(ensure(h, h->d_obj, (size_t)window * 8));
    OK(ensure(h, h->d_feas, (size_t)window));
    for (int64_t w0 = 0; w0 < B; w0 += window) {
        const int64_t wn = std::min(window, B - w0);
        CK(cudaMemcpyAsync(h->dX.p, X + (size_t)w0 * 3 * N, (size_t)wn * row_bytes, cudaMemcpyHostToDevice, h->stream));
        EvalOut out{};
        out.obj = (double *)h->d_obj.p;
        out.feasible = (unsigned char *)h->d_feas.p;
        OK(launch_on_main(h, (const double *)h->dX.p, wn, out, true));
        CK(launch_argmin(out.obj, out.feasible, wn, barrier, (double *)h->argmin_obj.p, (long long *)h->argmin_idx.p,
                         1024, h->stream));
        h->launches += 2;
        double *ho = (double *)h->h_small + 40;
        long long *hi = (long long *)h->h_small + 41;
        CK(cudaMemcpyAsync(ho, h->argmin_obj.p, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(hi, h->argmin_idx.p, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (*hi >= 0 && (bidx < 0 || *ho < best)) { // ties keep the earlier window's index
            best = *ho;
            bidx = w0 + *hi;
        }
    }
    *best_obj = bidx >= 0 ? best : INFINITY;
    *best_idx = bidx;
    return COV_OK;
}


// ------------------------------------------------------------------------------------------
// continuous variant: exact union area of the discs (no grid involved)
// ------------------------------------------------------------------------------------------
extern "C" int cov_union_area_batch(cov_handle *h, const double *X, int64_t B, int64_t N, double *area)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (N < 1 || N > kMaxUavs) return fail(h, COV_ERR_LIMIT, "cov_union_area_batch: N must be 1..1024");
    if (B < 0 || (B > 0 && (!X || !area))) return fail(h, COV_ERR_INVALID, "cov_union_area_batch: bad arguments");
    const size_t row_bytes = (size_t)3 * N * 8;
    const int64_t window = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(B, 1), (int64_t)((256ull << 20) / row_bytes)));
    OK(ensure(h, h->dX, (size_t)window * row_bytes));
    OK(ensure(h, h->d_obj, (size_t)window * 8));
    for (int64_t w0 = 0; w0 < B; w0 += window) {
        const int64_t wn = std::min(window, B - w0);
        CK(cudaMemcpyAsync(h->dX.p, X + (size_t)w0 * 3 * N, (size_t)wn * row_bytes, cudaMemcpyHostToDevice, h->stream));
        CK(launch_union_area((const double *)h->dX.p, (long long)wn, (int)N, (double *)h->d_obj.p, h->stream));
        h->launches += 1;
        CK(cudaMemcpyAsync(area + w0, h->d_obj.p, (size_t)wn * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return COV_OK;
}

// ------------------------------------------------------------------------------------------
// streams, memory, timing
// ------------------------------------------------------------------------------------------
extern "C" int cov_sync(cov_handle *h)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    CK(cudaStreamSynchronize(h->s_in));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->s_out));
    return COV_OK;
}
extern "C" void *cov_stream(cov_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int cov_set_stream(cov_handle *h, void *stream)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    CK(cudaStreamSynchronize(h->stream));
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return COV_OK;
}
extern "C" int cov_host_alloc(cov_handle *h, int64_t bytes, void **out)
{
    if (!h || !out || bytes < 0) return fail(h, COV_ERR_INVALID, "cov_host_alloc: bad arguments");
    DeviceGuard dg(h->device);
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, COV_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    }
    return COV_OK;
}
extern "C" int cov_host_free(cov_handle *h, void *p)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (p) CK(cudaFreeHost(p));
    return COV_OK;
}
extern "C" int cov_device_alloc(cov_handle *h, int64_t bytes, void **out)
{
    if (!h || !out || bytes < 0) return fail(h, COV_ERR_INVALID, "cov_device_alloc: bad arguments");
    DeviceGuard dg(h->device);
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, COV_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    return COV_OK;
}
extern "C" int cov_device_free(cov_handle *h, void *p)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (p) CK(cudaFree(p));
    return COV_OK;
}
extern "C" int cov_memcpy_h2d(cov_handle *h, void *dst, const void *src, int64_t bytes)
{
    if (!h || bytes < 0) return fail(h, COV_ERR_INVALID, "cov_memcpy_h2d: bad arguments");
    DeviceGuard dg(h->device);
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, h->stream));
    return COV_OK;
}
extern "C" int cov_memcpy_d2h(cov_handle *h, void *dst, const void *src, int64_t bytes)
{
    if (!h || bytes < 0) return fail(h, COV_ERR_INVALID, "cov_memcpy_d2h: bad arguments");
    DeviceGuard dg(h->device);
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
    return COV_OK;
}
extern "C" int64_t cov_launch_count(const cov_handle *h) { return h ? h->launches : 0; }

extern "C" int64_t cov_get_trace(const cov_handle *h, double *ms, int64_t cap)
{
    if (!h) return 0;
    const int64_t n = (int64_t)h->trace_ms.size();
    for (int64_t k = 0; k < n && k < cap && ms; ++k) ms[k] = h->trace_ms[k];
    return n;
}

extern "C" int cov_last_kernel_ms(cov_handle *h, double *ms)
{
    if (!h || !ms) return fail(h, COV_ERR_INVALID, "cov_last_kernel_ms: bad arguments");
    DeviceGuard dg(h->device);
    if (h->last_call_ms_cache < 0 || h->kernel_spans.size() > h->last_call_begin) {
        const double before = h->last_call_ms_cache < 0 ? 0 : h->last_call_ms_cache;
        CK(drain_spans(h));
        h->last_call_ms_cache += before;
    }
    *ms = h->last_call_ms_cache;
    return COV_OK;
}

extern "C" int cov_last_launch(const cov_handle *h, cov_launch_info *out)
{
    if (!h || !out) return COV_ERR_INVALID;
    if (h->last_info.block == 0) return COV_ERR_STATE;
    out->kernel = h->last_info.kernel;
    out->grid = h->last_info.grid;
    out->block = h->last_info.block;
    out->smem_bytes = h->last_info.smem_bytes;
    out->band_rows = h->last_info.band_rows;
    out->planes_in_smem = h->last_info.planes_in_smem;
    out->multi = h->last_info.multi;
    out->chunk = h->last_info.chunk;
    out->max_warps = h->last_info.max_warps;
    out->plane_mode = h->last_info.plane_mode;
    return COV_OK;
}

extern "C" int cov_kernel_time_total(cov_handle *h, double *ms, int64_t *launches)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    const double keep = h->last_call_ms_cache;
    const bool pending = h->kernel_spans.size() > h->last_call_begin;
    CK(drain_spans(h));
    if (keep >= 0 && pending) h->last_call_ms_cache += keep;
    else if (keep >= 0 && !pending) h->last_call_ms_cache = keep;
    if (ms) *ms = h->drained_ms;
    if (launches) *launches = h->drained_launches;
    return COV_OK;
}

extern "C" int cov_generate_candidates(cov_handle *h, double *dX, int64_t B, int64_t N, uint64_t seed,
                                       int64_t first_index, double lx, double ly, double h_min, double h_max,
                                       double tan_half_fov)
{
    if (!h) return fail(nullptr, COV_ERR_INVALID, "NULL handle");
    DeviceGuard dg(h->device);
    if (B < 0 || N < 1 || N > kMaxUavs || (B > 0 && !dX)) return fail(h, COV_ERR_INVALID, "cov_generate_candidates: bad arguments");
    CK(launch_generate(dX, (long long)B, (int)N, seed, (long long)first_index, lx, ly, h_min, h_max, tan_half_fov,
                       h->stream));
    h->launches += 1;
    return COV_OK;
}

// ------------------------------------------------------------------------------------------
// several GPUs of one box: contiguous candidate slices, one host thread per device
// ------------------------------------------------------------------------------------------
struct cov_multi {
    std::vector<cov_handle *> hs;
    std::string err;
};

extern "C" int cov_multi_create(const int *devices, int n, cov_multi **out)
{
    if (!out || n < 1 || !devices) return fail(nullptr, COV_ERR_INVALID, "cov_multi_create: bad arguments");
    *out = nullptr;
    cov_multi *m = new cov_multi();
    for (int k = 0; k < n; ++k) {
        cov_handle *h = nullptr;
        int rc = cov_create(devices[k], &h);
        if (rc != COV_OK) {
            for (cov_handle *q : m->hs) cov_destroy(q);
            delete m;
            return rc;
        }
        m->hs.push_back(h);
    }
    *out = m;
    return COV_OK;
}
extern "C" void cov_multi_destroy(cov_multi *m)
{
    if (!m) return;
    for (cov_handle *h : m->hs) cov_destroy(h);
    delete m;
}
extern "C" const char *cov_multi_last_error(const cov_multi *m) { return m ? m->err.c_str() : g_err_nohandle.c_str(); }
extern "C" int cov_multi_size(const cov_multi *m) { return m ? (int)m->hs.size() : 0; }
extern "C" cov_handle *cov_multi_handle(cov_multi *m, int k)
{
    return (m && k >= 0 && k < (int)m->hs.size()) ? m->hs[k] : nullptr;
}

static void shard(int64_t B, int G, int k, int64_t &b0, int64_t &bn)
{
    const int64_t per = (B + G - 1) / G; // ceil(B/G), SURVEY.md 8e
    b0 = std::min<int64_t>(B, per * k);
    bn = std::min<int64_t>(B, b0 + per) - b0;
}

extern "C" int cov_multi_eval_batch(cov_multi *m, const double *X, int64_t B, double *obj, int64_t *count,
                                    uint8_t *feasible)
{
    if (!m) return fail(nullptr, COV_ERR_INVALID, "NULL multi handle");
    if (B < 0 || (B > 0 && (!X || !obj))) {
        m->err = "cov_multi_eval_batch: bad arguments";
        return COV_ERR_INVALID;
    }
    const int G = (int)m->hs.size();
    std::vector<int> rcs((size_t)G, COV_OK);
    std::vector<std::thread> th;
    for (int k = 0; k < G; ++k)
        th.emplace_back([&, k]() {
            int64_t b0, bn;
            shard(B, G, k, b0, bn);
            if (bn <= 0) return;
            cov_handle *h = m->hs[k];
            const int N = h->have_params ? h->o.N : 0;
            rcs[k] = cov_eval_batch(h, X + (size_t)b0 * 3 * N, bn, obj + b0, count ? count + b0 : nullptr,
                                    feasible ? feasible + b0 : nullptr);
        });
    for (auto &t : th) t.join();
    for (int k = 0; k < G; ++k)
        if (rcs[k] != COV_OK) {
            m->err = "device shard " + std::to_string(k) + ": " + m->hs[k]->err;
            return rcs[k];
        }
    return COV_OK;
}

extern "C" int cov_multi_argmin(cov_multi *m, const double *X, int64_t B, int32_t barrier, double *best_obj,
                                int64_t *best_idx)
{
    if (!m) return fail(nullptr, COV_ERR_INVALID, "NULL multi handle");
    if (B < 0 || (B > 0 && !X) || !best_obj || !best_idx) {
        m->err = "cov_multi_argmin: bad arguments";
        return COV_ERR_INVALID;
    }
    const int G = (int)m->hs.size();
    std::vector<int> rcs((size_t)G, COV_OK);
    std::vector<double> bo((size_t)G, INFINITY);
    std::vector<int64_t> bi((size_t)G, -1);
    std::vector<std::thread> th;
    for (int k = 0; k < G; ++k)
        th.emplace_back([&, k]() {
            int64_t b0, bn;
            shard(B, G, k, b0, bn);
            if (bn <= 0) return;
            cov_handle *h = m->hs[k];
            const int N = h->have_params ? h->o.N : 0;
            rcs[k] = cov_argmin(h, X + (size_t)b0 * 3 * N, bn, barrier, &bo[k], &bi[k]);
            if (rcs[k] == COV_OK && bi[k] >= 0) bi[k] += b0;
        });
    for (auto &t : th) t.join();
    double best = INFINITY;
    int64_t idx = -1;
    for (int k = 0; k < G; ++k) {
        if (rcs[k] != COV_OK) {
            m->err = "device shard " + std::to_string(k) + ": " + m->hs[k]->err;
            return rcs[k];
        }
        if (bi[k] >= 0 && (idx < 0 || bo[k] < best)) { // 16-byte (min, index) pair per GPU, host gather
            best = bo[k];
            idx = bi[k];
        }
    }
    *best_obj = best;
    *best_idx = idx;
    return COV_OK;
}
