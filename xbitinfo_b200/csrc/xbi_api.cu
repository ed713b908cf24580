// C ABI of libxbitinfo_b200.so (declared in include/xbitinfo_b200.h).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <sys/mman.h>

#include "xbi_common.cuh"
#include "xbi_internal.h"

namespace xbi {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<unsigned long long> g_fast_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void note_fast_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_fast_launches.fetch_add(1, std::memory_order_relaxed);
}

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
static int fail_cuda(cudaError_t e, const char* where) {
    g_last_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return XBI_ERR_CUDA_BASE + (int)e;
}
#define XBI_CUDA(call, where)                          \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) return fail_cuda(e__, where); \
    } while (0)

static int make_job(CountJob& job, const void* x, int dtype, int64_t outer, int64_t n, int64_t inner,
                    unsigned flags, uint64_t mask_bits) {
    const int bytes = dtype_bytes(dtype);
    if (bytes == 0) return fail(XBI_ERR_ARG, "unknown dtype code");
    if (outer < 0 || n < 0 || inner < 0) return fail(XBI_ERR_ARG, "negative extent");
    if ((flags & XBI_MASK_NAN) && (flags & XBI_MASK_VALUE))
        return fail(XBI_ERR_ARG, "XBI_MASK_NAN and XBI_MASK_VALUE are mutually exclusive");
    if ((flags & (XBI_SIGNED_EXPONENT | XBI_MASK_NAN)) && !dtype_is_float(dtype))
        return fail(XBI_ERR_ARG, "XBI_SIGNED_EXPONENT / XBI_MASK_NAN need a floating-point dtype");
    if (reinterpret_cast<uintptr_t>(x) % bytes != 0) return fail(XBI_ERR_ARG, "data pointer is not element-aligned");
    job.x = x;
    job.dtype = dtype;
    job.transform = (flags & XBI_SIGNED_EXPONENT) != 0;
    job.mask_mode = (flags & XBI_MASK_NAN) ? kMaskNaN : (flags & XBI_MASK_VALUE) ? kMaskValue : kMaskNone;
    job.mask_bits = bytes == 8 ? mask_bits : (mask_bits & ((1ull << (8 * bytes)) - 1ull));
    job.outer = outer;
    job.n = n;
    job.inner = inner;
    return XBI_OK;
}

// ---------------------------------------------------------------------------------
// optional timing of the dominant (fast-path) kernel: event pairs around its launch
// ---------------------------------------------------------------------------------
struct Profile {
    static constexpr int kMax = 4096;
    bool enabled = false;
    int used = 0;
    cudaEvent_t ev[2 * kMax];
    int made = 0;
    unsigned long long bytes = 0;  // algorithmic bytes covered by the timed launches
};
static thread_local Profile g_prof;

static cudaEvent_t prof_event(int idx) {
    while (g_prof.made <= idx) {
        cudaEventCreate(&g_prof.ev[g_prof.made]);
        ++g_prof.made;
    }
    return g_prof.ev[idx];
}

// what becomes of the table of one counting job (the tail of count_pairs_impl: the finish kernel)
struct Finish {
    unsigned long long* counts = nullptr;  // nbits*4
    bool accumulate = true;                // counts += table; false: counts = table
    double* mi = nullptr;                  // K2 in the same launch when set
    int set_zero = 0;
    double z = 0.0;
    PeerExchange px = {};                  // world > 1: all-reduce over peer memory before counts / K2
};

static bool allow_generic_only() {
    static const bool on = [] {
        const char* env = std::getenv("XBI_ALLOW_GENERIC");
        return env && env[0] == '1';
    }();
    return on;
}

static FinishJob finish_job(const CountJob& job, const Workspace* ws, const Finish& fin) {
    FinishJob J = {};  // (no leftover ranges)
    J.x = job.x;
    J.n = job.n;
    J.inner = job.inner > 0 ? job.inner : 1;
    J.mask_lo = (uint32_t)(job.mask_bits & 0xFFFFFFFFu);
    J.mask_hi = (uint32_t)(job.mask_bits >> 32);
    J.ws = ws;
    J.counts = fin.counts;
    J.accumulate = fin.accumulate ? 1 : 0;
    J.mi = fin.mi;
    J.set_zero = fin.set_zero;
    J.z = fin.z;
    J.px = fin.px;
    return J;
}

// K1 of one job into the workspace, then ONE finish launch: fringe pairs + combine (+ exchange) (+ K2).
static int count_pairs_impl(const CountJob& job, unsigned flags, Workspace* ws, cudaStream_t stream, const Finish& fin) {
    const int64_t numel = job.outer * job.n * job.inner;
    const bool nothing = numel == 0 || job.n < 2;  // no pairs along the axis
    if (nothing && fin.accumulate && !fin.mi && fin.px.world <= 1) return XBI_OK;  // nothing to add
    XBI_CUDA(cudaMemsetAsync(ws, 0, sizeof(Workspace), stream), "cudaMemsetAsync(workspace)");
    FinishJob J = finish_job(job, ws, fin);
    if (!nothing) {
        const int64_t npairs_flat = numel - job.inner;
        int64_t dlo = 0, dhi = 0;
        bool edges_done = false;
        RowsLeftovers left;  // what the TMA kernel's producer warps counted on top of [dlo, dhi)
        if (!(flags & XBI_FORCE_GENERIC)) {
            const bool timed = g_prof.enabled && g_prof.used < Profile::kMax;
            // masked modes: sample the density of validity boundaries; the sparse and the dense flavour of
            // the fast kernels both read it and one of them returns at once
            if (flags & XBI_FORCE_DENSE) {
                static const unsigned long long kForceDense[2] = {1ull, 0ull};  // "every sampled pair is a boundary"
                XBI_CUDA(cudaMemcpyAsync(ws->plus.pad, kForceDense, sizeof(kForceDense), cudaMemcpyHostToDevice, stream),
                         "cudaMemcpyAsync(flavour)");
            } else if (!(flags & XBI_FORCE_SPARSE)) {
                XBI_CUDA(launch_mask_density(job, npairs_flat, ws, stream), "mask density sample");
            }
            if (timed) XBI_CUDA(cudaEventRecord(prof_event(2 * g_prof.used), stream), "cudaEventRecord");
            cudaError_t fast;
            if (job.inner == 1)
                fast = launch_pairs_tma(job, npairs_flat, ws, stream, &dlo, &dhi, &edges_done, &left);
            else
                fast = launch_pairs_cols(job, ws, stream, &dlo, &dhi, &edges_done);
            if (fast == cudaErrorSymbolNotFound) {
                // the driver does not hand out cuTensorMapEncodeTiled: the TMA kernel cannot run.  Never silently:
                // the generic kernel is 4-10x slower.
                cudaGetLastError();
                if (!allow_generic_only())
                    return fail(XBI_ERR_UNSUPPORTED,
                                "cuTensorMapEncodeTiled could not be resolved from the CUDA driver: the TMA fast path is "
                                "unavailable (set XBI_ALLOW_GENERIC=1 to run on the 4-10x slower generic kernel)");
                static std::atomic<bool> warned{false};
                if (!warned.exchange(true))
                    std::fprintf(stderr, "xbitinfo_b200: TMA fast path unavailable, counting on the generic kernel (XBI_ALLOW_GENERIC=1)\n");
                dlo = dhi = 0;
                left = RowsLeftovers{};
            } else if (fast != cudaSuccess) {
                return fail_cuda(fast, job.inner == 1 ? "pair counting (TMA rows)" : "pair counting (column walkers)");
            }
            if (timed) {
                XBI_CUDA(cudaEventRecord(prof_event(2 * g_prof.used + 1), stream), "cudaEventRecord");
                if (dhi > dlo) {
                    g_prof.bytes += (unsigned long long)(dhi - dlo) * dtype_bytes(job.dtype);
                    ++g_prof.used;
                }
            }
        }
        if (dhi <= dlo) { dlo = 0; dhi = 0; edges_done = false; left = RowsLeftovers{}; }
        // what the fast kernel left: flat pairs [0, dlo) and [dhi, npairs_flat) ...
        J.range_lo[0] = 0;   J.range_cnt[0] = dlo;
        J.range_lo[1] = dhi; J.range_cnt[1] = npairs_flat - dhi;
        if (left.flat) J.range_cnt[0] = J.range_cnt[1] = 0;  // (the TMA kernel's producer warps took them along)
        if (J.range_cnt[0] + J.range_cnt[1] > kFinishMaxFlat) {
            XBI_CUDA(launch_pairs_generic_flat(job, 0, dlo, &ws->plus, stream), "pair counting (generic, head)");
            XBI_CUDA(launch_pairs_generic_flat(job, dhi, npairs_flat, &ws->plus, stream), "pair counting (generic, tail)");
            J.range_cnt[0] = J.range_cnt[1] = 0;
        }
        // ... and the block-straddling pairs it did not remove itself
        const int64_t nedges = (job.outer - 1) * job.inner;
        if (edges_done && job.inner != 1) {
            // column walkers: every block-boundary pair has been subtracted by k_block_edges
        } else if (edges_done) {
            // inner == 1: edge t sits at flat index t*n + n-1; those inside [dlo, dhi) are done
            J.range_lo[2] = 0;          J.range_cnt[2] = dlo / job.n;
            J.range_lo[3] = dhi / job.n; J.range_cnt[3] = nedges - dhi / job.n;
        } else {
            J.range_lo[2] = 0; J.range_cnt[2] = nedges;
        }
        for (int r = 2; r < 4; ++r)
            if (J.range_cnt[r] < 0 || left.edges) J.range_cnt[r] = 0;
        if (J.range_cnt[2] + J.range_cnt[3] > kFinishMaxEdges) {
            for (int r = 2; r < 4; ++r)
                XBI_CUDA(launch_pairs_generic_edges(job, J.range_lo[r], J.range_lo[r] + J.range_cnt[r], &ws->minus, stream),
                         "pair counting (edge pairs)");
            J.range_cnt[2] = J.range_cnt[3] = 0;
        }
    }
    XBI_CUDA(launch_finish(job, J, stream), "finish (fringe + combine)");
    return XBI_OK;
}

// ---------------------------------------------------------------------------------
// host-buffer path: per-thread staging context
// ---------------------------------------------------------------------------------
struct HostCtx {
    int device = -1;
    size_t buf_bytes = 0;  // current size of each staging buffer
    size_t buf_cap = 0;    // the most they may grow to (XBI_STAGING_MB)
    void* buf[2] = {nullptr, nullptr};
    cudaStream_t stream[2] = {nullptr, nullptr};
    Workspace* ws[2] = {nullptr, nullptr};
    unsigned long long* counts = nullptr;  // 64*4
    double* mi = nullptr;                  // 64
    unsigned long long* counts_multi = nullptr;  // XBI_MAX_DIMS tables (xbi_bitinformation_host_multi)
    double* mi_multi = nullptr;
    // pinned staging for pageable caller memory (allocated on first need): two slots per direction
    static constexpr size_t kPinBytes = (size_t)32 << 20;
    void* pin_in[2] = {nullptr, nullptr};
    void* pin_out[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool in_busy[2] = {false, false};
    int in_next = 0;
    void release() {
        if (device < 0) return;
        cudaSetDevice(device);
        for (int i = 0; i < 2; ++i) {
            if (stream[i]) { cudaStreamSynchronize(stream[i]); cudaStreamDestroy(stream[i]); }
            if (buf[i]) cudaFree(buf[i]);
            if (ws[i]) cudaFree(ws[i]);
            stream[i] = nullptr; buf[i] = nullptr; ws[i] = nullptr;
            if (pin_in[i]) cudaFreeHost(pin_in[i]);
            if (pin_out[i]) cudaFreeHost(pin_out[i]);
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
            pin_in[i] = nullptr; pin_out[i] = nullptr; ev_in[i] = nullptr; ev_out[i] = nullptr; in_busy[i] = false;
        }
        if (counts) cudaFree(counts);
        if (mi) cudaFree(mi);
        if (counts_multi) cudaFree(counts_multi);
        if (mi_multi) cudaFree(mi_multi);
        counts = nullptr; mi = nullptr; counts_multi = nullptr; mi_multi = nullptr;
        device = -1;
        buf_bytes = 0;
        buf_cap = 0;
        in_next = 0;
    }
    // A host thread that used the *_host entry points gives its staging memory back when it exits (dask worker
    // threads come and go).  At process teardown the runtime may already be gone: every call then fails with
    // cudaErrorCudartUnloading and nothing is touched.
    ~HostCtx() {
        if (device < 0) return;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) release();
        else cudaGetLastError();
    }
};
static thread_local HostCtx g_ctx;

static int ensure_ctx(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(XBI_ERR_NO_DEVICE, "no CUDA device available (xbitinfo_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(XBI_ERR_ARG, "device index out of range");
    if (g_ctx.device == device) { XBI_CUDA(cudaSetDevice(device), "cudaSetDevice"); return XBI_OK; }
    g_ctx.release();
    XBI_CUDA(cudaSetDevice(device), "cudaSetDevice");
    size_t mb = 256;
    if (const char* env = std::getenv("XBI_STAGING_MB")) {
        const long v = std::atol(env);
        if (v >= 1 && v <= 65536) mb = (size_t)v;
    }
    g_ctx.buf_cap = mb << 20;
    g_ctx.device = device;  // (release() needs it; a half-built context is torn down again below)
    const int rc = [&]() -> int {
        for (int i = 0; i < 2; ++i) {
            XBI_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream[i], cudaStreamNonBlocking), "cudaStreamCreate");
            XBI_CUDA(cudaMalloc(&g_ctx.ws[i], sizeof(Workspace)), "cudaMalloc(workspace)");
        }
        XBI_CUDA(cudaMalloc(&g_ctx.counts, 64 * 4 * sizeof(unsigned long long)), "cudaMalloc(counts)");
        XBI_CUDA(cudaMalloc(&g_ctx.mi, 64 * sizeof(double)), "cudaMalloc(mi)");
        return XBI_OK;
    }();
    if (rc != XBI_OK) g_ctx.release();  // never leave a context that looks ready but holds null pointers
    return rc;
}

// The two staging buffers are sized by what the call at hand needs (at most XBI_STAGING_MB each, default 256):
// a worker thread that rounds 8 MB dask chunks holds 2 x 8 MB of HBM, not 2 x 256 MB.  They only ever grow.
static int ensure_staging(size_t want_bytes) {
    HostCtx& c = g_ctx;
    size_t want = want_bytes < c.buf_cap ? want_bytes : c.buf_cap;
    want = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);  // whole MB
    if (want < ((size_t)1 << 20)) want = (size_t)1 << 20;
    if (want > c.buf_cap) want = c.buf_cap;
    if (c.buf[0] && c.buf[1] && c.buf_bytes >= want) return XBI_OK;
    for (int i = 0; i < 2; ++i) {
        if (c.stream[i]) XBI_CUDA(cudaStreamSynchronize(c.stream[i]), "cudaStreamSynchronize");
        if (c.buf[i]) { cudaFree(c.buf[i]); c.buf[i] = nullptr; }
    }
    c.buf_bytes = 0;
    for (int i = 0; i < 2; ++i) {
        const cudaError_t e = cudaMalloc(&c.buf[i], want);
        if (e != cudaSuccess) {
            for (int k = 0; k < 2; ++k) { if (c.buf[k]) cudaFree(c.buf[k]); c.buf[k] = nullptr; }
            return fail_cuda(e, "cudaMalloc(staging)");
        }
    }
    c.buf_bytes = want;
    return XBI_OK;
}

// ---------------------------------------------------------------------------------
// pageable caller memory: threaded copies through the pinned ring (xbi_hostpool.cu)
// ---------------------------------------------------------------------------------
static bool staging_enabled() {
    static const bool on = [] {
        const char* env = std::getenv("XBI_HOST_STAGING");
        return !(env && env[0] == '0');
    }();
    return on;
}
// true when `p` is ordinary (not page-locked, not CUDA-managed) host memory
static bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}
static int ensure_pinned(HostCtx& c, bool in, bool out) {
    for (int i = 0; i < 2; ++i) {
        if (in && !c.pin_in[i]) {
            XBI_CUDA(cudaHostAlloc(&c.pin_in[i], HostCtx::kPinBytes, cudaHostAllocDefault), "cudaHostAlloc(staging in)");
            XBI_CUDA(cudaEventCreateWithFlags(&c.ev_in[i], cudaEventDisableTiming), "cudaEventCreate");
        }
        if (out && !c.pin_out[i]) {
            XBI_CUDA(cudaHostAlloc(&c.pin_out[i], HostCtx::kPinBytes, cudaHostAllocDefault), "cudaHostAlloc(staging out)");
            XBI_CUDA(cudaEventCreateWithFlags(&c.ev_out[i], cudaEventDisableTiming), "cudaEventCreate");
        }
    }
    return XBI_OK;
}
// rows x row_bytes from pageable `from` (pitch src_pitch) -> contiguous device memory, in pieces of one
// pinned slot: the pool fills slot k+1 while the DMA engine drains slot k.  *handled = false when a single
// row does not fit a slot (the caller then copies directly).
static int staged_h2d(HostCtx& c, void* dst_dev, const char* from, size_t src_pitch, size_t rows, size_t row_bytes,
                      cudaStream_t st, bool* handled) {
    *handled = false;
    const bool contiguous = rows == 1 || src_pitch == row_bytes;
    if (!contiguous && row_bytes > HostCtx::kPinBytes) return XBI_OK;
    if (int rc = ensure_pinned(c, true, false)) return rc;
    auto piece = [&](char* to_dev, const char* src, size_t pitch, size_t nrows, size_t rbytes) -> int {
        const int slot = c.in_next;
        c.in_next ^= 1;
        if (c.in_busy[slot]) XBI_CUDA(cudaEventSynchronize(c.ev_in[slot]), "cudaEventSynchronize(staging)");
        host_copy_parallel(c.pin_in[slot], rbytes, src, pitch, rbytes, nrows);
        XBI_CUDA(cudaMemcpyAsync(to_dev, c.pin_in[slot], nrows * rbytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D, staged)");
        XBI_CUDA(cudaEventRecord(c.ev_in[slot], st), "cudaEventRecord(staging)");
        c.in_busy[slot] = true;
        return XBI_OK;
    };
    char* to = static_cast<char*>(dst_dev);
    if (contiguous) {
        const size_t total = rows * row_bytes;
        for (size_t off = 0; off < total; off += HostCtx::kPinBytes) {
            const size_t len = total - off < HostCtx::kPinBytes ? total - off : HostCtx::kPinBytes;
            if (int rc = piece(to + off, from + off, len, 1, len)) return rc;
        }
    } else {
        const size_t rows_per = HostCtx::kPinBytes / row_bytes;
        for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
            const size_t k = rows - r0 < rows_per ? rows - r0 : rows_per;
            if (int rc = piece(to + r0 * row_bytes, from + r0 * src_pitch, src_pitch, k, row_bytes)) return rc;
        }
    }
    *handled = true;
    return XBI_OK;
}

}  // namespace xbi

using namespace xbi;

extern "C" {

int xbi_abi_version(void) { return XBI_ABI_VERSION; }
const char* xbi_last_error(void) { return g_last_error.c_str(); }
unsigned long long xbi_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
unsigned long long xbi_fast_path_launches(void) { return g_fast_launches.load(std::memory_order_relaxed); }
size_t xbi_workspace_bytes(void) { return sizeof(Workspace); }

void xbi_profile_enable(int on) {
    g_prof.enabled = on != 0;
    g_prof.used = 0;
    g_prof.bytes = 0;
}

int xbi_profile_collect(double* total_ms, unsigned long long* launches, unsigned long long* bytes) {
    double sum = 0.0;
    for (int i = 0; i < g_prof.used; ++i) {
        XBI_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]), "cudaEventSynchronize");
        float ms = 0.f;
        XBI_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]), "cudaEventElapsedTime");
        sum += ms;
    }
    if (total_ms) *total_ms = sum;
    if (launches) *launches = (unsigned long long)g_prof.used;
    if (bytes) *bytes = g_prof.bytes;
    return XBI_OK;
}

int xbi_count_pairs(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner, unsigned flags,
                    uint64_t mask_bits, unsigned long long* counts_dev, void* workspace_dev, size_t workspace_bytes,
                    void* stream) {
    if (!counts_dev || !workspace_dev) return fail(XBI_ERR_ARG, "null counts or workspace pointer");
    if (workspace_bytes < sizeof(Workspace)) return fail(XBI_ERR_ARG, "workspace smaller than xbi_workspace_bytes()");
    if (reinterpret_cast<uintptr_t>(workspace_dev) % 8 != 0) return fail(XBI_ERR_ARG, "workspace must be 8-byte aligned");
    CountJob job;
    if (int rc = make_job(job, x_dev, dtype, outer, n, inner, flags, mask_bits)) return rc;
    if (!x_dev && outer * n * inner > 0) return fail(XBI_ERR_ARG, "null data pointer");
    Finish fin;
    fin.counts = counts_dev;
    return count_pairs_impl(job, flags, static_cast<Workspace*>(workspace_dev), static_cast<cudaStream_t>(stream), fin);
}

static int bitinformation_dev_impl(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner, unsigned flags,
                                   uint64_t mask_bits, int set_zero_insignificant, double confidence,
                                   unsigned long long* counts_dev, double* mi_dev, void* workspace_dev,
                                   size_t workspace_bytes, void* stream, const PeerExchange& px) {
    if (!counts_dev || !workspace_dev) return fail(XBI_ERR_ARG, "null counts or workspace pointer");
    if (workspace_bytes < sizeof(Workspace)) return fail(XBI_ERR_ARG, "workspace smaller than xbi_workspace_bytes()");
    if (reinterpret_cast<uintptr_t>(workspace_dev) % 8 != 0) return fail(XBI_ERR_ARG, "workspace must be 8-byte aligned");
    CountJob job;
    if (int rc = make_job(job, x_dev, dtype, outer, n, inner, flags, mask_bits)) return rc;
    if (!x_dev && outer * n * inner > 0) return fail(XBI_ERR_ARG, "null data pointer");
    Finish fin;
    fin.counts = counts_dev;
    fin.accumulate = false;
    fin.mi = mi_dev;
    fin.px = px;
    if (mi_dev && set_zero_insignificant) {
        if (!(confidence > 0.0 && confidence < 1.0)) return fail(XBI_ERR_ARG, "confidence must be in (0,1)");
        fin.set_zero = 1;
        fin.z = normal_quantile(1.0 - (1.0 - confidence) / 2.0);
    }
    return count_pairs_impl(job, flags, static_cast<Workspace*>(workspace_dev), static_cast<cudaStream_t>(stream), fin);
}

int xbi_bitinformation_dev(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner, unsigned flags,
                           uint64_t mask_bits, int set_zero_insignificant, double confidence,
                           unsigned long long* counts_dev, double* mi_dev, void* workspace_dev, size_t workspace_bytes,
                           void* stream) {
    PeerExchange none = {};
    return bitinformation_dev_impl(x_dev, dtype, outer, n, inner, flags, mask_bits, set_zero_insignificant, confidence,
                                   counts_dev, mi_dev, workspace_dev, workspace_bytes, stream, none);
}

size_t xbi_workspace_bytes_multi(int naxes) {
    return (size_t)(naxes > 0 ? naxes : 1) * sizeof(Workspace) + 64;  // one workspace per axis + the "masked element seen" flag
}

int xbi_bitinformation_dev_multi(const void* x_dev, int dtype, int ndim, const int64_t* shape, int naxes, const int* axes,
                                 unsigned flags, uint64_t mask_bits, int set_zero_insignificant, double confidence,
                                 unsigned long long* counts_dev, double* mi_dev, void* workspace_dev, size_t workspace_bytes,
                                 void* stream) {
    if (ndim < 1 || ndim > XBI_MAX_DIMS) return fail(XBI_ERR_ARG, "ndim must be in 1..XBI_MAX_DIMS");
    if (naxes < 1 || naxes > XBI_MAX_DIMS) return fail(XBI_ERR_ARG, "naxes must be in 1..XBI_MAX_DIMS");
    if (!shape || !axes) return fail(XBI_ERR_ARG, "null shape / axes pointer");
    if (!counts_dev || !workspace_dev) return fail(XBI_ERR_ARG, "null counts or workspace pointer");
    if (workspace_bytes < xbi_workspace_bytes_multi(naxes)) return fail(XBI_ERR_ARG, "workspace smaller than xbi_workspace_bytes_multi(naxes)");
    if (reinterpret_cast<uintptr_t>(workspace_dev) % 8 != 0) return fail(XBI_ERR_ARG, "workspace must be 8-byte aligned");
    int64_t numel = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return fail(XBI_ERR_ARG, "negative extent");
        numel *= shape[d];
    }
    int pos_y = -1, pos_x = -1;  // where the two innermost axes sit in `axes`
    for (int a = 0; a < naxes; ++a) {
        if (axes[a] < 0 || axes[a] >= ndim) return fail(XBI_ERR_ARG, "axis out of range");
        for (int b = 0; b < a; ++b)
            if (axes[b] == axes[a]) return fail(XBI_ERR_ARG, "an axis is listed twice");
        if (axes[a] == ndim - 2) pos_y = a;
        if (axes[a] == ndim - 1) pos_x = a;
    }
    CountJob probe;
    if (int rc = make_job(probe, x_dev, dtype, 0, 0, 0, flags, mask_bits)) return rc;
    if (!x_dev && numel > 0) return fail(XBI_ERR_ARG, "null data pointer");
    const int nbits = 8 * dtype_bytes(dtype);
    const size_t table = (size_t)nbits * 4;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace* ws = static_cast<Workspace*>(workspace_dev);
    unsigned int* dirty_flag = reinterpret_cast<unsigned int*>(ws + naxes);
    double z = 0.0;
    if (mi_dev && set_zero_insignificant) {
        if (!(confidence > 0.0 && confidence < 1.0)) return fail(XBI_ERR_ARG, "confidence must be in (0,1)");
        z = normal_quantile(1.0 - (1.0 - confidence) / 2.0);
    }
    auto job_of = [&](int axis) {
        CountJob j = probe;
        j.outer = 1; j.inner = 1;
        for (int d = 0; d < axis; ++d) j.outer *= shape[d];
        j.n = shape[axis];
        for (int d = axis + 1; d < ndim; ++d) j.inner *= shape[d];
        return j;
    };
    auto fin_of = [&](int a) {
        Finish fin;
        fin.counts = counts_dev + (size_t)a * table;
        fin.accumulate = false;
        fin.mi = mi_dev ? mi_dev + (size_t)a * nbits : nullptr;
        fin.set_zero = (mi_dev && set_zero_insignificant) ? 1 : 0;
        fin.z = z;
        return fin;
    };
    bool fused = false;
    if (pos_y >= 0 && pos_x >= 0 && numel > 0 && !(flags & XBI_FORCE_GENERIC)) {
        const CountJob jy = job_of(ndim - 2);
        if (pairs_xy_eligible(jy)) {
            // the two innermost axes in ONE read (xbi_count_xy.cu)
            XBI_CUDA(cudaMemsetAsync(ws, 0, xbi_workspace_bytes_multi(naxes), st), "cudaMemsetAsync(workspace)");
            XBI_CUDA(launch_pairs_xy(jy, ws + pos_y, ws + pos_x, dirty_flag, st), "pair counting (two axes)");
            fused = true;
            if (jy.mask_mode != kMaskNone) {
                // the kernel counted as if nothing were masked and only looked for masked elements: if there
                // are any, both axes are recounted by the masked single-axis kernels
                unsigned int seen = 0;
                XBI_CUDA(cudaMemcpyAsync(&seen, dirty_flag, sizeof(seen), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(flag)");
                XBI_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
                if (seen) fused = false;
            }
            if (fused) {
                const CountJob jx = job_of(ndim - 1);
                XBI_CUDA(launch_finish(jy, finish_job(jy, ws + pos_y, fin_of(pos_y)), st), "finish (combine)");
                XBI_CUDA(launch_finish(jx, finish_job(jx, ws + pos_x, fin_of(pos_x)), st), "finish (combine)");
            }
        }
    }
    for (int a = 0; a < naxes; ++a) {
        if (fused && (a == pos_y || a == pos_x)) continue;
        if (int rc = count_pairs_impl(job_of(axes[a]), flags, ws + a, st, fin_of(a))) return rc;
    }
    return XBI_OK;
}

// ---------------------------------------------------------------------------------
// all-reduce over peer memory: exchange buffers (cudaMalloc + CUDA IPC) and the fused entry point
// ---------------------------------------------------------------------------------
size_t xbi_peer_bytes(void) { return kPeerWords * sizeof(unsigned long long); }
size_t xbi_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int xbi_peer_create(void** buffer_dev, void* handle_out) {
    if (!buffer_dev || !handle_out) return fail(XBI_ERR_ARG, "null pointer");
    void* p = nullptr;
    XBI_CUDA(cudaMalloc(&p, xbi_peer_bytes()), "cudaMalloc(peer exchange buffer)");
    cudaError_t e = cudaMemset(p, 0, xbi_peer_bytes());
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail_cuda(e, "cudaIpcGetMemHandle");
    }
    std::memcpy(handle_out, &h, sizeof(h));
    *buffer_dev = p;
    return XBI_OK;
}

int xbi_peer_open(const void* handle, void** mapped_dev) {
    if (!handle || !mapped_dev) return fail(XBI_ERR_ARG, "null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    XBI_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    *mapped_dev = p;
    return XBI_OK;
}

int xbi_peer_close(void* mapped_dev) {
    if (mapped_dev) XBI_CUDA(cudaIpcCloseMemHandle(mapped_dev), "cudaIpcCloseMemHandle");
    return XBI_OK;
}

int xbi_peer_destroy(void* buffer_dev) {
    if (buffer_dev) XBI_CUDA(cudaFree(buffer_dev), "cudaFree(peer exchange buffer)");
    return XBI_OK;
}

int xbi_peer_error_epoch(const void* own_buffer_dev, unsigned long long* epoch_out) {
    if (!own_buffer_dev || !epoch_out) return fail(XBI_ERR_ARG, "null pointer");
    XBI_CUDA(cudaMemcpy(epoch_out, static_cast<const unsigned long long*>(own_buffer_dev) + kPeerErrorWord,
                        sizeof(unsigned long long), cudaMemcpyDeviceToHost),
             "cudaMemcpy(peer error word)");
    return XBI_OK;
}

int xbi_bitinformation_dev_allreduce(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner, unsigned flags,
                                     uint64_t mask_bits, int set_zero_insignificant, double confidence,
                                     unsigned long long* counts_dev, double* mi_dev, void* workspace_dev,
                                     size_t workspace_bytes, void* stream, int world, int rank, uint64_t epoch,
                                     void* const* peer_buffers, double timeout_s) {
    if (world < 1 || world > kPeerMaxWorld) return fail(XBI_ERR_ARG, "world must be in 1..8 (the GPUs of one node)");
    if (rank < 0 || rank >= world) return fail(XBI_ERR_ARG, "rank out of range");
    if (epoch == 0) return fail(XBI_ERR_ARG, "epochs start at 1");
    if (!peer_buffers) return fail(XBI_ERR_ARG, "null peer buffer table");
    PeerExchange px = {};
    px.world = world;
    px.rank = rank;
    px.epoch = epoch;
    px.timeout_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 20.0) * 1e9);
    for (int p = 0; p < world; ++p) {
        if (!peer_buffers[p]) return fail(XBI_ERR_ARG, "null peer buffer");
        px.peer[p] = static_cast<unsigned long long*>(peer_buffers[p]);
    }
    return bitinformation_dev_impl(x_dev, dtype, outer, n, inner, flags, mask_bits, set_zero_insignificant, confidence,
                                   counts_dev, mi_dev, workspace_dev, workspace_bytes, stream, px);
}

int xbi_mutual_information_batched(const unsigned long long* counts_dev, int nbits, int64_t nbatch,
                                   int set_zero_insignificant, double confidence, double* mi_dev, void* stream) {
    if (nbatch < 0) return fail(XBI_ERR_ARG, "negative nbatch");
    if (nbits < 1 || nbits > 64) return fail(XBI_ERR_ARG, "nbits must be in 1..64");
    if (nbatch == 0) return XBI_OK;
    if (!counts_dev || !mi_dev) return fail(XBI_ERR_ARG, "null pointer");
    double z = 0.0;
    if (set_zero_insignificant) {
        if (!(confidence > 0.0 && confidence < 1.0)) return fail(XBI_ERR_ARG, "confidence must be in (0,1)");
        z = normal_quantile(1.0 - (1.0 - confidence) / 2.0);
    }
    XBI_CUDA(launch_mutual_information(counts_dev, nbits, nbatch, set_zero_insignificant ? 1 : 0, z, mi_dev,
                                       static_cast<cudaStream_t>(stream)),
             "mutual information");
    return XBI_OK;
}

int xbi_mutual_information(const unsigned long long* counts_dev, int nbits, int set_zero_insignificant,
                           double confidence, double* mi_dev, void* stream) {
    if (!counts_dev || !mi_dev) return fail(XBI_ERR_ARG, "null pointer");
    return xbi_mutual_information_batched(counts_dev, nbits, 1, set_zero_insignificant, confidence, mi_dev, stream);
}

// Chunk grid of an N-d C-contiguous array -> ChunkGrid of at most kMaxChunkDims dimensions.
// The grid is regular (chunk_shape) or given by per-dimension offset tables (nchunks_dim + bounds: for every
// dimension in turn nchunks_dim[d]+1 ascending offsets from 0 to shape[d]; the same table on the host for
// planning and on the device for the kernels; a dimension whose table is regular is treated as such).
// Neighbouring dimensions are merged when the inner one is not chunked and neither is the analysed
// one (the chunk order stays the C order of the chunk grid); dimensions of length 1 disappear.
// *nchunks: chunks in the grid; *empty: nothing to do (no elements, or no pairs along the axis).
static int plan_chunk_grid(int ndim, const int64_t* shape, const int64_t* chunk_shape, const int64_t* nchunks_dim,
                           const int64_t* bounds_host, const int64_t* bounds_dev, int axis, ChunkGrid& g,
                           int64_t* nchunks, bool* empty) {
    if (ndim < 1 || ndim > XBI_MAX_DIMS) return fail(XBI_ERR_ARG, "ndim must be in 1..XBI_MAX_DIMS");
    if (!shape || (!chunk_shape && !(nchunks_dim && bounds_host && bounds_dev)))
        return fail(XBI_ERR_ARG, "null shape / chunk description pointer");
    if (axis < -1 || axis >= ndim) return fail(XBI_ERR_ARG, "axis out of range");
    int64_t total_chunks = 1, numel = 1;
    int64_t cmax[XBI_MAX_DIMS];                 // (largest) chunk extent per dimension
    int64_t ncd[XBI_MAX_DIMS];                  // chunks per dimension
    const int64_t* table[XBI_MAX_DIMS];         // device offsets of the irregular dimensions, else nullptr
    size_t toff = 0;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return fail(XBI_ERR_ARG, "negative extent");
        int64_t nc;
        table[d] = nullptr;
        if (chunk_shape) {
            if (chunk_shape[d] < 1) return fail(XBI_ERR_ARG, "chunk extents must be positive");
            cmax[d] = chunk_shape[d] < shape[d] ? chunk_shape[d] : (shape[d] > 0 ? shape[d] : 1);
            nc = shape[d] > 0 ? (shape[d] + cmax[d] - 1) / cmax[d] : 0;
        } else {
            nc = nchunks_dim[d];
            if (nc < 0 || (nc == 0) != (shape[d] == 0)) return fail(XBI_ERR_ARG, "chunk count does not fit the extent");
            const int64_t* b = bounds_host + toff;
            if (b[0] != 0 || b[nc] != shape[d]) return fail(XBI_ERR_ARG, "chunk offsets must run from 0 to the extent");
            bool regular = true;
            cmax[d] = nc > 0 ? b[1] - b[0] : 1;
            for (int64_t q = 0; q < nc; ++q) {
                const int64_t ext = b[q + 1] - b[q];
                if (ext < 1) return fail(XBI_ERR_ARG, "chunk offsets must be strictly ascending");
                if (q + 1 < nc ? ext != b[1] - b[0] : ext > b[1] - b[0]) regular = false;
                if (ext > cmax[d]) cmax[d] = ext;
            }
            if (!regular) table[d] = bounds_dev + toff;
            toff += (size_t)nc + 1;
        }
        if (nc > 0 && total_chunks > (int64_t)0x7FFFFFFF / nc) return fail(XBI_ERR_UNSUPPORTED, "more than 2^31-1 chunks");
        ncd[d] = nc;
        total_chunks *= nc;
        numel *= shape[d];
    }
    *nchunks = total_chunks;
    *empty = numel == 0 || (axis >= 0 && shape[axis] < 2);
    if (numel == 0) return XBI_OK;

    // innermost first: merge dimension d into the group built so far when that group is unchunked
    int64_t mshape[XBI_MAX_DIMS], mchunk[XBI_MAX_DIMS], mstride[XBI_MAX_DIMS], mscale[XBI_MAX_DIMS], mnchunk[XBI_MAX_DIMS];
    const int64_t* mtable[XBI_MAX_DIMS];
    int maxis = -1, m = 0;  // groups are collected innermost first
    int64_t stride = 1;
    bool can_merge = false;  // the current (innermost so far) group may take another outer dimension
    for (int d = ndim - 1; d >= 0; --d) {
        const int64_t c = cmax[d];
        const bool is_axis = d == axis;
        if (shape[d] == 1 && !is_axis) { stride *= shape[d]; continue; }
        if (m > 0 && can_merge && !is_axis && mchunk[m - 1] == mshape[m - 1] && mtable[m - 1] == nullptr &&
            c <= (int64_t)0x7FFFFFFF / mshape[m - 1]) {
            mchunk[m - 1] = c * mshape[m - 1];
            mscale[m - 1] = mshape[m - 1];  // offsets of dimension d count whole inner groups
            mtable[m - 1] = table[d];
            mnchunk[m - 1] = ncd[d];
            mshape[m - 1] *= shape[d];
        } else {
            mshape[m] = shape[d];
            mchunk[m] = c;
            mstride[m] = stride;
            mscale[m] = 1;
            mtable[m] = table[d];
            mnchunk[m] = ncd[d];
            if (is_axis) maxis = m;
            can_merge = !is_axis;
            ++m;
        }
        stride *= shape[d];
    }
    if (m == 0) {  // a single element
        mshape[0] = 1; mchunk[0] = 1; mstride[0] = 1; mscale[0] = 1; mtable[0] = nullptr; mnchunk[0] = 1;
        m = 1;
    }
    if (m > kMaxChunkDims)
        return fail(XBI_ERR_UNSUPPORTED, "more than 5 chunked dimensions after merging; process the chunks one by one");
    double full = 1.0;
    for (int i = 0; i < kMaxChunkDims; ++i) {
        g.shape[i] = 1; g.stride[i] = 0; g.chunk[i] = 1; g.nchunk[i] = 1; g.bounds[i] = nullptr; g.scale[i] = 1;
    }
    g.axis = -1;
    g.vec16 = 0;  // (set by the walkers' planner, xbi_chunked_walk.cu)
    g.lo = kMaxChunkDims - m;
    for (int i = 0; i < m; ++i) {  // group i (innermost first) -> slot kMaxChunkDims-1-i
        const int slot = kMaxChunkDims - 1 - i;
        if (mchunk[i] > (int64_t)0x7FFFFFFF)
            return fail(XBI_ERR_UNSUPPORTED, "a chunk extent of 2^31 or more; use xbi_count_pairs on that chunk");
        g.shape[slot] = mshape[i];
        g.stride[slot] = mstride[i];
        g.chunk[slot] = (unsigned)mchunk[i];
        g.nchunk[slot] = (unsigned)mnchunk[i];
        g.bounds[slot] = reinterpret_cast<const long long*>(mtable[i]);
        g.scale[slot] = mscale[i];
        if (i == maxis) g.axis = slot;
        full *= (double)mchunk[i];
    }
    if (full >= 9.0e18) return fail(XBI_ERR_UNSUPPORTED, "chunk too large");
    unsigned long long full_numel = 1;
    for (int i = 0; i < kMaxChunkDims; ++i) full_numel *= g.chunk[i];
    const unsigned long long tile = 4096ull;
    const unsigned long long tiles_per_chunk = (full_numel + tile - 1) / tile;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long target_items = (unsigned long long)sms * 16ull;
    unsigned long long groups = (target_items + (unsigned long long)total_chunks - 1) / (unsigned long long)total_chunks;
    if (groups < 1) groups = 1;
    if (groups > tiles_per_chunk) groups = tiles_per_chunk;
    unsigned long long tpg = (tiles_per_chunk + groups - 1) / groups;
    groups = (tiles_per_chunk + tpg - 1) / tpg;
    if (tpg > 0xFFFFFFFFull || groups * (unsigned long long)total_chunks > 0xFFFFFFFFull)
        return fail(XBI_ERR_UNSUPPORTED, "chunk grid too large for one launch");
    g.groups = (unsigned)groups;
    g.tiles_per_group = (unsigned)tpg;
    g.items = (unsigned)(groups * (unsigned long long)total_chunks);
    return XBI_OK;
}

static int count_pairs_grid(const void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* chunk_shape,
                            const int64_t* nchunks_dim, const int64_t* bounds_host, const int64_t* bounds_dev, int axis,
                            unsigned flags, uint64_t mask_bits, unsigned long long* counts_dev, void* stream) {
    CountJob job;
    if (int rc = make_job(job, x_dev, dtype, 0, 0, 0, flags, mask_bits)) return rc;
    if (axis < 0) return fail(XBI_ERR_ARG, "axis out of range");
    ChunkGrid g;
    int64_t nchunks = 0;
    bool empty = false;
    if (int rc = plan_chunk_grid(ndim, shape, chunk_shape, nchunks_dim, bounds_host, bounds_dev, axis, g, &nchunks, &empty))
        return rc;
    if (nchunks == 0) return XBI_OK;
    if (!counts_dev) return fail(XBI_ERR_ARG, "null counts pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t table = (size_t)8 * dtype_bytes(dtype) * 4 * sizeof(unsigned long long);
    XBI_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)nchunks * table, st), "cudaMemsetAsync(counts)");
    if (empty) return XBI_OK;
    if (!x_dev) return fail(XBI_ERR_ARG, "null data pointer");
    XBI_CUDA(launch_pairs_chunked(x_dev, dtype, g, (unsigned)nchunks, job.transform, job.mask_mode, job.mask_bits,
                                  counts_dev, st, flags),
             "pair counting (chunk grid)");
    return XBI_OK;
}

int xbi_count_pairs_chunked(const void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* chunk_shape,
                            int axis, unsigned flags, uint64_t mask_bits, unsigned long long* counts_dev, void* stream) {
    if (!chunk_shape) return fail(XBI_ERR_ARG, "null chunk_shape pointer");
    return count_pairs_grid(x_dev, dtype, ndim, shape, chunk_shape, nullptr, nullptr, nullptr, axis, flags, mask_bits,
                            counts_dev, stream);
}

int xbi_count_pairs_chunked_v(const void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* nchunks_dim,
                              const int64_t* bounds_host, const int64_t* bounds_dev, int axis, unsigned flags,
                              uint64_t mask_bits, unsigned long long* counts_dev, void* stream) {
    if (!nchunks_dim || !bounds_host || !bounds_dev) return fail(XBI_ERR_ARG, "null chunk table pointer");
    return count_pairs_grid(x_dev, dtype, ndim, shape, nullptr, nchunks_dim, bounds_host, bounds_dev, axis, flags,
                            mask_bits, counts_dev, stream);
}

static int bitround_grid(void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* chunk_shape,
                         const int64_t* nchunks_dim, const int64_t* bounds_host, const int64_t* bounds_dev,
                         const int32_t* keepbits_dev, void* stream) {
    if (!dtype_is_float(dtype)) return fail(XBI_ERR_ARG, "Only float arrays (16-64bit) can be bit-rounded");
    ChunkGrid g;
    int64_t nchunks = 0;
    bool empty = false;
    if (int rc = plan_chunk_grid(ndim, shape, chunk_shape, nchunks_dim, bounds_host, bounds_dev, -1, g, &nchunks, &empty))
        return rc;
    if (nchunks == 0 || empty) return XBI_OK;
    if (!x_dev || !keepbits_dev) return fail(XBI_ERR_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(x_dev) % dtype_bytes(dtype) != 0)
        return fail(XBI_ERR_ARG, "data pointer is not element-aligned");
    XBI_CUDA(launch_bitround_chunked(x_dev, dtype, g, keepbits_dev, static_cast<cudaStream_t>(stream)),
             "bitround (chunk grid)");
    return XBI_OK;
}

int xbi_bitround_chunked(void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* chunk_shape,
                         const int32_t* keepbits_dev, void* stream) {
    if (!chunk_shape) return fail(XBI_ERR_ARG, "null chunk_shape pointer");
    return bitround_grid(x_dev, dtype, ndim, shape, chunk_shape, nullptr, nullptr, nullptr, keepbits_dev, stream);
}

int xbi_bitround_chunked_v(void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* nchunks_dim,
                           const int64_t* bounds_host, const int64_t* bounds_dev, const int32_t* keepbits_dev,
                           void* stream) {
    if (!nchunks_dim || !bounds_host || !bounds_dev) return fail(XBI_ERR_ARG, "null chunk table pointer");
    return bitround_grid(x_dev, dtype, ndim, shape, nullptr, nchunks_dim, bounds_host, bounds_dev, keepbits_dev, stream);
}

static int check_bitround_args(int dtype, int64_t numel, int keepbits) {
    if (!dtype_is_float(dtype)) return fail(XBI_ERR_ARG, "Only float arrays (16-64bit) can be bit-rounded");
    const int mant = dtype == XBI_F16 ? 10 : dtype == XBI_F32 ? 23 : 52;
    if (keepbits < 0) return fail(XBI_ERR_ARG, "keepbits must be zero or positive");
    if (keepbits > mant) return fail(XBI_ERR_ARG, "keepbits too large for given dtype");
    if (numel < 0) return fail(XBI_ERR_ARG, "negative numel");
    return XBI_OK;
}

int xbi_bitround_inplace(void* x_dev, int dtype, int64_t numel, int keepbits, void* stream) {
    if (int rc = check_bitround_args(dtype, numel, keepbits)) return rc;
    if (numel == 0) return XBI_OK;
    if (!x_dev) return fail(XBI_ERR_ARG, "null data pointer");
    if (reinterpret_cast<uintptr_t>(x_dev) % dtype_bytes(dtype) != 0)
        return fail(XBI_ERR_ARG, "data pointer is not element-aligned");
    XBI_CUDA(launch_bitround(x_dev, dtype, numel, keepbits, static_cast<cudaStream_t>(stream)), "bitround");
    return XBI_OK;
}

int xbi_bitround(const void* src_dev, void* dst_dev, int dtype, int64_t numel, int keepbits, void* stream) {
    if (int rc = check_bitround_args(dtype, numel, keepbits)) return rc;
    if (numel == 0) return XBI_OK;
    if (!src_dev || !dst_dev) return fail(XBI_ERR_ARG, "null data pointer");
    const int bytes = dtype_bytes(dtype);
    const uintptr_t a = reinterpret_cast<uintptr_t>(src_dev), b = reinterpret_cast<uintptr_t>(dst_dev);
    if (a % bytes != 0 || b % bytes != 0) return fail(XBI_ERR_ARG, "data pointer is not element-aligned");
    const uint64_t nb = (uint64_t)numel * (uint64_t)bytes;
    if (a != b && a < b + nb && b < a + nb) return fail(XBI_ERR_ARG, "source and destination overlap");
    XBI_CUDA(launch_bitround_to(src_dev, dst_dev, dtype, numel, keepbits, static_cast<cudaStream_t>(stream)), "bitround");
    return XBI_OK;
}

int xbi_shuffle(const void* src_dev, void* dst_dev, int itemsize, int64_t numel, int kind, int inverse, void* stream) {
    if (itemsize != 2 && itemsize != 4 && itemsize != 8) return fail(XBI_ERR_ARG, "itemsize must be 2, 4 or 8");
    if (kind != XBI_SHUFFLE_BYTES && kind != XBI_SHUFFLE_BITS) return fail(XBI_ERR_ARG, "unknown shuffle kind");
    if (numel < 0) return fail(XBI_ERR_ARG, "negative numel");
    if (numel == 0) return XBI_OK;
    if (!src_dev || !dst_dev) return fail(XBI_ERR_ARG, "null data pointer");
    if (kind == XBI_SHUFFLE_BITS && numel % 8 != 0)
        return fail(XBI_ERR_ARG, "bit shuffle needs a number of elements that is a multiple of 8");
    const uintptr_t a = reinterpret_cast<uintptr_t>(src_dev), b = reinterpret_cast<uintptr_t>(dst_dev);
    const uint64_t bytes = (uint64_t)numel * (uint64_t)itemsize;
    if (a < b + bytes && b < a + bytes) return fail(XBI_ERR_ARG, "source and destination overlap");
    XBI_CUDA(launch_shuffle(src_dev, dst_dev, itemsize, numel, kind == XBI_SHUFFLE_BITS, inverse == 0, 0,
                            static_cast<cudaStream_t>(stream)),
             "shuffle");
    return XBI_OK;
}

int xbi_bitround_shuffle(const void* src_dev, void* dst_dev, int dtype, int64_t numel, int keepbits, int kind,
                         void* stream) {
    if (int rc = check_bitround_args(dtype, numel, keepbits)) return rc;
    if (kind != XBI_SHUFFLE_BYTES && kind != XBI_SHUFFLE_BITS) return fail(XBI_ERR_ARG, "unknown shuffle kind");
    if (numel == 0) return XBI_OK;
    if (!src_dev || !dst_dev) return fail(XBI_ERR_ARG, "null data pointer");
    if (kind == XBI_SHUFFLE_BITS && numel % 8 != 0)
        return fail(XBI_ERR_ARG, "bit shuffle needs a number of elements that is a multiple of 8");
    const int itemsize = dtype_bytes(dtype);
    const uintptr_t a = reinterpret_cast<uintptr_t>(src_dev), b = reinterpret_cast<uintptr_t>(dst_dev);
    const uint64_t bytes = (uint64_t)numel * (uint64_t)itemsize;
    if (a < b + bytes && b < a + bytes) return fail(XBI_ERR_ARG, "source and destination overlap");
    const int mant = dtype == XBI_F16 ? 10 : dtype == XBI_F32 ? 23 : 52;
    XBI_CUDA(launch_shuffle(src_dev, dst_dev, itemsize, numel, kind == XBI_SHUFFLE_BITS, true, mant - keepbits,
                            static_cast<cudaStream_t>(stream)),
             "bitround + shuffle");
    return XBI_OK;
}

// Slab plan for streaming a host (outer, n, inner) array through a staging buffer.
//   mode 0: whole blocks   -- `blocks` consecutive outer indices per slab (1-D copy)
//   mode 1: column strips  -- all n rows x `cols` columns of one outer index (2-D copy)
//   mode 2: row strips     -- `rows` (+1 halo) rows x `cols` columns of one outer index
int xbi_bitinformation_host(const void* x_host, int dtype, int64_t outer, int64_t n, int64_t inner, unsigned flags,
                            uint64_t mask_bits, int set_zero_insignificant, double confidence, double* mi_host,
                            unsigned long long* counts_host, int device) {
    if (!mi_host) return fail(XBI_ERR_ARG, "null mi_host");
    CountJob probe;
    if (int rc = make_job(probe, x_host, dtype, outer, n, inner, flags, mask_bits)) return rc;
    if (int rc = ensure_ctx(device)) return rc;
    HostCtx& c = g_ctx;
    const int bytes = dtype_bytes(dtype);
    const int nbits = 8 * bytes;
    const int64_t numel = outer * n * inner;
    if (numel > 0 && !x_host) return fail(XBI_ERR_ARG, "null data pointer");
    if (int rc = ensure_staging((size_t)numel * bytes)) return rc;
    XBI_CUDA(cudaMemsetAsync(c.counts, 0, 64 * 4 * sizeof(unsigned long long), c.stream[0]), "memset(counts)");
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");

    const char* src = static_cast<const char*>(x_host);
    const int64_t cap = (int64_t)(c.buf_bytes / bytes);  // elements per staging buffer
    const bool stage = numel > 0 && staging_enabled() && is_pageable(x_host);
    int which = 0;
    auto submit = [&](const void* from, size_t src_pitch, int64_t rows, int64_t cols, int64_t so, int64_t sn,
                      int64_t si) -> int {
        cudaStream_t st = c.stream[which];
        // the stream's previous slab (copy + kernels) must be finished before its buffer is reused:
        // same-stream ordering guarantees that.
        bool staged = false;
        if (stage)
            if (int rc = staged_h2d(c, c.buf[which], static_cast<const char*>(from), src_pitch, (size_t)rows,
                                    (size_t)cols * bytes, st, &staged))
                return rc;
        if (staged) {
        } else if (src_pitch == (size_t)cols * bytes || rows == 1) {
            XBI_CUDA(cudaMemcpyAsync(c.buf[which], from, (size_t)rows * cols * bytes, cudaMemcpyHostToDevice, st),
                     "cudaMemcpyAsync(H2D)");
        } else {
            XBI_CUDA(cudaMemcpy2DAsync(c.buf[which], (size_t)cols * bytes, from, src_pitch, (size_t)cols * bytes,
                                       (size_t)rows, cudaMemcpyHostToDevice, st),
                     "cudaMemcpy2DAsync(H2D)");
        }
        CountJob job = probe;
        job.x = c.buf[which];
        job.outer = so; job.n = sn; job.inner = si;
        Finish fin;
        fin.counts = c.counts;
        if (int rc = count_pairs_impl(job, flags, c.ws[which], st, fin)) return rc;
        which ^= 1;
        return XBI_OK;
    };

    if (numel > 0 && n >= 2) {
        const int64_t block = n * inner;
        if (block <= cap) {
            const int64_t per = cap / block;
            for (int64_t o = 0; o < outer; o += per) {
                const int64_t k = (outer - o < per) ? (outer - o) : per;
                if (int rc = submit(src + (size_t)o * block * bytes, (size_t)block * k * bytes, 1, block * k, k, n, inner))
                    return rc;
            }
        } else {
            // strips of one outer index.  Column strips keep 16-byte multiples so the fast
            // path stays eligible; row strips carry one halo row (the first row of the next strip).
            const int64_t gran = 64;
            int64_t cols, rows_per = 0;
            const bool all_rows = n <= cap;
            if (all_rows) {
                cols = cap / n;
                if (cols >= inner) cols = inner;
                else if (cols >= gran) cols -= cols % gran;
            } else {
                cols = inner < gran ? inner : gran;
                rows_per = cap / cols - 1;
                if (rows_per < 1)
                    return fail(XBI_ERR_UNSUPPORTED, "staging buffer too small for this shape; raise XBI_STAGING_MB");
            }
            const size_t pitch = (size_t)inner * bytes;
            for (int64_t o = 0; o < outer; ++o) {
                for (int64_t c0 = 0; c0 < inner; c0 += cols) {
                    const int64_t w = (inner - c0 < cols) ? (inner - c0) : cols;
                    if (all_rows) {
                        const char* from = src + ((size_t)(o * n) * inner + c0) * bytes;
                        if (int rc = submit(from, pitch, n, w, 1, n, w)) return rc;
                    } else {
                        for (int64_t r0 = 0; r0 < n - 1; r0 += rows_per) {
                            const int64_t take = (n - r0 < rows_per + 1) ? (n - r0) : (rows_per + 1);
                            const char* from = src + ((size_t)(o * n + r0) * inner + c0) * bytes;
                            if (int rc = submit(from, pitch, take, w, 1, take, w)) return rc;
                        }
                    }
                }
            }
        }
    }
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
    XBI_CUDA(cudaStreamSynchronize(c.stream[1]), "sync");
    if (int rc = xbi_mutual_information(c.counts, nbits, set_zero_insignificant, confidence, c.mi, c.stream[0]))
        return rc;
    XBI_CUDA(cudaMemcpyAsync(mi_host, c.mi, nbits * sizeof(double), cudaMemcpyDeviceToHost, c.stream[0]), "D2H(mi)");
    if (counts_host)
        XBI_CUDA(cudaMemcpyAsync(counts_host, c.counts, (size_t)nbits * 4 * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, c.stream[0]),
                 "D2H(counts)");
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
    return XBI_OK;
}

int xbi_bitinformation_host_multi(const void* x_host, int dtype, int ndim, const int64_t* shape, int naxes,
                                  const int* axes, unsigned flags, uint64_t mask_bits, int set_zero_insignificant,
                                  double confidence, double* mi_host, unsigned long long* counts_host, int device) {
    if (!mi_host) return fail(XBI_ERR_ARG, "null mi_host");
    if (ndim < 1 || ndim > XBI_MAX_DIMS) return fail(XBI_ERR_ARG, "ndim must be in 1..XBI_MAX_DIMS");
    if (naxes < 1 || naxes > XBI_MAX_DIMS) return fail(XBI_ERR_ARG, "naxes must be in 1..XBI_MAX_DIMS");
    if (!shape || !axes) return fail(XBI_ERR_ARG, "null shape / axes pointer");
    int64_t numel = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return fail(XBI_ERR_ARG, "negative extent");
        numel *= shape[d];
    }
    for (int a = 0; a < naxes; ++a)
        if (axes[a] < 0 || axes[a] >= ndim) return fail(XBI_ERR_ARG, "axis out of range");
    CountJob probe;
    if (int rc = make_job(probe, x_host, dtype, 0, 0, 0, flags, mask_bits)) return rc;
    const int bytes = dtype_bytes(dtype);
    const int nbits = 8 * bytes;
    auto dims_of = [&](int axis, int64_t rows0, int64_t* outer, int64_t* n, int64_t* inner) {
        // (outer, n, inner) of an array whose first extent is rows0 instead of shape[0]
        *outer = 1; *inner = 1;
        for (int d = 0; d < axis; ++d) *outer *= (d == 0 ? rows0 : shape[d]);
        *n = axis == 0 ? rows0 : shape[axis];
        for (int d = axis + 1; d < ndim; ++d) *inner *= shape[d];
    };
    if (int rc = ensure_ctx(device)) return rc;
    HostCtx& c = g_ctx;
    const int64_t row = shape[0] > 0 ? numel / shape[0] : 0;  // elements of one index of the first dimension
    if (int rc = ensure_staging((size_t)(numel + row) * bytes)) return rc;
    const int64_t cap = (int64_t)(c.buf_bytes / bytes);
    if (naxes == 1 || numel == 0 || row == 0 || 2 * row > cap) {
        // nothing to share between the axes (or one row of the first dimension does not fit the staging
        // buffer next to its halo): one pass over the host array per axis
        for (int a = 0; a < naxes; ++a) {
            int64_t outer, n, inner;
            dims_of(axes[a], shape[0], &outer, &n, &inner);
            if (int rc = xbi_bitinformation_host(x_host, dtype, outer, n, inner, flags, mask_bits, set_zero_insignificant,
                                                 confidence, mi_host + (size_t)a * nbits,
                                                 counts_host ? counts_host + (size_t)a * nbits * 4 : nullptr, device))
                return rc;
        }
        return XBI_OK;
    }
    if (!x_host) return fail(XBI_ERR_ARG, "null data pointer");
    const size_t table = (size_t)nbits * 4;
    if (!c.counts_multi) {
        XBI_CUDA(cudaMalloc(&c.counts_multi, XBI_MAX_DIMS * 64 * 4 * sizeof(unsigned long long)), "cudaMalloc(counts)");
        XBI_CUDA(cudaMalloc(&c.mi_multi, XBI_MAX_DIMS * 64 * sizeof(double)), "cudaMalloc(mi)");
    }
    XBI_CUDA(cudaMemsetAsync(c.counts_multi, 0, naxes * table * sizeof(unsigned long long), c.stream[0]), "memset(counts)");
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
    // Slabs of whole rows of the first dimension, each slab on the device ONCE for all the axes.  Along the
    // first dimension a slab carries the first row of the next slab as halo (the pair across the cut belongs
    // to it); the other axes count the slab's own rows only.
    const bool stage = staging_enabled() && is_pageable(x_host);
    const int64_t per = cap / row - 1;  // own rows per slab (>= 1)
    const char* src = static_cast<const char*>(x_host);
    int which = 0;
    for (int64_t r0 = 0; r0 < shape[0]; r0 += per, which ^= 1) {
        const int64_t own = shape[0] - r0 < per ? shape[0] - r0 : per;
        const int64_t halo = r0 + own < shape[0] ? 1 : 0;
        cudaStream_t st = c.stream[which];
        const char* from = src + (size_t)r0 * row * bytes;
        const size_t nb = (size_t)(own + halo) * row * bytes;
        bool staged = false;
        if (stage)
            if (int rc = staged_h2d(c, c.buf[which], from, nb, 1, nb, st, &staged)) return rc;
        if (!staged) XBI_CUDA(cudaMemcpyAsync(c.buf[which], from, nb, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D)");
        for (int a = 0; a < naxes; ++a) {
            CountJob job = probe;
            job.x = c.buf[which];
            dims_of(axes[a], axes[a] == 0 ? own + halo : own, &job.outer, &job.n, &job.inner);
            Finish fin;
            fin.counts = c.counts_multi + (size_t)a * table;
            if (int rc = count_pairs_impl(job, flags, c.ws[which], st, fin)) return rc;
        }
    }
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
    XBI_CUDA(cudaStreamSynchronize(c.stream[1]), "sync");
    if (int rc = xbi_mutual_information_batched(c.counts_multi, nbits, naxes, set_zero_insignificant, confidence,
                                                c.mi_multi, c.stream[0]))
        return rc;
    XBI_CUDA(cudaMemcpyAsync(mi_host, c.mi_multi, (size_t)naxes * nbits * sizeof(double), cudaMemcpyDeviceToHost, c.stream[0]),
             "D2H(mi)");
    if (counts_host)
        XBI_CUDA(cudaMemcpyAsync(counts_host, c.counts_multi, (size_t)naxes * table * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, c.stream[0]),
                 "D2H(counts)");
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
    return XBI_OK;
}

int xbi_bitround_host(const void* src_host, void* dst_host, int dtype, int64_t numel, int keepbits, int device) {
    if (int rc = check_bitround_args(dtype, numel, keepbits)) return rc;
    if (numel == 0) return XBI_OK;
    if (!src_host || !dst_host) return fail(XBI_ERR_ARG, "null pointer");
    if (int rc = ensure_ctx(device)) return rc;
    HostCtx& c = g_ctx;
    const int bytes = dtype_bytes(dtype);
    const bool stage_in = staging_enabled() && is_pageable(src_host);
    const bool stage_out = staging_enabled() && is_pageable(dst_host);
    {
        const size_t all = (size_t)numel * bytes;
        if (int rc = ensure_staging((stage_in || stage_out) && all > HostCtx::kPinBytes ? HostCtx::kPinBytes : all)) return rc;
    }
    if (stage_in || stage_out) {
        // slabs of one pinned slot; slot k of each direction belongs to stream k.  While the GPU works on
        // slab i the pool copies slab i-1's result out and slab i+1 in.
        if (int rc = ensure_pinned(c, stage_in, stage_out)) return rc;
        if (stage_out && src_host != dst_host) {
            // a fresh output array is first touched here: with 2 MB pages the copy-out threads take 512x fewer
            // page faults (numpy gives the same advice for its own large allocations; malloc'ed buffers do not)
            const uintptr_t lo = (reinterpret_cast<uintptr_t>(dst_host) + ((uintptr_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1);
            const uintptr_t hi = (reinterpret_cast<uintptr_t>(dst_host) + (size_t)numel * bytes) & ~(((uintptr_t)2 << 20) - 1);
            if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
        }
        const size_t slot = HostCtx::kPinBytes < c.buf_bytes ? HostCtx::kPinBytes : c.buf_bytes;
        const int64_t per = (int64_t)(slot / bytes);
        struct Pending { bool active; char* dst; size_t nbytes; } pend[2] = {{false, nullptr, 0}, {false, nullptr, 0}};
        auto drain = [&](int k) -> int {
            if (!pend[k].active) return XBI_OK;
            XBI_CUDA(cudaEventSynchronize(c.ev_out[k]), "cudaEventSynchronize(staging out)");
            host_copy_parallel(pend[k].dst, pend[k].nbytes, c.pin_out[k], pend[k].nbytes, pend[k].nbytes, 1);
            pend[k].active = false;
            return XBI_OK;
        };
        int k = 0;
        for (int64_t off = 0; off < numel; off += per, k ^= 1) {
            const int64_t cnt = (numel - off < per) ? (numel - off) : per;
            const size_t nb = (size_t)cnt * bytes;
            cudaStream_t st = c.stream[k];
            if (int rc = drain(k)) return rc;
            const char* from = static_cast<const char*>(src_host) + (size_t)off * bytes;
            const void* h2d_src = from;
            if (stage_in) {
                if (c.in_busy[k]) XBI_CUDA(cudaEventSynchronize(c.ev_in[k]), "cudaEventSynchronize(staging in)");
                host_copy_parallel(c.pin_in[k], nb, from, nb, nb, 1);
                h2d_src = c.pin_in[k];
            }
            XBI_CUDA(cudaMemcpyAsync(c.buf[k], h2d_src, nb, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D)");
            if (stage_in) {
                XBI_CUDA(cudaEventRecord(c.ev_in[k], st), "cudaEventRecord(staging in)");
                c.in_busy[k] = true;
            }
            XBI_CUDA(launch_bitround(c.buf[k], dtype, cnt, keepbits, st), "bitround");
            char* to = static_cast<char*>(dst_host) + (size_t)off * bytes;
            if (stage_out) {
                XBI_CUDA(cudaMemcpyAsync(c.pin_out[k], c.buf[k], nb, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H, staged)");
                XBI_CUDA(cudaEventRecord(c.ev_out[k], st), "cudaEventRecord(staging out)");
                pend[k] = {true, to, nb};
            } else {
                XBI_CUDA(cudaMemcpyAsync(to, c.buf[k], nb, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
            }
        }
        if (int rc = drain(0)) return rc;
        if (int rc = drain(1)) return rc;
        XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
        XBI_CUDA(cudaStreamSynchronize(c.stream[1]), "sync");
        return XBI_OK;
    }
    const int64_t cap = (int64_t)(c.buf_bytes / bytes);
    int which = 0;
    for (int64_t off = 0; off < numel; off += cap) {
        const int64_t k = (numel - off < cap) ? (numel - off) : cap;
        cudaStream_t st = c.stream[which];
        XBI_CUDA(cudaMemcpyAsync(c.buf[which], static_cast<const char*>(src_host) + (size_t)off * bytes,
                                 (size_t)k * bytes, cudaMemcpyHostToDevice, st),
                 "cudaMemcpyAsync(H2D)");
        XBI_CUDA(launch_bitround(c.buf[which], dtype, k, keepbits, st), "bitround");
        XBI_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_host) + (size_t)off * bytes, c.buf[which], (size_t)k * bytes,
                                 cudaMemcpyDeviceToHost, st),
                 "cudaMemcpyAsync(D2H)");
        which ^= 1;
    }
    XBI_CUDA(cudaStreamSynchronize(c.stream[0]), "sync");
    XBI_CUDA(cudaStreamSynchronize(c.stream[1]), "sync");
    return XBI_OK;
}

void xbi_host_release(void) { g_ctx.release(); }

}  // extern "C"
