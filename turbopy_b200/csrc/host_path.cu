// End-to-end calls on HOST buffers: the path a turboPy user with numpy arrays
// takes (BorisPush.push(position, momentum, ...) on (n,3) host arrays,
// turbopy/computetools.py:407).  The particle arrays are streamed through the
// device in chunks on three internal streams so that the H2D copy of chunk c+1,
// the kernel of chunk c and the D2H copy of chunk c-1 overlap (PCIe is full
// duplex); the arithmetic is done by the same device kernels as the resident
// path -- there is no CPU compute here.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace tp {

int gather_push_device(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g, int axis,
                       int formula, uint64_t* status, cudaStream_t s, int nsteps, int32_t* bounds, int64_t bounds_len,
                       double* rec_scratch, const struct MomRequest* mr);

namespace {

constexpr int kBufs = 6;                       // most buffers / streams in flight; host_bufs() of them are used
// Pipeline depth: chunks in flight, each on its own stream (H2D -> kernel -> D2H).  Three keep both copy directions
// busy when the link is symmetric (one GPU: 95 % of the duplex ceiling); with several ranks on one host the two
// directions run at different, fluctuating rates and a deeper pipeline absorbs it (tunable: TURBOPY_B200_HOST_BUFS).
inline int host_bufs() {
    static const int n = [] {
        const char* e = std::getenv("TURBOPY_B200_HOST_BUFS");
        const int v = e ? std::atoi(e) : 4;
        return std::max(2, std::min(kBufs, v));
    }();
    return n;
}
constexpr int64_t kChunk = 4ll << 20;          // most particles per chunk (192 MB of state)
constexpr int64_t kMinChunk = 256ll << 10;

// Chunk length: about n/16 (tunable: TURBOPY_B200_HOST_CHUNKS; 8/16/32/64 parts measured
// 0.91/0.93/0.90/0.84 G pushes/s end to end at 16 Mi particles) so that copies and
// kernels of different chunks overlap and the un-overlapped fill / drain of the
// pipeline (first H2D, last D2H) stays a small part of the call; within
// [256 Ki, 4 Mi] particles, even (keeps SoA blocks 16-B aligned).
inline int64_t chunk_for(int64_t n) {
    static const int parts = [] {
        const char* e = std::getenv("TURBOPY_B200_HOST_CHUNKS");
        const int v = e ? std::atoi(e) : 16;
        return v >= 1 ? v : 16;
    }();
    int64_t c = (n + parts - 1) / parts;
    c = std::max(kMinChunk, std::min(kChunk, c));
    c = std::min(c, std::max<int64_t>(n, 2));
    return ((c + 1) / 2) * 2;
}

struct HostCtx {
    cudaStream_t stream[kBufs] = {};
    double* dev[kBufs] = {};                            // 12 * cap doubles each: pos, mom, E, B
    cudaEvent_t grid_ready = nullptr;                   // the uploaded grid (stream 0) is complete
    int64_t cap = 0;
    double* grid_dev = nullptr; int64_t grid_cap = 0;   // uploaded grid (r + components)
    uint64_t* status_dev = nullptr;
    int64_t h2d = 0, d2h = 0;
    bool init = false;
};
// One context (streams + staging buffers) per device, created on the device that is current at the first host
// call on it; g_host_mutex serialises the host entry points (ctypes releases the GIL around them).
constexpr int kMaxDevices = 16;
HostCtx g_ctxs[kMaxDevices];
std::mutex g_host_mutex;
int g_last_device = 0;

HostCtx& current_ctx() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    g_last_device = dev;
    return g_ctxs[dev];
}

int ensure_ctx(int64_t chunk, int64_t grid_doubles) {
    HostCtx& c = current_ctx();
    if (!c.init) {
        for (int b = 0; b < kBufs; ++b) TP_CUDA_TRY(cudaStreamCreateWithFlags(&c.stream[b], cudaStreamNonBlocking));
        TP_CUDA_TRY(cudaMalloc(&c.status_dev, 4 * sizeof(uint64_t)));
        TP_CUDA_TRY(cudaEventCreateWithFlags(&c.grid_ready, cudaEventDisableTiming));
        c.init = true;
    }
    if (chunk > c.cap) {
        for (int b = 0; b < kBufs; ++b) {
            if (c.dev[b]) cudaFree(c.dev[b]);
            c.dev[b] = nullptr;
            if (b < host_bufs()) TP_CUDA_TRY(cudaMalloc(&c.dev[b], sizeof(double) * 12 * chunk));
        }
        c.cap = chunk;
    }
    if (grid_doubles > c.grid_cap) {
        if (c.grid_dev) cudaFree(c.grid_dev);
        c.grid_dev = nullptr;
        TP_CUDA_TRY(cudaMalloc(&c.grid_dev, sizeof(double) * grid_doubles));
        c.grid_cap = grid_doubles;
    }
    return TP_OK;
}

// Copy a chunk [i0, i0+m) of an (n,3) host array to/from a device block of 3*ld doubles.
// SoA host arrays are copied component by component (device block is SoA with
// ld = chunk length); AoS host arrays are one contiguous copy (device block AoS).
int copy_chunk(double* dev, double* host, int64_t sn, int64_t sc, int64_t i0, int64_t m, int64_t ld, bool to_dev,
               cudaStream_t s, int64_t& bytes) {
    const cudaMemcpyKind kind = to_dev ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (sn == 3 && sc == 1) {
        double* h = host + 3 * i0;
        TP_CUDA_TRY(to_dev ? cudaMemcpyAsync(dev, h, sizeof(double) * 3 * m, kind, s)
                           : cudaMemcpyAsync(h, dev, sizeof(double) * 3 * m, kind, s));
    } else {
        for (int c = 0; c < 3; ++c) {
            double* h = host + c * sc + i0;
            double* d = dev + c * ld;
            TP_CUDA_TRY(to_dev ? cudaMemcpyAsync(d, h, sizeof(double) * m, kind, s)
                               : cudaMemcpyAsync(h, d, sizeof(double) * m, kind, s));
        }
    }
    bytes += static_cast<int64_t>(sizeof(double)) * 3 * m;
    return TP_OK;
}

int check_host_layout(const tp_particles* p) {
    if (!p || p->n < 0 || (p->n > 0 && (!p->pos || !p->mom))) return TP_E_ARG;
    const int lp = layout_of(p->pos_stride_n, p->pos_stride_c, p->n);
    const int lm = layout_of(p->mom_stride_n, p->mom_stride_c, p->n);
    if (lp < 0 || lm < 0) return TP_E_LAYOUT;
    return TP_OK;
}

// Shared driver: `launch(chunk particles, buffer index, first particle, stream)` runs the kernel.
template <typename Launch>
int run_chunks_unsafe(const tp_particles* p, const double* E_host, bool e_pp, const double* B_host, bool b_pp,
                      Launch launch);

// On ANY error the three streams are drained before returning: asynchronous copies issued earlier may still be
// targeting the caller's arrays, which the caller is free to release as soon as the call has failed.
template <typename Launch>
int run_chunks(const tp_particles* p, const double* E_host, bool e_pp, const double* B_host, bool b_pp,
               Launch launch) {
    const int rc = run_chunks_unsafe(p, E_host, e_pp, B_host, b_pp, launch);
    if (rc != TP_OK) {
        HostCtx& c = current_ctx();
        for (int b = 0; b < kBufs; ++b)
            if (c.stream[b]) cudaStreamSynchronize(c.stream[b]);
        cudaGetLastError();
    }
    return rc;
}

template <typename Launch>
int run_chunks_unsafe(const tp_particles* p, const double* E_host, bool e_pp, const double* B_host, bool b_pp,
                      Launch launch) {
    HostCtx& c = current_ctx();
    const int64_t n = p->n;
    const int64_t chunk = chunk_for(n);
    int rc = TP_OK;
    int idx = 0;
    for (int64_t i0 = 0; i0 < n; i0 += chunk, ++idx) {
        const int b = idx % host_bufs();
        const int64_t m = std::min(chunk, n - i0);
        cudaStream_t s = c.stream[b];
        double* base = c.dev[b];
        if ((rc = copy_chunk(base, p->pos, p->pos_stride_n, p->pos_stride_c, i0, m, chunk, true, s, c.h2d))) return rc;
        if ((rc = copy_chunk(base + 3 * chunk, p->mom, p->mom_stride_n, p->mom_stride_c, i0, m, chunk, true, s, c.h2d)))
            return rc;
        if (e_pp && (rc = copy_chunk(base + 6 * chunk, const_cast<double*>(E_host), 3, 1, i0, m, chunk, true, s, c.h2d)))
            return rc;
        if (b_pp && (rc = copy_chunk(base + 9 * chunk, const_cast<double*>(B_host), 3, 1, i0, m, chunk, true, s, c.h2d)))
            return rc;
        tp_particles dp;
        const bool pos_aos = p->pos_stride_n == 3, mom_aos = p->mom_stride_n == 3;
        dp.pos = base;              dp.pos_stride_n = pos_aos ? 3 : 1; dp.pos_stride_c = pos_aos ? 1 : chunk;
        dp.mom = base + 3 * chunk;  dp.mom_stride_n = mom_aos ? 3 : 1; dp.mom_stride_c = mom_aos ? 1 : chunk;
        dp.n = m;
        if ((rc = launch(dp, base, chunk, s))) return rc;
        if ((rc = copy_chunk(base, p->pos, p->pos_stride_n, p->pos_stride_c, i0, m, chunk, false, s, c.d2h))) return rc;
        if ((rc = copy_chunk(base + 3 * chunk, p->mom, p->mom_stride_n, p->mom_stride_c, i0, m, chunk, false, s, c.d2h)))
            return rc;
    }
    for (int b = 0; b < kBufs; ++b) TP_CUDA_TRY(cudaStreamSynchronize(c.stream[b]));
    return TP_OK;
}

}  // namespace
}  // namespace tp

using namespace tp;

extern "C" int tp_boris_push_host_f64(const tp_particles* p, const tp_push_consts* k, const double* E_host,
                                      int e_per_particle, const double* B_host, int b_per_particle) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    int rc = check_host_layout(p);
    if (rc) return rc;
    if (!k || !E_host || !B_host) return TP_E_ARG;
    if (p->n == 0) return TP_OK;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    const int64_t chunk = chunk_for(p->n);
    if ((rc = ensure_ctx(chunk, 0))) return rc;
    current_ctx().h2d = current_ctx().d2h = 0;
    return run_chunks(p, E_host, e_per_particle != 0, B_host, b_per_particle != 0,
                      [&](const tp_particles& dp, double* base, int64_t ch, cudaStream_t s) {
                          const double* E = e_per_particle ? base + 6 * ch : E_host;
                          const double* B = b_per_particle ? base + 9 * ch : B_host;
                          return tp_boris_push_f64(&dp, k, E, e_per_particle ? TP_FIELD_PER_PARTICLE : TP_FIELD_UNIFORM_HOST,
                                                   3, 1, B, b_per_particle ? TP_FIELD_PER_PARTICLE : TP_FIELD_UNIFORM_HOST,
                                                   3, 1, s);
                      });
}

extern "C" int tp_gather_push_host_f64(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g,
                                       int axis, int formula, uint64_t* status_host) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    int rc = check_host_layout(p);
    if (rc) return rc;
    if (!k || !g || !g->r) return TP_E_ARG;
    if (g->ng < 2) return TP_E_GRID;
    // Upload the grid: r, then the field components.  A component with node stride 1 is one contiguous copy.
    // Strided components (the columns of a Grid.generate_field(3) array: stride 3) are NOT copied row by row (a 2-D
    // copy of 8-byte rows runs at a few hundred MB/s): the memory ranges they live in are merged (columns of one
    // (ng,3) array share a range) and each range is uploaded once, contiguously; the device reads it with the same
    // stride.  Ranges whose stride is so large that the gaps dominate fall back to the 2-D copy.
    struct Range { uintptr_t lo, hi; };
    Range rng[6]; int nr = 0;
    for (int cc = 0; cc < 6; ++cc) {
        if (!g->comp[cc] || g->comp_stride[cc] == 1) continue;
        if (g->comp_stride[cc] < 1 || g->comp_stride[cc] > 8) continue;          // sparse: 2-D copy below
        const uintptr_t lo = reinterpret_cast<uintptr_t>(g->comp[cc]);
        rng[nr++] = {lo, lo + (static_cast<uintptr_t>(g->ng - 1) * g->comp_stride[cc] + 1) * sizeof(double)};
    }
    std::sort(rng, rng + nr, [](const Range& x, const Range& y) { return x.lo < y.lo; });
    Range seg[6]; int ns = 0;
    for (int i = 0; i < nr; ++i) {
        if (ns && rng[i].lo <= seg[ns - 1].hi) seg[ns - 1].hi = std::max(seg[ns - 1].hi, rng[i].hi);
        else seg[ns++] = rng[i];
    }
    int ndense = 0;
    for (int cc = 0; cc < 6; ++cc)
        ndense += (g->comp[cc] && (g->comp_stride[cc] == 1 || g->comp_stride[cc] > 8)) ? 1 : 0;
    const int64_t ngp = (g->ng + 1) & ~1ll;                        // keep every block 16-byte aligned
    int64_t seg_doubles = 0;
    for (int i = 0; i < ns; ++i) seg_doubles += ((static_cast<int64_t>(seg[i].hi - seg[i].lo) / 8) + 2) & ~1ll;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    const int64_t chunk = chunk_for(p->n);
    if ((rc = ensure_ctx(chunk, ngp * (1 + ndense) + seg_doubles))) return rc;
    HostCtx& c = current_ctx();
    c.h2d = c.d2h = 0;
    cudaStream_t s0 = c.stream[0];
    tp_grid_fields dg = *g;
    dg.r = c.grid_dev;
    TP_CUDA_TRY(cudaMemcpyAsync(c.grid_dev, g->r, sizeof(double) * g->ng, cudaMemcpyHostToDevice, s0));
    c.h2d += sizeof(double) * g->ng;
    double* next = c.grid_dev + ngp;
    double* seg_dev[6];
    for (int i = 0; i < ns; ++i) {
        const int64_t nd = static_cast<int64_t>(seg[i].hi - seg[i].lo) / 8;
        seg_dev[i] = next;
        TP_CUDA_TRY(cudaMemcpyAsync(next, reinterpret_cast<const void*>(seg[i].lo), sizeof(double) * nd,
                                    cudaMemcpyHostToDevice, s0));
        c.h2d += sizeof(double) * nd;
        next += (nd + 2) & ~1ll;
    }
    for (int cc = 0; cc < 6; ++cc) {
        if (!g->comp[cc]) continue;
        const int64_t st = g->comp_stride[cc];
        if (st != 1 && st >= 1 && st <= 8) {                        // lives in one of the uploaded ranges
            const uintptr_t u = reinterpret_cast<uintptr_t>(g->comp[cc]);
            for (int i = 0; i < ns; ++i)
                if (u >= seg[i].lo && u < seg[i].hi) dg.comp[cc] = seg_dev[i] + (u - seg[i].lo) / 8;
            dg.comp_stride[cc] = st;
            continue;
        }
        double* dst = next;
        next += ngp;
        if (st == 1) {
            TP_CUDA_TRY(cudaMemcpyAsync(dst, g->comp[cc], sizeof(double) * g->ng, cudaMemcpyHostToDevice, s0));
        } else {
            TP_CUDA_TRY(cudaMemcpy2DAsync(dst, sizeof(double), g->comp[cc], sizeof(double) * st,
                                          sizeof(double), g->ng, cudaMemcpyHostToDevice, s0));
        }
        c.h2d += sizeof(double) * g->ng;
        dg.comp[cc] = dst;
        dg.comp_stride[cc] = 1;
    }
    if ((rc = tp_status_reset(c.status_dev, s0))) return rc;
    // the chunk pipeline starts right away: its particle uploads overlap the grid upload, only the KERNELS wait for it
    TP_CUDA_TRY(cudaEventRecord(c.grid_ready, s0));
    if (p->n > 0) {
        rc = run_chunks(p, nullptr, false, nullptr, false,
                        [&](const tp_particles& dp, double*, int64_t, cudaStream_t s) {
                            if (s != s0) TP_CUDA_TRY(cudaStreamWaitEvent(s, c.grid_ready, 0));
                            return gather_push_device(&dp, k, &dg, axis, formula, c.status_dev, s, 1, nullptr, 0, nullptr, nullptr);
                        });
        if (rc) return rc;
    }
    if (status_host) {
        TP_CUDA_TRY(cudaMemcpy(status_host, c.status_dev, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        c.d2h += 4 * sizeof(uint64_t);
    }
    return TP_OK;
}

// Page-lock / unlock a caller-owned host range so that the chunk copies above run as true asynchronous DMA
// (pageable numpy arrays: 0.14 G pushes/s end to end; the same arrays page-locked: 0.93).  The LIFETIME of the
// registration is the caller's business: the Python layer ties it to the owning ndarray with a weakref finalizer.
extern "C" int tp_host_pin(void* ptr, int64_t bytes) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!ptr || bytes <= 0) return TP_E_ARG;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost) return TP_OK;   // already page-locked
    cudaGetLastError();
    const cudaError_t rc = cudaHostRegister(ptr, static_cast<size_t>(bytes), cudaHostRegisterDefault);
    if (rc != cudaSuccess) cudaGetLastError();                  // not fatal: the copies fall back to staged transfers
    return static_cast<int>(rc);
}

extern "C" int tp_host_unpin(void* ptr) {
    if (!ptr) return TP_E_ARG;
    const cudaError_t rc = cudaHostUnregister(ptr);
    if (rc != cudaSuccess) cudaGetLastError();
    return static_cast<int>(rc);
}

extern "C" int tp_host_last_transfer(int64_t* h2d_bytes, int64_t* d2h_bytes) {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    if (h2d_bytes) *h2d_bytes = g_ctxs[g_last_device].h2d;
    if (d2h_bytes) *d2h_bytes = g_ctxs[g_last_device].d2h;
    return TP_OK;
}

extern "C" int tp_host_release(void) {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    int keep = 0;
    cudaGetDevice(&keep);
    for (int dev = 0; dev < kMaxDevices; ++dev) {
        HostCtx& c = g_ctxs[dev];
        if (!c.init) continue;
        cudaSetDevice(dev);
        for (int b = 0; b < kBufs; ++b) {
            if (c.dev[b]) cudaFree(c.dev[b]);
            c.dev[b] = nullptr;
            if (c.stream[b]) cudaStreamDestroy(c.stream[b]);
            c.stream[b] = nullptr;
        }
        if (c.grid_dev) cudaFree(c.grid_dev);
        if (c.status_dev) cudaFree(c.status_dev);
        if (c.grid_ready) cudaEventDestroy(c.grid_ready);
        c.grid_dev = nullptr; c.status_dev = nullptr; c.grid_ready = nullptr; c.cap = 0; c.grid_cap = 0; c.init = false;
    }
    cudaSetDevice(keep);
    return TP_OK;
}

// What the host link of THIS device sustains right now: pinned host <-> device cudaMemcpyAsync of `bytes` per
// direction, `iters` times, H2D alone, D2H alone and both at once on two streams (GB/s each, CUDA events).  It is
// the ceiling of the end-to-end host path (48 B in + 48 B out per push); run on all ranks at once it shows what the
// box's host memory / PCIe fabric gives each GPU under that load.
extern "C" int tp_host_link_probe(int64_t bytes, int iters, double* h2d_gbs, double* d2h_gbs, double* duplex_gbs) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (bytes <= 0 || iters <= 0) return TP_E_ARG;
    void *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    auto cleanup = [&](int rc) {
        if (h0) cudaFreeHost(h0); if (h1) cudaFreeHost(h1);
        if (d0) cudaFree(d0); if (d1) cudaFree(d1);
        if (s0) cudaStreamDestroy(s0); if (s1) cudaStreamDestroy(s1);
        if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (e2) cudaEventDestroy(e2);
        return rc;
    };
#define TP_PROBE_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cleanup(static_cast<int>(_e)); } while (0)
    TP_PROBE_TRY(cudaHostAlloc(&h0, bytes, cudaHostAllocDefault));
    TP_PROBE_TRY(cudaHostAlloc(&h1, bytes, cudaHostAllocDefault));
    std::memset(h0, 1, bytes); std::memset(h1, 2, bytes);          // first touch on the calling thread's NUMA node
    TP_PROBE_TRY(cudaMalloc(&d0, bytes));
    TP_PROBE_TRY(cudaMalloc(&d1, bytes));
    TP_PROBE_TRY(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    TP_PROBE_TRY(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    TP_PROBE_TRY(cudaEventCreate(&e0)); TP_PROBE_TRY(cudaEventCreate(&e1)); TP_PROBE_TRY(cudaEventCreate(&e2));
    auto timed = [&](bool up, bool down, double* out) -> int {
        for (int pass = 0; pass < 2; ++pass) {                      // pass 0 warms up
            cudaEventRecord(e0, s0);
            cudaStreamWaitEvent(s1, e0, 0);
            for (int i = 0; i < iters; ++i) {
                if (up) cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, s0);
                if (down) cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, s1);
            }
            cudaEventRecord(e2, s1);
            cudaStreamWaitEvent(s0, e2, 0);
            cudaEventRecord(e1, s0);
            const cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (out) *out = static_cast<double>(bytes) * iters / (ms * 1e-3) / 1e9;     // per direction
        return TP_OK;
    };
    int rc;
    if ((rc = timed(true, false, h2d_gbs)) || (rc = timed(false, true, d2h_gbs)) || (rc = timed(true, true, duplex_gbs)))
        return cleanup(rc);
#undef TP_PROBE_TRY
    return cleanup(TP_OK);
}
