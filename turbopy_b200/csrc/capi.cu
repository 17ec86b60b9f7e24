// Library-level entry points: version, errors, device info, CUDA-graph capture,
// launch accounting, and the on-device self test of the division helper.
#include "common.cuh"

#include <cstdio>

namespace tp {

int64_t g_launch_count = 0;

const DeviceProps& device_props() {
    static DeviceProps props = [] {
        DeviceProps d;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            return d;
        }
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return d;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return d;
        d.device = dev;
        d.sm_count = p.multiProcessorCount;
        d.cc_major = p.major;
        d.cc_minor = p.minor;
        d.l2_bytes = p.l2CacheSize;
        d.smem_optin = static_cast<int64_t>(p.sharedMemPerBlockOptin);
        // The kernels are built for sm_100a only; anything else cannot run them.
        d.status = (p.major == 10) ? TP_OK : TP_E_NODEVICE;
        return d;
    }();
    return props;
}

__global__ void selftest_division_kernel(int64_t n, uint64_t seed, unsigned long long* mismatches) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned long long bad = 0, taken = 0;
    for (int64_t i = gtid; i < n; i += gstride) {
        // splitmix64 stream per pair
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(2 * i + 1);
        auto next = [&z]() {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            return x ^ (x >> 31);
        };
        uint64_t ma = next() & 0xFFFFFFFFFFFFFull, mb = next() & 0xFFFFFFFFFFFFFull;
        const uint64_t sel = next();
        int mode = static_cast<int>(sel & 15);
        if (mode == 1) mb = 0xFFFFFFFFFFFFFull - (mb & 0xff);   // denominator mantissa ~ all ones
        if (mode == 2) mb &= 0xfff;                             // denominator ~ power of two
        if (mode == 3) ma = 0xFFFFFFFFFFFFFull - (ma & 0xff);
        if (mode == 4) ma &= 0xffff;
        int ea = static_cast<int>((sel >> 8) % 200) - 100;
        int eb = static_cast<int>((sel >> 24) % 200) - 100;
        if (mode == 5) ea = -1000 - static_cast<int>((sel >> 40) % 22);   // tiny numerator (slow path)
        if (mode == 6) eb = 1000 + static_cast<int>((sel >> 40) % 23);    // huge denominator
        if (mode == 7) eb = -1000 - static_cast<int>((sel >> 40) % 22);   // tiny denominator
        double a = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(1023 + ea) << 52) | ma));
        double b = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(1023 + eb) << 52) | mb));
        if (sel & (1ull << 60)) a = -a;
        if (sel & (1ull << 61)) b = -b;
        if (mode == 8) a = (sel & (1ull << 62)) ? -0.0 : 0.0;            // signed zero numerator
        if (mode == 9) a = __longlong_as_double(0x7ff0000000000000ll);   // inf numerator
        if (mode == 10) b = __longlong_as_double(0x7ff0000000000000ll);  // inf denominator
        if (mode == 11) b = 0.0;
#ifndef TP_GUARD_STRICT
        if (mode == 12) mode = 13;                                       // the default build does not claim the hi == 0 hole (common.cuh)
#endif
        if (mode == 12 || mode == 13) {                                  // subnormal numerators; 12: high word zero (< 2^-1042)
            uint64_t m = (mode == 12) ? (ma & 0xFFFFFFFFull) : (ma | 0x100000000ull);
            if (m == 0) m = 1;
            a = __longlong_as_double(static_cast<long long>(m | ((sel & (1ull << 60)) ? 0x8000000000000000ull : 0ull)));
            b = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(1023 + static_cast<int>((sel >> 24) % 3)) << 52) | mb));
        }
        // The contract of the FAST policy: whenever the guard and the finiteness
        // test pass (b > 0 as for every denominator of the hot path), the quotient
        // has the bits of IEEE a/b; otherwise the caller recomputes with a/b.
        b = fabs(b);
        Guard g;
        const Recip<true> r = make_recip<true>(b, g);
        const double q = quot<true>(a, r, g);
        const double t = a / b;
        const bool accepted = guard_ok(g) && finite3(q, q, q);
        if (accepted && __double_as_longlong(q) != __double_as_longlong(t)) ++bad;
        if (accepted) ++taken;
    }
    if (bad) atomicAdd(mismatches, bad);
    atomicAdd(mismatches + 1, taken);
}

// rcp_inrange / sqrt_inrange against the IEEE library results (1.0/b, sqrt(a)).
__global__ void selftest_roots_kernel(int64_t n, uint64_t seed, unsigned long long* out) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned long long bad = 0, taken = 0;
    for (int64_t i = gtid; i < n; i += gstride) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(2 * i + 1);
        auto next = [&z]() {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            return x ^ (x >> 31);
        };
        uint64_t m = next() & 0xFFFFFFFFFFFFFull;
        const uint64_t sel = next();
        int mode = static_cast<int>(sel & 15);
        if (mode == 1) m = 0xFFFFFFFFFFFFFull - (m & 0xff);     // mantissa ~ all ones
        if (mode == 2) m &= 0xfff;                              // ~ power of two
        if (mode == 3) m = (m & 0xff) | 0x6A09E667F3BCDull;     // ~ sqrt(2) pattern
        int e = static_cast<int>((sel >> 8) % 1200) - 600;      // spans inside and outside [2^-500, 2^500]
        if (mode == 4) e = static_cast<int>((sel >> 8) % 2046) - 1022;
        double v = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(1023 + e) << 52) | m));
        if (mode == 5) v = 0.0;
        if (mode == 6) v = __longlong_as_double(0x7ff0000000000000ll);
        if (mode == 7) v = -v;
        Guard g;
        guard_range(g, v);
        const double s = sqrt_inrange(v);
        const double rcp = rcp_inrange(v);
        Guard gr;
        guard_range(gr, rcp);
        if (guard_ok(g)) {                                      // sqrt argument accepted
            ++taken;
            const double ref = sqrt(v);                         // negative argument: both NaN (payload may differ;
            const bool same = __double_as_longlong(s) == __double_as_longlong(ref) || (s != s && ref != ref);
            if (!same) ++bad;                                   //  the kernels treat any NaN as "recompute")
        }
        if (guard_ok(gr) && v > 0.0) {                          // reciprocal accepted
            ++taken;
            if (__double_as_longlong(rcp) != __double_as_longlong(1.0 / v)) ++bad;
        }
    }
    if (bad) atomicAdd(out, bad);
    atomicAdd(out + 1, taken);
}

}  // namespace tp

using namespace tp;

extern "C" int tp_selftest_roots(int64_t n, uint64_t seed, int64_t* mismatches, int64_t* fast_taken,
                                 tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!mismatches || n < 0) return TP_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned long long* dev = nullptr;
    TP_CUDA_TRY(cudaMalloc(&dev, 2 * sizeof(unsigned long long)));
    cudaMemsetAsync(dev, 0, 2 * sizeof(unsigned long long), s);
    selftest_roots_kernel<<<d.sm_count * 4, 256, 0, s>>>(n, seed, dev);
    count_launch();
    unsigned long long host[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(host), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dev);
    *mismatches = static_cast<int64_t>(host[0]);
    if (fast_taken) *fast_taken = static_cast<int64_t>(host[1]);
    return static_cast<int>(e);
}

extern "C" int tp_version(void) { return TP_VERSION; }

extern "C" const char* tp_error_string(int code) {
    switch (code) {
        case TP_OK: return "ok";
        case TP_E_ARG: return "turbopy_b200: bad argument";
        case TP_E_LAYOUT: return "turbopy_b200: unsupported array strides";
        case TP_E_NODEVICE: return "turbopy_b200: no sm_100 (B200) CUDA device available; there is no CPU fallback";
        case TP_E_GRID: return "turbopy_b200: invalid grid (need >= 2 nodes, dr > 0)";
        case TP_E_MODE: return "turbopy_b200: unknown field mode / formula / axis";
        case TP_E_GRAPH: return "turbopy_b200: CUDA graph capture misuse";
        case TP_E_NCCL: return "turbopy_b200: libnccl.so.2 could not be loaded";
        default: break;
    }
    if (code < TP_E_NCCL && code > TP_E_NCCL - 64) return "turbopy_b200: NCCL call failed (ncclResult_t = -100 - code)";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "turbopy_b200: unknown error";
}

extern "C" int tp_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes,
                              int64_t* smem_optin_bytes) {
    const DeviceProps& d = device_props();
    if (d.device < 0) return TP_E_NODEVICE;
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    if (l2_bytes) *l2_bytes = d.l2_bytes;
    if (smem_optin_bytes) *smem_optin_bytes = d.smem_optin;
    return d.status;
}

extern "C" int64_t tp_launch_count(void) { return g_launch_count; }

extern "C" int tp_selftest_division(int64_t n, uint64_t seed, int64_t* mismatches, int64_t* fast_taken,
                                    tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!mismatches || n < 0) return TP_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned long long* dev = nullptr;
    TP_CUDA_TRY(cudaMalloc(&dev, 2 * sizeof(unsigned long long)));
    cudaMemsetAsync(dev, 0, 2 * sizeof(unsigned long long), s);
    selftest_division_kernel<<<d.sm_count * 4, 256, 0, s>>>(n, seed, dev);
    count_launch();
    unsigned long long host[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(host), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dev);
    *mismatches = static_cast<int64_t>(host[0]);
    if (fast_taken) *fast_taken = static_cast<int64_t>(host[1]);
    return static_cast<int>(e);
}

// ---- CUDA-graph capture ------------------------------------------------------
// Thin wrappers so a host language without CUDA bindings can record a sequence
// of the calls above (e.g. K gather+push steps, or gather+push + reduction) and
// replay it with one launch.  A replay launches the captured kernels again; the
// launch counter is advanced by the number of kernel nodes.
namespace {
struct GraphExec {
    cudaGraphExec_t exec = nullptr;
    int kernel_nodes = 0;
    int64_t count_at_begin = 0;
};
int64_t g_capture_count_at_begin = 0;
}  // namespace

extern "C" int tp_graph_begin(tp_stream_t stream) {
    g_capture_count_at_begin = g_launch_count;
    return static_cast<int>(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal));
}

extern "C" int tp_graph_end(tp_stream_t stream, void** graph_exec) {
    if (!graph_exec) return TP_E_ARG;
    cudaGraph_t graph = nullptr;
    TP_CUDA_TRY(cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &graph));
    if (!graph) return TP_E_GRAPH;
    GraphExec* g = new GraphExec;
    cudaError_t e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { delete g; return static_cast<int>(e); }
    g->kernel_nodes = static_cast<int>(g_launch_count - g_capture_count_at_begin);
    g_launch_count = g_capture_count_at_begin;      // captured launches did not run
    *graph_exec = g;
    return TP_OK;
}

extern "C" int tp_graph_launch(void* graph_exec, tp_stream_t stream) {
    if (!graph_exec) return TP_E_ARG;
    GraphExec* g = static_cast<GraphExec*>(graph_exec);
    TP_CUDA_TRY(cudaGraphLaunch(g->exec, static_cast<cudaStream_t>(stream)));
    count_launch(g->kernel_nodes);
    return TP_OK;
}

extern "C" int tp_graph_destroy(void* graph_exec) {
    if (!graph_exec) return TP_OK;
    GraphExec* g = static_cast<GraphExec*>(graph_exec);
    cudaError_t e = g->exec ? cudaGraphExecDestroy(g->exec) : cudaSuccess;
    delete g;
    return static_cast<int>(e);
}
