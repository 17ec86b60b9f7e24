// PoissonSolver1DRadial.solve (turbopy/computetools.py:35-59) as
// two fp64 prefix scans.  Line-by-line spec: oracle/poisson.py.
//
//   a[i]  = (r[i]*s[i])*dr                      I1 = cumsum(a)
//   g[i]  = (I1[i]*dr)/r[i],  g[0] := i0 = 2*g[1] - g[2],  g -= i0
//   I2    = cumsum(g),        out = I2 - I2[n-1]
//
// Each scan is reduce-then-scan over 2048-element tiles (per-tile sums, one-CTA
// scan of the tile sums, per-tile scan with the tile's offset): no CTA ever
// waits on another one.  numpy's cumsum is strictly sequential, so results
// agree to rounding (tolerance in tests/test_gpu_poisson.py), not bit for bit.
// Traffic: 80 B/point over the five passes (r,s | r,s,I1 | I1,r,I2 | I2,out).
#include "common.cuh"

#include <algorithm>

namespace tp {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct PoissonArgs {
    const double* r; const double* s; double* I1; double* out;
    double* tile_sums;          // ntiles doubles (reused by both scans)
    double* scalars;            // [0] = i0, [1] = total of the second scan
    int64_t n; int64_t ntiles; double dr;
};

__device__ __forceinline__ double block_exclusive_scan(double v, double& total) {
    __shared__ double warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    double wbase = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        if (w < warp) wbase += warp_tot[w];
        tot += warp_tot[w];
    }
    __syncthreads();
    total = tot;
    return wbase + (inc - v);
}

// element of the first scan's input
__device__ __forceinline__ double a_of(const PoissonArgs& p, int64_t i) { return (p.r[i] * p.s[i]) * p.dr; }

// element of the second scan's input (needs I1 and i0)
__device__ __forceinline__ double g_of(const PoissonArgs& p, int64_t i, double i0) {
    if (i == 0) return i0 - i0;
    return (p.I1[i] * p.dr) / p.r[i] - i0;
}

template <int PASS>      // 1: a[], 2: g[]
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const __grid_constant__ PoissonArgs p) {
    const double i0 = PASS == 2 ? p.scalars[0] : 0.0;
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int64_t base = t * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int64_t i = base + k;
            if (i < p.n) acc += PASS == 1 ? a_of(p, i) : g_of(p, i, i0);
        }
        double total;
        block_exclusive_scan(acc, total);
        if (threadIdx.x == 0) p.tile_sums[t] = total;
        __syncthreads();
    }
}

// one CTA: exclusive scan of the tile sums in place; PASS 1 also forms i0 from
// the first three elements in numpy's sequential order.
template <int PASS>
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(const __grid_constant__ PoissonArgs p) {
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < p.ntiles; base += kScanThreads) {
        const int64_t t = base + threadIdx.x;
        const double v = t < p.ntiles ? p.tile_sums[t] : 0.0;
        double total;
        const double ex = block_exclusive_scan(v, total);
        const double carry = carry_s;
        if (t < p.ntiles) p.tile_sums[t] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (PASS == 1) {
            // i0 = 2*g[1] - g[2] with I1[1] = a0 + a1, I1[2] = I1[1] + a2  (computetools.py:50-53)
            double i0 = 0.0;
            if (p.n >= 3) {
                const double I1_1 = a_of(p, 0) + a_of(p, 1);
                const double I1_2 = I1_1 + a_of(p, 2);
                i0 = 2.0 * ((I1_1 * p.dr) / p.r[1]) - (I1_2 * p.dr) / p.r[2];
            }
            p.scalars[0] = i0;
        } else {
            p.scalars[1] = carry_s;                 // I2[n-1]
        }
    }
}

template <int PASS>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const __grid_constant__ PoissonArgs p) {
    const double i0 = PASS == 2 ? p.scalars[0] : 0.0;
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int64_t base = t * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
        double v[kScanItems];
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int64_t i = base + k;
            v[k] = i < p.n ? (PASS == 1 ? a_of(p, i) : g_of(p, i, i0)) : 0.0;
            acc += v[k];
        }
        double total;
        double run = p.tile_sums[t] + block_exclusive_scan(acc, total);
        double* dst = PASS == 1 ? p.I1 : p.out;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int64_t i = base + k;
            run += v[k];
            if (i < p.n) dst[i] = run;
            if (PASS == 2 && i == p.n - 1) p.scalars[1] = run;     // I2[n-1] as stored: out[n-1] becomes exactly 0
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) poisson_shift_kernel(double* out, const double* scalars, int64_t n) {
    const double last = scalars[1];
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kScanThreads;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kScanThreads + threadIdx.x; i < n; i += gstride)
        out[i] = out[i] - last;
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_poisson_scratch_doubles(int64_t n) {
    return n + (n + kScanTile - 1) / kScanTile + 8;      // I1 + tile sums + scalars
}

extern "C" int tp_poisson_radial_solve_f64(const double* sources, const double* r, double mean_dr, double* out,
                                           int64_t n, double* scratch, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (n < 3 || !sources || !r || !out || !scratch) return TP_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PoissonArgs p;
    p.r = r; p.s = sources; p.out = out; p.n = n; p.dr = mean_dr;
    p.ntiles = (n + kScanTile - 1) / kScanTile;
    p.I1 = scratch;
    p.tile_sums = scratch + n;
    p.scalars = scratch + n + p.ntiles;
    const int blocks = static_cast<int>(std::min<int64_t>(p.ntiles, static_cast<int64_t>(d.sm_count) * 8));
    scan_tile_sums_kernel<1><<<blocks, kScanThreads, 0, s>>>(p);
    scan_tile_offsets_kernel<1><<<1, kScanThreads, 0, s>>>(p);
    scan_apply_kernel<1><<<blocks, kScanThreads, 0, s>>>(p);
    scan_tile_sums_kernel<2><<<blocks, kScanThreads, 0, s>>>(p);
    scan_tile_offsets_kernel<2><<<1, kScanThreads, 0, s>>>(p);
    scan_apply_kernel<2><<<blocks, kScanThreads, 0, s>>>(p);
    poisson_shift_kernel<<<blocks, kScanThreads, 0, s>>>(out, p.scalars, n);
    count_launch(7);
    return static_cast<int>(cudaGetLastError());
}
