// PoissonSolver1DRadial.solve (turbopy/computetools.py:35-59) as
// two chained fp64 prefix scans.  Line-by-line spec: oracle/poisson.py.
//
//   a[i]  = (r[i]*s[i])*dr                      I1 = cumsum(a)
//   g[i]  = (I1[i]*dr)/r[i],  g[0] := i0 = 2*g[1] - g[2],  g -= i0
//   I2    = cumsum(g),        out = I2 - I2[n-1]
//
// Reduce-then-scan over 2048-element tiles, no CTA ever waits on another one:
//   K1  per-tile sums of a                                   (reads r, s)
//   K2  one CTA: exclusive scan of the tile sums -> off1[t]; i0 from the first three elements
//   K3  per tile: a -> I1 = off1 + local scan -> g -> per-tile sums of g   (reads r, s; nothing stored)
//   K4  one CTA: exclusive scan -> off2[t]; total = off2[T-1] + G[T-1]
//   K5  per tile: a -> I1 -> g -> I2 = off2 + local scan -> out = I2 - total   (reads r, s; writes out)
// I1 is recomputed instead of stored and the final shift is folded into K5:
// 56 B/point of traffic (16 algorithmic), all of it coalesced -- a warp owns
// 8 rows of 32 consecutive elements and scans them with shuffles.  The tile sum
// of K3 is the last value of the SAME local scan K5 runs, so out[n-1] is
// exactly 0 like the reference's I2 - I2[-1].
// numpy's cumsum is strictly sequential; a parallel scan associates differently,
// so results agree to rounding (tolerance in tests/test_gpu_poisson.py).
#include "common.cuh"

#include <algorithm>

namespace tp {

constexpr int kScanThreads = 256;
constexpr int kScanRows = 8;                               // rows of 32 elements per warp
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanTile = kScanThreads * kScanRows;        // 2048

struct PoissonArgs {
    const double* r; const double* s; double* out;
    double* tile1;              // ntiles doubles: sums, then exclusive offsets of scan 1
    double* tile2;              // the same for scan 2
    double* scalars;            // [0] = i0, [1] = total of scan 2 (= I2[n-1])
    int64_t n; int64_t ntiles; double dr;
};

__device__ __forceinline__ double warp_inclusive(double v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Inclusive scan of the tile's 2048 values held as v[k] = element (warp*256 + k*32 + lane).
// Returns the tile total; v[] becomes the inclusive prefix inside the tile.
__device__ __forceinline__ double tile_scan(double (&v)[kScanRows]) {
    __shared__ double warp_tot[kScanWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double carry = 0.0;
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) {
        const double inc = warp_inclusive(v[k]);
        v[k] = carry + inc;
        carry = carry + __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) warp_tot[warp] = carry;
    __syncthreads();
    double wbase = 0.0;
#pragma unroll
    for (int w = 0; w < kScanWarps - 1; ++w)
        if (w < warp) wbase += warp_tot[w];
    double total = 0.0;
#pragma unroll
    for (int w = 0; w < kScanWarps - 1; ++w) total += warp_tot[w];
    total = total + warp_tot[kScanWarps - 1];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) v[k] = wbase + v[k];
    return total;
}

__device__ __forceinline__ int64_t elem_index(int64_t tile, int k) {
    return tile * kScanTile + (threadIdx.x >> 5) * (32 * kScanRows) + k * 32 + (threadIdx.x & 31);
}

__device__ __forceinline__ double a_at(const PoissonArgs& p, int64_t i) { return (p.r[i] * p.s[i]) * p.dr; }

// a[] of the tile (0 beyond n)
__device__ __forceinline__ void load_a(const PoissonArgs& p, int64_t t, double (&v)[kScanRows], double (&rr)[kScanRows]) {
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) {
        const int64_t i = elem_index(t, k);
        rr[k] = i < p.n ? p.r[i] : 1.0;
        v[k] = i < p.n ? (rr[k] * p.s[i]) * p.dr : 0.0;
    }
}

// v: inclusive local scan of a (tile_scan output) -> g of the tile (0 beyond n)
__device__ __forceinline__ void to_g(const PoissonArgs& p, int64_t t, double off1, double i0, const double (&rr)[kScanRows],
                                     double (&v)[kScanRows]) {
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) {
        const int64_t i = elem_index(t, k);
        const double I1 = off1 + v[k];
        double g = ((I1 * p.dr) / rr[k]) - i0;
        if (i == 0) g = i0 - i0;                       // g[0] := i0, then g - i0
        v[k] = i < p.n ? g : 0.0;
    }
}

__global__ void __launch_bounds__(kScanThreads) poisson_sums1_kernel(const __grid_constant__ PoissonArgs p) {
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        double v[kScanRows], rr[kScanRows];
        load_a(p, t, v, rr);
        const double total = tile_scan(v);
        if (threadIdx.x == 0) p.tile1[t] = total;
    }
}

// one CTA: tile sums -> exclusive offsets, in place; returns the grand total through *total_out
__device__ __forceinline__ void scan_tile_sums(double* sums, int64_t ntiles, double* total_out) {
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += kScanTile) {
        double v[kScanRows], in[kScanRows];
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            const int64_t i = base + (threadIdx.x >> 5) * (32 * kScanRows) + k * 32 + (threadIdx.x & 31);
            in[k] = i < ntiles ? sums[i] : 0.0;
            v[k] = in[k];
        }
        const double total = tile_scan(v);
        const double carry = carry_s;
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            const int64_t i = base + (threadIdx.x >> 5) * (32 * kScanRows) + k * 32 + (threadIdx.x & 31);
            if (i < ntiles) sums[i] = carry + (v[k] - in[k]);            // exclusive
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

__global__ void __launch_bounds__(kScanThreads) poisson_offsets1_kernel(const __grid_constant__ PoissonArgs p) {
    scan_tile_sums(p.tile1, p.ntiles, nullptr);
    if (threadIdx.x == 0) {
        // i0 = 2*g[1] - g[2] with I1[1] = a0 + a1, I1[2] = I1[1] + a2 in numpy's sequential order (:50-53)
        const double I1_1 = a_at(p, 0) + a_at(p, 1);
        const double I1_2 = I1_1 + a_at(p, 2);
        p.scalars[0] = 2.0 * ((I1_1 * p.dr) / p.r[1]) - (I1_2 * p.dr) / p.r[2];
    }
}

__global__ void __launch_bounds__(kScanThreads) poisson_sums2_kernel(const __grid_constant__ PoissonArgs p) {
    const double i0 = p.scalars[0];
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        double v[kScanRows], rr[kScanRows];
        load_a(p, t, v, rr);
        tile_scan(v);
        to_g(p, t, p.tile1[t], i0, rr, v);
        const double total = tile_scan(v);               // the same scan K5 runs: its last value IS the tile sum
        if (threadIdx.x == 0) p.tile2[t] = total;
    }
}

__global__ void __launch_bounds__(kScanThreads) poisson_offsets2_kernel(const __grid_constant__ PoissonArgs p) {
    __shared__ double last_sum;
    if (threadIdx.x == 0) last_sum = p.tile2[p.ntiles - 1];
    __syncthreads();
    scan_tile_sums(p.tile2, p.ntiles, nullptr);
    __syncthreads();
    // I2[n-1] exactly as K5 forms it for the last element: off2[T-1] + (local inclusive value = tile sum)
    if (threadIdx.x == 0) p.scalars[1] = p.tile2[p.ntiles - 1] + last_sum;
}

__global__ void __launch_bounds__(kScanThreads) poisson_apply_kernel(const __grid_constant__ PoissonArgs p) {
    const double i0 = p.scalars[0], last = p.scalars[1];
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        double v[kScanRows], rr[kScanRows];
        load_a(p, t, v, rr);
        tile_scan(v);
        to_g(p, t, p.tile1[t], i0, rr, v);
        tile_scan(v);
        const double off2 = p.tile2[t];
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            const int64_t i = elem_index(t, k);
            if (i < p.n) p.out[i] = (off2 + v[k]) - last;
        }
    }
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_poisson_scratch_doubles(int64_t n) {
    return 2 * ((n + kScanTile - 1) / kScanTile) + 8;      // two arrays of tile sums + scalars
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
    p.tile1 = scratch;
    p.tile2 = scratch + p.ntiles;
    p.scalars = scratch + 2 * p.ntiles;
    const int blocks = static_cast<int>(std::min<int64_t>(p.ntiles, static_cast<int64_t>(d.sm_count) * 8));
    poisson_sums1_kernel<<<blocks, kScanThreads, 0, s>>>(p);
    poisson_offsets1_kernel<<<1, kScanThreads, 0, s>>>(p);
    poisson_sums2_kernel<<<blocks, kScanThreads, 0, s>>>(p);
    poisson_offsets2_kernel<<<1, kScanThreads, 0, s>>>(p);
    poisson_apply_kernel<<<blocks, kScanThreads, 0, s>>>(p);
    count_launch(5);
    return static_cast<int>(cudaGetLastError());
}
