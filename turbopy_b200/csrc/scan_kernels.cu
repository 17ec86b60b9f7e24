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
// 56 B/point of traffic (16 algorithmic).  A thread owns 8 CONSECUTIVE elements
// (two 256-bit loads), scans them serially in registers, and only the 256 thread
// totals go through shuffles: one shuffle chain and one __syncthreads per tile
// scan.  The tile sum of K3 is the last value of the SAME local scan K5 runs, so
// out[n-1] is exactly 0 like the reference's I2 - I2[-1].  The result does not
// depend on the launch geometry (fixed tile size, fixed association), so it is
// reproducible run to run; grids of at most kSmallTiles tiles run K1..K5 as one
// single-CTA launch with the same arithmetic.
// numpy's cumsum is strictly sequential; a parallel scan associates differently,
// so results agree to rounding (tolerance in tests/test_gpu_poisson.py).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace tp {

constexpr int kScanThreads = 256;
constexpr int kScanRows = 8;                               // consecutive elements per thread
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanTile = kScanThreads * kScanRows;        // 2048
constexpr int kSmallTiles = 8;                             // <= 16384 points: one launch, one CTA

struct PoissonArgs {
    const double* r; const double* s; double* out;
    double* tile1;              // ntiles doubles: sums, then exclusive offsets of scan 1
    double* tile2;              // the same for scan 2
    double* scalars;            // [0] = i0, [1] = total of scan 2 (= I2[n-1])
    int64_t n; int64_t ntiles; double dr;
    bool vec;                   // r, s, out 32-byte aligned: 256-bit accesses
    bool lin;                   // r == nullptr: node coordinates rebuilt from the linspace parameters below
    LinGrid lg;
};

// Inclusive scan of the tile's 2048 values held as v[k] = element (thread*8 + k).
// Returns the tile total; v[] becomes the inclusive prefix inside the tile.
// `phase` alternates between consecutive calls so that one barrier per scan is enough.
__device__ __forceinline__ double tile_scan(double (&v)[kScanRows], int& phase) {
    __shared__ double warp_tot[2][kScanWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 1; k < kScanRows; ++k) v[k] = v[k - 1] + v[k];
    double inc = v[kScanRows - 1];                           // thread total -> inclusive over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[phase][warp] = inc;
    __syncthreads();
    double wbase = 0.0, total = 0.0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) {
        if (w == warp) wbase = total;
        total += warp_tot[phase][w];
    }
    phase ^= 1;
    const double base = wbase + (inc - v[kScanRows - 1]);    // everything before this thread's first element
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) v[k] = base + v[k];
    return total;
}

__device__ __forceinline__ int64_t elem_index(int64_t tile, int k) {
    return tile * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanRows + k;
}

__device__ __forceinline__ double r_at(const PoissonArgs& p, int64_t i) { return p.lin ? lin_r(p.lg, i) : p.r[i]; }
__device__ __forceinline__ double a_at(const PoissonArgs& p, int64_t i) { return (r_at(p, i) * p.s[i]) * p.dr; }

__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// the thread's 8 elements of `src` (fill beyond n)
__device__ __forceinline__ void load8(const PoissonArgs& p, const double* src, int64_t t, double fill, double (&v)[kScanRows]) {
    const int64_t i0 = elem_index(t, 0);
    if (p.vec && i0 + kScanRows <= p.n) {
        ld4(src + i0, v[0], v[1], v[2], v[3]);
        ld4(src + i0 + 4, v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = i0 + k < p.n ? src[i0 + k] : fill;
    }
}

// a[] of the tile (0 beyond n)
__device__ __forceinline__ void load_a(const PoissonArgs& p, int64_t t, double (&v)[kScanRows], double (&rr)[kScanRows]) {
    if (p.lin) {
        const int64_t i0 = elem_index(t, 0);
        const double di0 = static_cast<double>(i0);
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) rr[k] = i0 + k < p.n ? lin_r_at(p.lg, di0, k, i0 + k) : 1.0;
    } else {
        load8(p, p.r, t, 1.0, rr);
    }
    load8(p, p.s, t, 0.0, v);
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) v[k] = (rr[k] * v[k]) * p.dr;
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

__device__ __forceinline__ void sums1_body(const PoissonArgs& p, int& phase) {
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        double v[kScanRows], rr[kScanRows];
        load_a(p, t, v, rr);
        const double total = tile_scan(v, phase);
        if (threadIdx.x == 0) p.tile1[t] = total;
    }
}

// one CTA: tile sums -> exclusive offsets, in place
__device__ __forceinline__ void scan_tile_sums(double* sums, int64_t ntiles, int& phase) {
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += kScanTile) {
        double v[kScanRows], in[kScanRows];
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            const int64_t i = base + threadIdx.x * kScanRows + k;
            in[k] = i < ntiles ? sums[i] : 0.0;
            v[k] = in[k];
        }
        const double total = tile_scan(v, phase);
        const double carry = carry_s;
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            const int64_t i = base + threadIdx.x * kScanRows + k;
            if (i < ntiles) sums[i] = carry + (v[k] - in[k]);            // exclusive
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__device__ __forceinline__ void offsets1_body(const PoissonArgs& p, int& phase) {
    scan_tile_sums(p.tile1, p.ntiles, phase);
    if (threadIdx.x == 0) {
        // i0 = 2*g[1] - g[2] with I1[1] = a0 + a1, I1[2] = I1[1] + a2 in numpy's sequential order (:50-53)
        const double I1_1 = a_at(p, 0) + a_at(p, 1);
        const double I1_2 = I1_1 + a_at(p, 2);
        p.scalars[0] = 2.0 * ((I1_1 * p.dr) / r_at(p, 1)) - (I1_2 * p.dr) / r_at(p, 2);
    }
}

__device__ __forceinline__ void sums2_body(const PoissonArgs& p, int& phase) {
    const double i0 = p.scalars[0];
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        double v[kScanRows], rr[kScanRows];
        load_a(p, t, v, rr);
        tile_scan(v, phase);
        to_g(p, t, p.tile1[t], i0, rr, v);
        const double total = tile_scan(v, phase);        // the same scan K5 runs: its last value IS the tile sum
        if (threadIdx.x == 0) p.tile2[t] = total;
    }
}

__device__ __forceinline__ void offsets2_body(const PoissonArgs& p, int& phase) {
    __shared__ double last_sum;
    if (threadIdx.x == 0) last_sum = p.tile2[p.ntiles - 1];
    __syncthreads();
    scan_tile_sums(p.tile2, p.ntiles, phase);
    // I2[n-1] exactly as K5 forms it for the last element: off2[T-1] + (local inclusive value = tile sum)
    if (threadIdx.x == 0) p.scalars[1] = p.tile2[p.ntiles - 1] + last_sum;
}

__device__ __forceinline__ void apply_body(const PoissonArgs& p, int& phase) {
    const double i0 = p.scalars[0], last = p.scalars[1];
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        double v[kScanRows], rr[kScanRows];
        load_a(p, t, v, rr);
        tile_scan(v, phase);
        to_g(p, t, p.tile1[t], i0, rr, v);
        tile_scan(v, phase);
        const double off2 = p.tile2[t];
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = (off2 + v[k]) - last;
        const int64_t e0 = elem_index(t, 0);
        if (p.vec && e0 + kScanRows <= p.n) {
            st4(p.out + e0, v[0], v[1], v[2], v[3]);
            st4(p.out + e0 + 4, v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int k = 0; k < kScanRows; ++k)
                if (e0 + k < p.n) p.out[e0 + k] = v[k];
        }
    }
}

__global__ void __launch_bounds__(kScanThreads) poisson_sums1_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    sums1_body(p, phase);
}
__global__ void __launch_bounds__(kScanThreads) poisson_offsets1_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    offsets1_body(p, phase);
}
__global__ void __launch_bounds__(kScanThreads) poisson_sums2_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    sums2_body(p, phase);
}
__global__ void __launch_bounds__(kScanThreads) poisson_offsets2_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    offsets2_body(p, phase);
}
__global__ void __launch_bounds__(kScanThreads) poisson_apply_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    apply_body(p, phase);
}

// K1..K5 back to back in ONE CTA (small grids: the solve is launch-latency bound).  __syncthreads orders the
// CTA's own global writes (tile sums, scalars) before its later reads; arithmetic identical to the 5-launch path.
__global__ void __launch_bounds__(kScanThreads) poisson_small_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    sums1_body(p, phase);
    __syncthreads();
    offsets1_body(p, phase);
    __syncthreads();
    sums2_body(p, phase);
    __syncthreads();
    offsets2_body(p, phase);
    __syncthreads();
    apply_body(p, phase);
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_poisson_scratch_doubles(int64_t n) {
    return 2 * ((n + kScanTile - 1) / kScanTile) + 8;      // two arrays of tile sums + scalars
}

static int poisson_solve(const double* sources, const double* r, const tp_linspace* grid, double mean_dr, double* out,
                         int64_t n, double* scratch, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (n < 3 || !sources || (!r && !grid) || !out || !scratch) return TP_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PoissonArgs p;
    p.r = r; p.s = sources; p.out = out; p.n = n; p.dr = mean_dr;
    p.lin = r == nullptr;
    p.lg = LinGrid{0.0, 0.0, 0.0, 0};
    if (p.lin) {
        if (grid->n != n) return TP_E_ARG;
        p.lg = LinGrid{grid->start, grid->span, grid->step, grid->n};
    }
    p.ntiles = (n + kScanTile - 1) / kScanTile;
    p.tile1 = scratch;
    p.tile2 = scratch + p.ntiles;
    p.scalars = scratch + 2 * p.ntiles;
    p.vec = ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(sources) | reinterpret_cast<uintptr_t>(out)) & 31) == 0;
    if (p.ntiles <= kSmallTiles) {
        poisson_small_kernel<<<1, kScanThreads, 0, s>>>(p);
        count_launch(1);
        return static_cast<int>(cudaGetLastError());
    }
    // one resident wave per kernel (the result does not depend on the grid: tiles are fixed)
    auto wave = [&](auto kernel) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kScanThreads, 0) != cudaSuccess || v < 1) v = 3;
        if (const char* e = std::getenv("TURBOPY_B200_SCAN_BLOCKS")) v = std::max(1, std::atoi(e));   // tuning knob
        return static_cast<int>(std::min<int64_t>(p.ntiles, static_cast<int64_t>(d.sm_count) * v));
    };
    static int b1 = 0, b2 = 0, b3 = 0;
    static int64_t planned_tiles = -1;
    if (planned_tiles != p.ntiles) {
        b1 = wave(poisson_sums1_kernel); b2 = wave(poisson_sums2_kernel); b3 = wave(poisson_apply_kernel);
        planned_tiles = p.ntiles;
    }
    poisson_sums1_kernel<<<b1, kScanThreads, 0, s>>>(p);
    poisson_offsets1_kernel<<<1, kScanThreads, 0, s>>>(p);
    poisson_sums2_kernel<<<b2, kScanThreads, 0, s>>>(p);
    poisson_offsets2_kernel<<<1, kScanThreads, 0, s>>>(p);
    poisson_apply_kernel<<<b3, kScanThreads, 0, s>>>(p);
    count_launch(5);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_poisson_radial_solve_f64(const double* sources, const double* r, double mean_dr, double* out,
                                           int64_t n, double* scratch, tp_stream_t stream) {
    if (!r) return TP_E_ARG;
    return poisson_solve(sources, r, nullptr, mean_dr, out, n, scratch, stream);
}

extern "C" int tp_poisson_radial_solve_lin_f64(const double* sources, const tp_linspace* grid, double mean_dr, double* out,
                                               int64_t n, double* scratch, tp_stream_t stream) {
    if (!grid) return TP_E_ARG;
    return poisson_solve(sources, nullptr, grid, mean_dr, out, n, scratch, stream);
}
