// PoissonSolver1DRadial.solve (turbopy/computetools.py:35-59) as
// two chained fp64 prefix scans.  Line-by-line spec: oracle/poisson.py.
//
//   a[i]  = (r[i]*s[i])*dr                      I1 = cumsum(a)
//   g[i]  = (I1[i]*dr)/r[i],  g[0] := i0 = 2*g[1] - g[2],  g -= i0
//   I2    = cumsum(g),        out = I2 - I2[n-1]
//
// Reduce-then-scan over 2048-element tiles, no CTA ever waits on another one.  Round 2: TWO passes
// over the inputs instead of three.  Inside tile t, I1[i] = c_t + l_i (c_t = everything before the tile, l_i = the
// tile's own running sum), so the tile sum of g that the second scan needs is, up to rounding,
//   G_t = c_t * B_t + C_t - cnt_t * i0      with  B_t = sum dr/r[i],  C_t = sum l_i * (dr/r[i])
// and B_t, C_t do not depend on c_t: the first pass can form them next to the sums of a.
//   K1  per tile: a -> local scan -> A_t = sum a, B_t, C_t                        (reads r, s; nothing stored)
//   K2  one CTA: c_t = exclusive scan of A_t; i0 from the first three elements; G_t; off2_t = exclusive scan of
//       G_t; then the LAST tile through K3's own code: last = I2[n-1] exactly as K3 forms it
//   K3  per tile: a -> I1 = c_t + local scan -> g -> I2 = off2_t + local scan -> out = I2 - last   (reads r, s; writes out)
// 40 B/point of traffic reading r, 24 B/point on a linspace grid (16 algorithmic); three passes cost 56 / 32.
// The offsets off2_t are sums of the G_t, not of the rounded g[i] K3 adds up: the two agree to a few ulps per
// tile, the same size as the rounding any association of the running sum carries (tolerance in
// tests/test_gpu_poisson.py), and out[n-1] is exactly 0 like the reference's I2 - I2[-1] because K2 computes
// `last` with the very instructions K3 uses for element n-1.
// A thread owns 8 CONSECUTIVE elements (two 256-bit loads), scans them serially in registers, and only the 256
// thread totals go through shuffles: one shuffle chain and one barrier per tile scan.  The result does not depend
// on the launch geometry (fixed tile size, fixed association), so it is reproducible run to run; grids of at most
// kSmallTiles tiles run K1..K3 as one single-CTA launch with the same arithmetic.
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
constexpr int kOffThreads = 512;                           // K2: one CTA, 4096 tile sums per iteration (1024 threads spill)
constexpr int kOffWarps = kOffThreads / 32;

struct PoissonArgs {
    const double* r; const double* s; double* out;
    double* tile1;              // ntiles doubles: A_t = sums of a, then c_t = exclusive offsets of scan 1
    double* tileB;              // B_t = sum of dr/r[i] over the tile's elements (i >= 1)
    double* tileC;              // C_t = sum of l_i * (dr/r[i])
    double* tile2;              // off2_t = exclusive offsets of scan 2
    double* scalars;            // [0] = i0, [1] = I2[n-1]
    int64_t n; int64_t ntiles; double dr;
    bool vec;                   // r, s, out 32-byte aligned: 256-bit accesses
    bool tvec;                  // the per-tile arrays are 32-byte aligned (K2's 256-bit accesses)
    bool interior;              // interior tiles take the test-free path (A-B switch TURBOPY_B200_SCAN_INTERIOR=0)
    int pf_dist;                // K1 / K3: L2 bulk prefetch of the tile this many iterations ahead (0 = off)
    bool lin;                   // r == nullptr: node coordinates rebuilt from the linspace parameters below
    LinGrid lg;
};

// barrier over the first WARPS warps of the CTA (id 0 over the whole CTA is __syncthreads)
template <int WARPS, int BAR>
__device__ __forceinline__ void scan_barrier() {
    asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(WARPS * 32) : "memory");
}

// Inclusive scan of WARPS*256 values held as v[k] = element (thread*8 + k).
// Returns the total; v[] becomes the inclusive prefix inside the tile.
// `phase` alternates between consecutive calls so that one barrier per scan is enough.
template <int WARPS, int BAR>
__device__ __forceinline__ double tile_scan(double (&v)[kScanRows], int& phase) {
    __shared__ double warp_tot[2][WARPS];
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
    scan_barrier<WARPS, BAR>();
    double wbase = 0.0, total = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        if (w == warp) wbase = total;
        total += warp_tot[phase][w];
    }
    phase ^= 1;
    const double base = wbase + (inc - v[kScanRows - 1]);    // everything before this thread's first element
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) v[k] = base + v[k];
    return total;
}

// Sums of two per-thread values over the 256-thread tile CTA, fixed association (xor tree inside the warp, then the
// eight warps in order).  One barrier; the result is valid in thread 0 only.
template <int BAR>
__device__ __forceinline__ void tile_sum2(double& x, double& y, int& phase) {
    __shared__ double part[2][kScanWarps][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, o);
        y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if (lane == 0) { part[phase][warp][0] = x; part[phase][warp][1] = y; }
    scan_barrier<kScanWarps, BAR>();
    if (threadIdx.x == 0) {
        x = part[phase][0][0]; y = part[phase][0][1];
#pragma unroll
        for (int w = 1; w < kScanWarps; ++w) { x += part[phase][w][0]; y += part[phase][w][1]; }
    }
    phase ^= 1;
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

// IN (interior) tiles lie strictly inside (0, n-1): every element exists and neither element 0 (g[0] := i0) nor
// element n-1 (linspace's exact end point) is among them, so the per-element index tests -- 64-bit compares and
// selects, a third of the instructions of the generic tile -- are dropped.  Same arithmetic, same bits; the tile
// kind is CTA-uniform.  (ncu, 2^26+1 points on a linspace grid: 68 / 77 thread instructions per point in K1 / K3
// with issue slots 60-66 % busy -- the passes are bound by instruction issue and dependent latency there, not by HBM.)
__device__ __forceinline__ bool interior_tile(const PoissonArgs& p, int64_t t) {
    return p.interior && t > 0 && (t + 1) * kScanTile < p.n;
}

// the thread's 8 elements of `src` (fill beyond n)
template <bool IN>
__device__ __forceinline__ void load8(const PoissonArgs& p, const double* src, int64_t t, double fill, double (&v)[kScanRows]) {
    const int64_t i0 = elem_index(t, 0);
    if (p.vec && (IN || i0 + kScanRows <= p.n)) {
        ld4(src + i0, v[0], v[1], v[2], v[3]);
        ld4(src + i0 + 4, v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = (IN || i0 + k < p.n) ? src[i0 + k] : fill;
    }
}

// K2's accesses to the per-tile arrays: 8 consecutive entries per thread, plain (coherent) 256-bit accesses --
// one scalar load per entry made the single CTA issue 12 k sector requests per 4096 tiles (9 us per iteration)
__device__ __forceinline__ void ld4c(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void tiles_load8(const PoissonArgs& p, const double* src, int64_t t0, double (&v)[kScanRows]) {
    if (p.tvec && t0 + kScanRows <= p.ntiles) {
        ld4c(src + t0, v[0], v[1], v[2], v[3]);
        ld4c(src + t0 + 4, v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = t0 + k < p.ntiles ? src[t0 + k] : 0.0;
    }
}
__device__ __forceinline__ void tiles_store8(const PoissonArgs& p, double* dst, int64_t t0, const double (&v)[kScanRows]) {
    if (p.tvec && t0 + kScanRows <= p.ntiles) {
        st4(dst + t0, v[0], v[1], v[2], v[3]);
        st4(dst + t0 + 4, v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanRows; ++k)
            if (t0 + k < p.ntiles) dst[t0 + k] = v[k];
    }
}

// K1 / K3: the tile this CTA reads d iterations from now goes to L2 with one bulk prefetch per array
// (TURBOPY_B200_SCAN_PREFETCH=d).  Measured at 256 Mi points on the final kernels: on a linspace grid, where the
// passes wait on their own loads (long_scoreboard on top), d = 0 / 1 / 2: 1.65 / 1.60 / 1.62 ms; when r is read the
// four CTAs per SM already keep HBM at 89 % and the prefetches only disturb it: 1.85 / 1.91 / 2.14 ms (d = 3: 2.58).
// Default: d = 1 on linspace grids, off otherwise.
__device__ __forceinline__ void prefetch_tile_inputs(const PoissonArgs& p, int64_t t) {
    if (threadIdx.x != 0 || p.pf_dist == 0 || !p.vec) return;
    const int64_t tn = t + static_cast<int64_t>(p.pf_dist) * gridDim.x;
    if (tn >= p.ntiles) return;
    const int64_t first = tn * kScanTile;
    const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(kScanTile), p.n - first) * sizeof(double)) & ~15u;
    if (bytes == 0u) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.s + first), "r"(bytes) : "memory");
    if (!p.lin) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.r + first), "r"(bytes) : "memory");
}

// a[] of the tile (0 beyond n)
template <bool IN>
__device__ __forceinline__ void load_a(const PoissonArgs& p, int64_t t, double (&v)[kScanRows], double (&rr)[kScanRows]) {
    if (p.lin) {
        const int64_t i0 = elem_index(t, 0);
        const double di0 = static_cast<double>(i0);
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            if (IN) rr[k] = p.lg.start + p.lg.span * ((di0 + static_cast<double>(k)) * p.lg.step);   // lin_r_at, i != n-1
            else rr[k] = i0 + k < p.n ? lin_r_at(p.lg, di0, k, i0 + k) : 1.0;
        }
    } else {
        load8<IN>(p, p.r, t, 1.0, rr);
    }
    load8<IN>(p, p.s, t, 0.0, v);
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) v[k] = (rr[k] * v[k]) * p.dr;
}

// v: inclusive local scan of a (tile_scan output) -> g of the tile (0 beyond n)
template <bool IN>
__device__ __forceinline__ void to_g(const PoissonArgs& p, int64_t t, double off1, double i0, const double (&rr)[kScanRows],
                                     double (&v)[kScanRows]) {
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) {
        const double I1 = off1 + v[k];
        double g = ((I1 * p.dr) / rr[k]) - i0;
        if (!IN) {
            const int64_t i = elem_index(t, k);
            if (i == 0) g = i0 - i0;                   // g[0] := i0, then g - i0
            g = i < p.n ? g : 0.0;
        }
        v[k] = g;
    }
}

// number of elements of tile t that carry a "- i0" (every element below n except element 0)
__device__ __forceinline__ double tile_count(const PoissonArgs& p, int64_t t) {
    const int64_t lo = t * kScanTile, hi = min(p.n, lo + kScanTile);
    return static_cast<double>(hi - lo - (t == 0 ? 1 : 0));
}

// K1: A_t, B_t, C_t of every tile
template <bool IN, int BAR>
__device__ __forceinline__ void sums_tile(const PoissonArgs& p, int64_t t, int& phase, int& phase2) {
    double v[kScanRows], rr[kScanRows];
    prefetch_tile_inputs(p, t);
    load_a<IN>(p, t, v, rr);
    const double total = tile_scan<kScanWarps, BAR>(v, phase);           // v[k] = l_i
    double bs = 0.0, cs = 0.0;
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) {
        double w = p.dr / rr[k];
        if (!IN) {
            const int64_t i = elem_index(t, k);
            w = (i >= 1 && i < p.n) ? w : 0.0;                          // element 0 (r may be 0) is g[0] := i0
        }
        bs += w;
        cs += v[k] * w;
    }
    tile_sum2<BAR>(bs, cs, phase2);
    if (threadIdx.x == 0) { p.tile1[t] = total; p.tileB[t] = bs; p.tileC[t] = cs; }
}

template <int BAR>
__device__ __forceinline__ void sums_body(const PoissonArgs& p, int& phase, int& phase2) {
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        if (interior_tile(p, t)) sums_tile<true, BAR>(p, t, phase, phase2);
        else sums_tile<false, BAR>(p, t, phase, phase2);
    }
}

// I2[n-1] with the instructions K3 runs for that element: the first 256 threads of the calling CTA walk the last
// tile like apply_body does and the owner of element n-1 stores off2 + (its inclusive local value).
template <int BAR>
__device__ __forceinline__ void last_body(const PoissonArgs& p, double i0, int& phase) {
    const int64_t t = p.ntiles - 1;
    double v[kScanRows], rr[kScanRows];
    load_a<false>(p, t, v, rr);
    tile_scan<kScanWarps, BAR>(v, phase);
    to_g<false>(p, t, p.tile1[t], i0, rr, v);
    tile_scan<kScanWarps, BAR>(v, phase);
    const double off2 = p.tile2[t];
#pragma unroll
    for (int k = 0; k < kScanRows; ++k)
        if (elem_index(t, k) == p.n - 1) p.scalars[1] = off2 + v[k];
}

// K2 (one CTA of WARPS warps): A_t -> c_t in place, G_t -> off2_t, i0, I2[n-1]
template <int WARPS>
__device__ __forceinline__ void offsets_body(const PoissonArgs& p, int& phase) {
    __shared__ double carry_s[2], i0_s;
    constexpr int kChunk = WARPS * 32 * kScanRows;
    if (threadIdx.x == 0) {
        // i0 = 2*g[1] - g[2] with I1[1] = a0 + a1, I1[2] = I1[1] + a2 in numpy's sequential order (:50-53)
        const double I1_1 = a_at(p, 0) + a_at(p, 1);
        const double I1_2 = I1_1 + a_at(p, 2);
        const double i0 = 2.0 * ((I1_1 * p.dr) / r_at(p, 1)) - (I1_2 * p.dr) / r_at(p, 2);
        p.scalars[0] = i0;
        i0_s = i0; carry_s[0] = 0.0; carry_s[1] = 0.0;
    }
    __syncthreads();
    const double i0 = i0_s;
    for (int64_t base = 0; base < p.ntiles; base += kChunk) {
        double v[kScanRows], in[kScanRows], tb[kScanRows], tc[kScanRows];
        const int64_t t0 = base + static_cast<int64_t>(threadIdx.x) * kScanRows;
        tiles_load8(p, p.tile1, t0, in);                                  // all three loads fly before the first scan
        tiles_load8(p, p.tileB, t0, tb);
        tiles_load8(p, p.tileC, t0, tc);
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = in[k];
        const double total1 = tile_scan<WARPS, 0>(v, phase);
        const double carry1 = carry_s[0], carry2 = carry_s[1];
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            const int64_t t = t0 + k;
            const double c = carry1 + (v[k] - in[k]);                     // exclusive: everything before tile t
            in[k] = c;
            const double g = t < p.ntiles ? (c * tb[k] + tc[k]) - tile_count(p, t) * i0 : 0.0;
            tb[k] = g;
            v[k] = g;
        }
        tiles_store8(p, p.tile1, t0, in);
        const double total2 = tile_scan<WARPS, 0>(v, phase);
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = carry2 + (v[k] - tb[k]);
        tiles_store8(p, p.tile2, t0, v);
        __syncthreads();
        if (threadIdx.x == 0) { carry_s[0] = carry1 + total1; carry_s[1] = carry2 + total2; }
        __syncthreads();
    }
    // the CTA's own global writes (c_t, off2_t of the last tile) are ordered before these reads by the barrier above
    if (threadIdx.x < kScanThreads) last_body<1>(p, i0, phase);
}

// K3
template <bool IN, int BAR>
__device__ __forceinline__ void apply_tile(const PoissonArgs& p, int64_t t, double i0, double last, int& phase) {
    double v[kScanRows], rr[kScanRows];
    prefetch_tile_inputs(p, t);
    load_a<IN>(p, t, v, rr);
    tile_scan<kScanWarps, BAR>(v, phase);
    to_g<IN>(p, t, p.tile1[t], i0, rr, v);
    tile_scan<kScanWarps, BAR>(v, phase);
    const double off2 = p.tile2[t];
#pragma unroll
    for (int k = 0; k < kScanRows; ++k) v[k] = (off2 + v[k]) - last;
    const int64_t e0 = elem_index(t, 0);
    if (p.vec && (IN || e0 + kScanRows <= p.n)) {
        st4(p.out + e0, v[0], v[1], v[2], v[3]);
        st4(p.out + e0 + 4, v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanRows; ++k)
            if (IN || e0 + k < p.n) p.out[e0 + k] = v[k];
    }
}

template <int BAR>
__device__ __forceinline__ void apply_body(const PoissonArgs& p, int& phase) {
    const double i0 = p.scalars[0], last = p.scalars[1];
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        if (interior_tile(p, t)) apply_tile<true, BAR>(p, t, i0, last, phase);
        else apply_tile<false, BAR>(p, t, i0, last, phase);
    }
}

__global__ void __launch_bounds__(kScanThreads) poisson_sums_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0, phase2 = 0;
    sums_body<0>(p, phase, phase2);
}
__global__ void __launch_bounds__(kOffThreads) poisson_offsets_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    offsets_body<kOffWarps>(p, phase);
}
__global__ void __launch_bounds__(kScanThreads, 4) poisson_apply_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0;
    apply_body<0>(p, phase);
}

// K1..K3 back to back in ONE CTA (small grids: the solve is launch-latency bound).  __syncthreads orders the
// CTA's own global writes (tile sums, scalars) before its later reads; arithmetic identical to the 3-launch path
// (the exclusive offsets of at most kSmallTiles tiles come out of the first warp's shuffles in both).
__global__ void __launch_bounds__(kScanThreads) poisson_small_kernel(const __grid_constant__ PoissonArgs p) {
    int phase = 0, phase2 = 0;
    sums_body<0>(p, phase, phase2);
    __syncthreads();
    offsets_body<kScanWarps>(p, phase);
    __syncthreads();
    apply_body<0>(p, phase);
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_poisson_scratch_doubles(int64_t n) {
    const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
    return 4 * ((ntiles + 3) & ~int64_t(3)) + 8;           // A/c, B, C, off2 per tile (each padded to 32 bytes) + scalars
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
    const int64_t ts = (p.ntiles + 3) & ~int64_t(3);
    p.tile1 = scratch;
    p.tileB = scratch + ts;
    p.tileC = scratch + 2 * ts;
    p.tile2 = scratch + 3 * ts;
    p.scalars = scratch + 4 * ts;
    p.tvec = (reinterpret_cast<uintptr_t>(scratch) & 31) == 0;
    static const int pf_dist = [] { const char* e = std::getenv("TURBOPY_B200_SCAN_PREFETCH"); return e ? std::max(0, std::atoi(e)) : -1; }();
    p.pf_dist = pf_dist >= 0 ? pf_dist : (p.lin ? 1 : 0);      // default: one tile ahead on linspace grids only (measured, see prefetch_tile_inputs)
    const char* in_env = std::getenv("TURBOPY_B200_SCAN_INTERIOR");       // read per call: tests compare both paths
    p.interior = !(in_env && in_env[0] == '0');
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
    static int b1 = 0, b3 = 0;
    static int64_t planned_tiles = -1;
    if (planned_tiles != p.ntiles) {
        b1 = wave(poisson_sums_kernel); b3 = wave(poisson_apply_kernel);
        planned_tiles = p.ntiles;
    }
    poisson_sums_kernel<<<b1, kScanThreads, 0, s>>>(p);
    poisson_offsets_kernel<<<1, kOffThreads, 0, s>>>(p);
    poisson_apply_kernel<<<b3, kScanThreads, 0, s>>>(p);
    count_launch(3);
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
