// Kernel 2 of the hot path: FiniteDifference 1-D stencils
// (turbopy/computetools.py:104-138 direct stencils, :189-223 del2_radial, and
// the banded dia_matrix operators of :140-187,225-381 applied as `D @ f`).
// Line-by-line spec: oracle/fd.py.  HBM-bound streaming kernels:
//   centered / constant-diagonal operators  16 B/point  (8 read + 8 written)
//   upwind_left                             24 B/point  (y, cell_widths read; d written)
//   del2_radial apply                       24 B/point  (f, r read; out written; coefficients rebuilt
//                                                        in registers instead of 3 stored diagonals)
//   stored-diagonal operators               (ndiag+2)*8 B/point
// Each thread produces FOUR consecutive outputs per iteration from a register
// window of the operand (two 128-bit loads + halo scalars that hit L1, two
// 128-bit stores): 64 B in flight per thread, 2048 threads per SM.
// Compiled with -fmad=false.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace tp {

constexpr int kFdThreads = 256;
constexpr int kQuad = 4;

// Persistent CTAs per SM, per kernel.  Measured on the B200 (tools/sweep_fd.py, 16 Mi and 256 Mi points, % of the
// HBM copy peak at 2 / 3 / 4 / 5 / 6 / 8 / 12 / 16 CTAs of 256 threads per SM): these streaming kernels peak around
// 1024-1280 threads per SM and lose 10-20 % on a ragged second wave (e.g. 6 or 8 CTAs when 5 are resident):
//   centered      67 / 84 / 95 / 98 / 77 / 92 / 91 / 92        -> 5
//   upwind        74 / 91 / 100 / 83 / 84 / 84 / 93 / 96       -> 4
//   del2_radial   77 / 96 / 99 / 81 / 90 / 103 / 105 / 105     -> 12
//   ddx, del2, BC 74 / 89 / 98 / 88 / 83 / 80 / 75 / 97        -> 4
//   radial_curl   52 / 71 / 82 / 93 / 67 / 82 / 84 / 84        -> 5
template <typename K>
static int fd_blocks(K kernel, int64_t n_quads, int per_sm) {
    (void)kernel;
    const DeviceProps& d = device_props();
    static const int forced = [] {
        const char* e = std::getenv("TURBOPY_B200_FD_BLOCKS");   // tuning knob
        return e ? std::max(1, std::atoi(e)) : 0;
    }();
    if (forced) per_sm = forced;
    const int64_t full = static_cast<int64_t>(d.sm_count) * per_sm;
    const int64_t need = (n_quads + kFdThreads - 1) / kFdThreads;
    return static_cast<int>(need < 1 ? 1 : (need < full ? need : full));
}

// w[k] = y[i - H + k] for k in [0, 4 + 2H); out-of-range entries are 0 (never used).
template <int H>
__device__ __forceinline__ void load_window(const double* __restrict__ y, int64_t i, int64_t n, bool vec,
                                            double (&w)[kQuad + 2 * H]) {
    if (vec && i >= H && i + kQuad + H <= n) {
        const double2 a = ld2(y + i), b = ld2(y + i + 2);
#pragma unroll
        for (int k = 0; k < H; ++k) { w[k] = y[i - H + k]; w[kQuad + H + k] = y[i + kQuad + k]; }
        w[H] = a.x; w[H + 1] = a.y; w[H + 2] = b.x; w[H + 3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < kQuad + 2 * H; ++k) {
            const int64_t j = i - H + k;
            w[k] = (j >= 0 && j < n) ? y[j] : 0.0;
        }
    }
}

__device__ __forceinline__ void store_quad(double* __restrict__ d, int64_t i, int64_t n, bool vec, const double (&o)[kQuad]) {
    if (vec && i + kQuad <= n) {
        st2(d + i, make_double2(o[0], o[1]));
        st2(d + i + 2, make_double2(o[2], o[3]));
    } else {
#pragma unroll
        for (int k = 0; k < kQuad; ++k)
            if (i + k < n) d[i + k] = o[k];
    }
}

#define TP_QUAD_LOOP(n)                                                                    \
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kFdThreads + threadIdx.x;      \
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kFdThreads;                  \
    const int64_t nquads = ((n) + kQuad - 1) / kQuad;                                      \
    for (int64_t qd = gtid; qd < nquads; qd += gstride)

// Quotient by the constant 2*dr: FAST policy with a per-value check, IEEE otherwise.
struct ConstDiv { double b, rb; int fast_ok; };

__device__ __forceinline__ double div_const(double a, const ConstDiv& c) {
    Guard g;
    const double q = quot<true>(a, const_recip<true>(c.b, c.rb), g);
    if (__builtin_expect(c.fast_ok && guard_ok(g) && finite3(q, q, q), 1)) return q;
    return a / c.b;
}

// d[i] = (y[i+1] - y[i-1]) / two_dr for 0 < i < n-1; d[0] = d[n-1] = 0.
// Five CTAs per SM are launched (table above), so five must be RESIDENT: the bound caps the kernel at 48 registers.
// Without it a later change to the Guard helpers took the kernel from 48 to 52 registers, four CTAs fitted, the fifth
// ran as a ragged second wave and the stencil fell from 98 % to 72 % of the copy peak (256 Mi points: 0.66 -> 0.90 ms;
// found in round 2's final kernel table, tools/micro/fd_centered_probe.py).
__global__ void __launch_bounds__(kFdThreads, 5) fd_centered_kernel(const double* __restrict__ y,
                                                                 double* __restrict__ d, int64_t n,
                                                                 ConstDiv two_dr, bool vec) {
    TP_QUAD_LOOP(n) {
        const int64_t i = qd * kQuad;
        double w[kQuad + 2], o[kQuad];
        load_window<1>(y, i, n, vec, w);
#pragma unroll
        for (int k = 0; k < kQuad; ++k) {
            const int64_t j = i + k;
            o[k] = (j > 0 && j < n - 1) ? div_const(w[k + 2] - w[k], two_dr) : 0.0;
        }
        store_quad(d, i, n, vec, o);
    }
}

// d[i] = (y[i] - y[i-1]) / w[i-1] for i >= 1; d[0] = 0.  LIN: w[i-1] = r[i] - r[i-1] rebuilt from the linspace
// parameters (Grid.cell_widths = r[1:] - r[:-1], turbopy/core.py:670) instead of being read.
template <bool LIN>
__global__ void __launch_bounds__(kFdThreads) fd_upwind_kernel(const double* __restrict__ y,
                                                               const double* __restrict__ cw, const LinGrid lg,
                                                               double* __restrict__ d, int64_t n, bool vec) {
    TP_QUAD_LOOP(n) {
        const int64_t i = qd * kQuad;
        double w[kQuad + 2], o[kQuad];
        load_window<1>(y, i, n, vec, w);
        const double di = LIN ? static_cast<double>(i) : 0.0;
        double rprev = (LIN && i > 0) ? lin_r_at(lg, di, -1, i - 1) : 0.0;
#pragma unroll
        for (int k = 0; k < kQuad; ++k) {
            const int64_t j = i + k;
            double width = 1.0;
            if (LIN) {
                const double rj = j < n ? lin_r_at(lg, di, k, j) : 0.0;
                width = rj - rprev;
                rprev = rj;
            } else if (j > 0 && j < n) {
                width = cw[j - 1];
            }
            o[k] = (j > 0 && j < n) ? (w[k + 1] - w[k]) / width : 0.0;
        }
        store_quad(d, i, n, vec, o);
    }
}

// out = del2_radial() @ f with the rows rebuilt from r:
//   below[i] = (-g1)/r[i] + g2, diag = -2*g2, above[i] = g1/r[i] + g2,
//   above[0] = 0 + 2*g2 (the r=0 boundary rows, computetools.py:206,217)
//   out[i] = ((0 + below[i]*f[i-1]) + diag*f[i]) + above[i]*f[i+1]
// (scipy's dia matvec accumulates one diagonal at a time into zeros).
__device__ __forceinline__ double del2_row(int64_t i, int64_t n, double fm, double fc, double fp, double ri,
                                           double g1, double g2, double diag, double above0) {
    double acc = 0.0;
    // g1/r[i] and (-g1)/r[i] = -(g1/r[i]) exactly: one quotient, FAST with a local check
    Guard g;
    const Recip<true> rr = make_recip<true>(ri, g);
    double qa = quot<true>(g1, rr, g);
    if (__builtin_expect(!(guard_ok(g) && finite3(qa, qa, qa) && ri > 0.0), 0)) qa = g1 / ri;
    if (i > 0) acc = acc + ((-qa) + g2) * fm;
    acc = acc + diag * fc;
    if (i < n - 1) {
        const double above = (i == 0) ? above0 : (qa + g2);
        acc = acc + above * fp;
    }
    return acc;
}

// AXPY: out = f + dtau * (D f), the explicit update `f + dtau * (D @ f)` of a diffusion-type field module in ONE
// pass (numpy evaluates D @ f, then dtau * that, then the sum: three rounded operations per point, reproduced here
// without contraction) -- 16 B per point instead of the 16 + 24 of a matvec followed by an axpy.
template <bool LIN, bool AXPY = false>
__global__ void __launch_bounds__(kFdThreads) fd_del2_radial_kernel(const double* __restrict__ f,
                                                                    const double* __restrict__ r, const LinGrid lg,
                                                                    double* __restrict__ out, int64_t n, double g1,
                                                                    double g2, bool vec, double dtau = 0.0) {
    const double diag = -2.0 * g2;
    const double above0 = 0.0 + 2.0 * g2;
    TP_QUAD_LOOP(n) {
        const int64_t i = qd * kQuad;
        double w[kQuad + 2], rr[kQuad], o[kQuad];
        load_window<1>(f, i, n, vec, w);
        if (LIN) {
            const double di = static_cast<double>(i);
#pragma unroll
            for (int k = 0; k < kQuad; ++k) rr[k] = i + k < n ? lin_r_at(lg, di, k, i + k) : 1.0;
        } else {
            load_window<0>(r, i, n, vec, rr);
        }
#pragma unroll
        for (int k = 0; k < kQuad; ++k) {
            o[k] = (i + k < n) ? del2_row(i + k, n, w[k], w[k + 1], w[k + 2], rr[k], g1, g2, diag, above0) : 0.0;
            if (AXPY) o[k] = w[k + 1] + dtau * o[k];
        }
        store_quad(out, i, n, vec, o);
    }
}

// w[base + off] for a warp-uniform runtime offset in [-2, 2] WITHOUT dynamic
// register indexing (which would push the window into local memory).
template <int N>
__device__ __forceinline__ double pick(const double (&w)[N], int base, int off) {
    // base is a compile-time constant after unrolling; the switch is warp-uniform
    switch (off) {
        case -2: return w[base - 2];
        case -1: return w[base - 1];
        case 0: return w[base];
        case 1: return w[base + 1];
        default: return w[base + 2];
    }
}

// Generic banded apply in scipy's dia-matvec order, stored diagonals.
struct DiaArgs {
    const double* data; const double* x; double* out; int64_t n; int ndiag; int off[5];
};

__global__ void __launch_bounds__(kFdThreads) fd_dia_kernel(const __grid_constant__ DiaArgs a, bool vec) {
    TP_QUAD_LOOP(a.n) {
        const int64_t i = qd * kQuad;
        double w[kQuad + 4], o[kQuad] = {0.0, 0.0, 0.0, 0.0};
        load_window<2>(a.x, i, a.n, vec, w);
        for (int k = 0; k < a.ndiag; ++k) {
            const double* dk = a.data + static_cast<int64_t>(k) * a.n;
#pragma unroll
            for (int q = 0; q < kQuad; ++q) {
                const int64_t j = i + q + a.off[k];
                if (i + q < a.n && j >= 0 && j < a.n) o[q] = o[q] + dk[j] * pick(w, q + 2, a.off[k]);
            }
        }
        store_quad(a.out, i, a.n, vec, o);
    }
}

// Banded apply for operators whose stored diagonals are CONSTANT apart from a
// few boundary entries (ddx, ddr, del2, BC_*: turbopy/computetools.py:140-155,
// 225-381): the coefficients travel as kernel arguments, nothing but x is read
// (16 B/point instead of (ndiag+2)*8).  Same summation order as scipy's
// dia_matvec; entry (k, j) means data[k][j].
constexpr int kMaxOverrides = 8;
struct DiaConstArgs {
    const double* x; double* out; int64_t n;
    int ndiag; int off[5]; double coef[5];
    int nov; int ov_k[kMaxOverrides]; int64_t ov_j[kMaxOverrides]; double ov_v[kMaxOverrides];
};

template <int Q>
__device__ __forceinline__ double dia_const_row(const DiaConstArgs& a, int64_t i, const double (&w)[kQuad + 4],
                                                bool edge) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (k >= a.ndiag) break;
        const int64_t j = i + a.off[k];
        if (j < 0 || j >= a.n) continue;
        double c = a.coef[k];
        if (edge)                                       // overrides only live next to the boundaries
            for (int o = 0; o < a.nov; ++o)
                if (a.ov_k[o] == k && a.ov_j[o] == j) c = a.ov_v[o];
        acc = acc + c * pick(w, Q + 2, a.off[k]);
    }
    return acc;
}

__global__ void __launch_bounds__(kFdThreads) fd_dia_const_kernel(const __grid_constant__ DiaConstArgs a, bool vec) {
    TP_QUAD_LOOP(a.n) {
        const int64_t i = qd * kQuad;
        double w[kQuad + 4], o[kQuad];
        load_window<2>(a.x, i, a.n, vec, w);
        const bool edge = (i < 8) || (i + kQuad + 8 > a.n);       // quad within reach of an override column
        o[0] = dia_const_row<0>(a, i, w, edge);
        o[1] = (i + 1 < a.n) ? dia_const_row<1>(a, i + 1, w, edge) : 0.0;
        o[2] = (i + 2 < a.n) ? dia_const_row<2>(a, i + 2, w, edge) : 0.0;
        o[3] = (i + 3 < a.n) ? dia_const_row<3>(a, i + 3, w, edge) : 0.0;
        store_quad(a.out, i, a.n, vec, o);
    }
}

// The same operator with the offset pattern known at compile time (the five
// patterns the reference's builders use): interior quads run unchecked with
// static window indices; quads near a boundary take the generic rows above.
template <int ND, int O0, int O1, int O2>
__global__ void __launch_bounds__(kFdThreads) fd_dia_const_t_kernel(const __grid_constant__ DiaConstArgs a, bool vec) {
    const double c0 = a.coef[0], c1 = a.coef[1], c2 = ND > 2 ? a.coef[2] : 0.0;
    constexpr int LO = (O0 < O1 ? O0 : O1) < (ND > 2 ? O2 : 0) ? (O0 < O1 ? O0 : O1) : (ND > 2 ? O2 : 0);
    constexpr int HI = (O0 > O1 ? O0 : O1) > (ND > 2 ? O2 : 0) ? (O0 > O1 ? O0 : O1) : (ND > 2 ? O2 : 0);
    constexpr int HL = LO < 0 ? -LO : 0, HR = HI > 0 ? HI : 0;      // halo actually touched by this pattern
    TP_QUAD_LOOP(a.n) {
        const int64_t i = qd * kQuad;
        double w[kQuad + 4], o[kQuad];
        const bool edge = (i < 8) || (i + kQuad + 8 > a.n);
        if (__builtin_expect(!edge, 1)) {
            if (vec) {
                const double2 p0 = ld2(a.x + i), p1 = ld2(a.x + i + 2);
                w[2] = p0.x; w[3] = p0.y; w[4] = p1.x; w[5] = p1.y;
            } else {
#pragma unroll
                for (int k = 0; k < kQuad; ++k) w[2 + k] = a.x[i + k];
            }
#pragma unroll
            for (int k = 1; k <= HL; ++k) w[2 - k] = a.x[i - k];
#pragma unroll
            for (int k = 0; k < HR; ++k) w[2 + kQuad + k] = a.x[i + kQuad + k];
#pragma unroll
            for (int q = 0; q < kQuad; ++q) {
                double acc = 0.0;                       // scipy: y starts from zeros
                acc = acc + c0 * w[q + 2 + O0];
                acc = acc + c1 * w[q + 2 + O1];
                if (ND > 2) acc = acc + c2 * w[q + 2 + O2];
                o[q] = acc;
            }
        } else {
            load_window<2>(a.x, i, a.n, vec, w);
            o[0] = dia_const_row<0>(a, i, w, true);
            o[1] = (i + 1 < a.n) ? dia_const_row<1>(a, i + 1, w, true) : 0.0;
            o[2] = (i + 2 < a.n) ? dia_const_row<2>(a, i + 2, w, true) : 0.0;
            o[3] = (i + 3 < a.n) ? dia_const_row<3>(a, i + 3, w, true) : 0.0;
        }
        store_quad(a.out, i, a.n, vec, o);
    }
}

template <int ND, int O0, int O1, int O2>
static bool launch_dia_const_t(const DiaConstArgs& a, bool vec, cudaStream_t s) {
    if (a.ndiag != ND || a.off[0] != O0 || a.off[1] != O1 || (ND > 2 && a.off[2] != O2)) return false;
    auto kern = fd_dia_const_t_kernel<ND, O0, O1, O2>;
    kern<<<quads(kern, a.n, 4), kFdThreads, 0, s>>>(a, vec);
    return true;
}

static ConstDiv host_recip(double b) {
    ConstDiv c;
    c.b = b;
    c.rb = 1.0 / b;
    c.fast_ok = const_divisor_ok(b) ? 1 : 0;
    return c;
}

template <typename K>
static int quads(K kernel, int64_t n, int per_sm) { return fd_blocks(kernel, (n + kQuad - 1) / kQuad, per_sm); }

}  // namespace tp

using namespace tp;

extern "C" int tp_fd_centered_f64(const double* y, double* d, int64_t n, double two_dr, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!y || !d))) return TP_E_ARG;
    if (n == 0) return TP_OK;
    const bool vec = aligned16(y) && aligned16(d);
    fd_centered_kernel<<<quads(fd_centered_kernel, n, 5), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(y, d, n, host_recip(two_dr), vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

static int fill_lin(const tp_linspace* g, int64_t n, LinGrid& lg) {
    if (!g || g->n != n) return TP_E_ARG;
    lg.start = g->start; lg.span = g->span; lg.step = g->step; lg.n = g->n;
    return TP_OK;
}

extern "C" int tp_fd_upwind_left_f64(const double* y, const double* widths, double* d, int64_t n,
                                     tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!y || !d)) || (n > 1 && !widths)) return TP_E_ARG;
    if (n == 0) return TP_OK;
    const bool vec = aligned16(y) && aligned16(d);
    fd_upwind_kernel<false><<<quads(fd_upwind_kernel<false>, n, 4), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        y, widths, LinGrid{}, d, n, vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_upwind_left_lin_f64(const double* y, const tp_linspace* grid, double* d, int64_t n, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!y || !d))) return TP_E_ARG;
    if (n == 0) return TP_OK;
    LinGrid lg;
    int rc = fill_lin(grid, n, lg);
    if (rc) return rc;
    const bool vec = aligned16(y) && aligned16(d);
    fd_upwind_kernel<true><<<quads(fd_upwind_kernel<true>, n, 4), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        y, nullptr, lg, d, n, vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_del2_radial_apply_f64(const double* f, const double* r, double* out, int64_t n, double g1,
                                           double g2, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!f || !r || !out))) return TP_E_ARG;
    if (n == 0) return TP_OK;
    const bool vec = aligned16(f) && aligned16(r) && aligned16(out);
    fd_del2_radial_kernel<false><<<quads(fd_del2_radial_kernel<false>, n, 12), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        f, r, LinGrid{}, out, n, g1, g2, vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_del2_radial_apply_lin_f64(const double* f, const tp_linspace* grid, double* out, int64_t n, double g1,
                                               double g2, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!f || !out))) return TP_E_ARG;
    if (n == 0) return TP_OK;
    LinGrid lg;
    int rc = fill_lin(grid, n, lg);
    if (rc) return rc;
    const bool vec = aligned16(f) && aligned16(out);
    fd_del2_radial_kernel<true><<<quads(fd_del2_radial_kernel<true>, n, 4), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        f, nullptr, lg, out, n, g1, g2, vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_del2_radial_axpy_f64(const double* f, const double* r, const tp_linspace* grid, double* out,
                                          int64_t n, double g1, double g2, double dtau, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!f || !out || (!r && !grid))) || f == out) return TP_E_ARG;      // not in place: rows read neighbours
    if (n == 0) return TP_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (grid) {
        LinGrid lg;
        int rc = fill_lin(grid, n, lg);
        if (rc) return rc;
        fd_del2_radial_kernel<true, true><<<quads(fd_del2_radial_kernel<true, true>, n, 4), kFdThreads, 0, s>>>(
            f, nullptr, lg, out, n, g1, g2, aligned16(f) && aligned16(out), dtau);
    } else {
        fd_del2_radial_kernel<false, true><<<quads(fd_del2_radial_kernel<false, true>, n, 12), kFdThreads, 0, s>>>(
            f, r, LinGrid{}, out, n, g1, g2, aligned16(f) && aligned16(r) && aligned16(out), dtau);
    }
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_dia_apply_f64(const double* data, const int* offsets, int ndiag, const double* x, double* out,
                                   int64_t n, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || ndiag < 0 || ndiag > 5 || (n > 0 && (!x || !out)) || (ndiag > 0 && (!data || !offsets)))
        return TP_E_ARG;
    if (n == 0) return TP_OK;
    DiaArgs a;
    std::memset(&a, 0, sizeof(a));
    a.data = data; a.x = x; a.out = out; a.n = n; a.ndiag = ndiag;
    for (int k = 0; k < ndiag; ++k) {
        if (offsets[k] < -2 || offsets[k] > 2) return TP_E_ARG;
        a.off[k] = offsets[k];
    }
    fd_dia_kernel<<<quads(fd_dia_kernel, n, 5), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, aligned16(x) && aligned16(out));
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_dia_const_apply_f64(const double* coef, const int* offsets, int ndiag, const int* ov_diag,
                                         const int64_t* ov_col, const double* ov_val, int nov, const double* x,
                                         double* out, int64_t n, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || ndiag < 0 || ndiag > 5 || nov < 0 || nov > kMaxOverrides || (n > 0 && (!x || !out)) ||
        (ndiag > 0 && (!coef || !offsets)) || (nov > 0 && (!ov_diag || !ov_col || !ov_val)))
        return TP_E_ARG;
    if (n == 0) return TP_OK;
    DiaConstArgs a;
    std::memset(&a, 0, sizeof(a));
    a.x = x; a.out = out; a.n = n; a.ndiag = ndiag; a.nov = nov;
    for (int k = 0; k < ndiag; ++k) {
        if (offsets[k] < -2 || offsets[k] > 2) return TP_E_ARG;
        a.off[k] = offsets[k];
        a.coef[k] = coef[k];
    }
    for (int o = 0; o < nov; ++o) {
        if (ov_diag[o] < 0 || ov_diag[o] >= ndiag) return TP_E_ARG;
        if (!(ov_col[o] >= 0 && ov_col[o] < n && (ov_col[o] < 4 || ov_col[o] >= n - 4))) return TP_E_ARG;
        a.ov_k[o] = ov_diag[o]; a.ov_j[o] = ov_col[o]; a.ov_v[o] = ov_val[o];
    }
    const bool vec = aligned16(x) && aligned16(out);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool done = launch_dia_const_t<2, -1, 1, 0>(a, vec, s) ||       // ddx, ddr
                      launch_dia_const_t<3, -1, 0, 1>(a, vec, s) ||       // del2
                      launch_dia_const_t<3, 0, 1, 2>(a, vec, s) ||        // BC_left_extrap / avg / quad
                      launch_dia_const_t<2, 0, 1, 0>(a, vec, s) ||        // BC_left_flat
                      launch_dia_const_t<3, -2, -1, 0>(a, vec, s);        // BC_right_extrap
    if (!done) fd_dia_const_kernel<<<quads(fd_dia_const_kernel, n, 4), kFdThreads, 0, s>>>(a, vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
