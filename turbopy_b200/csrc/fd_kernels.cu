// Kernel 2 of the hot path: FiniteDifference 1-D stencils
// (turbopy/computetools.py:104-138 direct stencils, :189-223 del2_radial, and
// the banded dia_matrix operators of :140-187,225-381 applied as `D @ f`).
// Line-by-line spec: oracle/fd.py.  HBM-bound streaming kernels:
//   centered          16 B/point  (8 read + 8 written)
//   upwind_left       24 B/point  (y, cell_widths read; d written)
//   del2_radial apply 24 B/point  (f, r read; out written; coefficients rebuilt
//                                  in registers instead of 3 stored diagonals)
// Each thread produces two consecutive outputs (one 128-bit store); the +-1
// neighbours come from L1.  Compiled with -fmad=false.
#include "common.cuh"

#include <cmath>
#include <cstring>

namespace tp {

constexpr int kFdThreads = 256;

static int fd_blocks(int64_t n_pairs) {
    const DeviceProps& d = device_props();
    const int64_t full = static_cast<int64_t>(d.sm_count) * 8;
    const int64_t need = (n_pairs + kFdThreads - 1) / kFdThreads;
    return static_cast<int>(need < 1 ? 1 : (need < full ? need : full));
}

// d[i] = (y[i+1] - y[i-1]) / two_dr for 0 < i < n-1; d[0] = d[n-1] = 0.
// Quotient by the constant 2*dr: FAST policy with a per-value check, IEEE otherwise.
struct ConstDiv { double b, rb; int fast_ok; };

__device__ __forceinline__ double div_const(double a, const ConstDiv& c) {
    Guard g;
    const double q = quot<true>(a, const_recip<true>(c.b, c.rb), g);
    if (__builtin_expect(c.fast_ok && guard_ok(g) && finite3(q, q, q), 1)) return q;
    return a / c.b;
}

__global__ void __launch_bounds__(kFdThreads) fd_centered_kernel(const double* __restrict__ y,
                                                                 double* __restrict__ d, int64_t n,
                                                                 ConstDiv two_dr, bool vec) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kFdThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kFdThreads;
    const int64_t npairs = (n + 1) >> 1;
    for (int64_t pr = gtid; pr < npairs; pr += gstride) {
        const int64_t i = pr * 2;
        double o0 = 0.0, o1 = 0.0;
        if (vec && i >= 2 && i + 3 < n) {                      // interior fast path
            const double ym = y[i - 1];
            const double2 yc = ld2(y + i);
            const double yp = y[i + 2];
            o0 = div_const(yc.y - ym, two_dr);
            o1 = div_const(yp - yc.x, two_dr);
            st2(d + i, make_double2(o0, o1));
        } else {
            if (i > 0 && i < n - 1) o0 = div_const(y[i + 1] - y[i - 1], two_dr);
            d[i] = o0;
            if (i + 1 < n) {
                if (i + 1 < n - 1) o1 = div_const(y[i + 2] - y[i], two_dr);
                d[i + 1] = o1;
            }
        }
    }
}

// d[i] = (y[i] - y[i-1]) / w[i-1] for i >= 1; d[0] = 0.
__global__ void __launch_bounds__(kFdThreads) fd_upwind_kernel(const double* __restrict__ y,
                                                               const double* __restrict__ w,
                                                               double* __restrict__ d, int64_t n, bool vec) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kFdThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kFdThreads;
    const int64_t npairs = (n + 1) >> 1;
    for (int64_t pr = gtid; pr < npairs; pr += gstride) {
        const int64_t i = pr * 2;
        if (vec && i >= 2 && i + 1 < n) {
            const double ym = y[i - 1];
            const double2 yc = ld2(y + i);
            const double w0 = w[i - 1], w1 = w[i];
            st2(d + i, make_double2((yc.x - ym) / w0, (yc.y - yc.x) / w1));
        } else {
            d[i] = (i > 0) ? (y[i] - y[i - 1]) / w[i - 1] : 0.0;
            if (i + 1 < n) d[i + 1] = (y[i + 1] - y[i]) / w[i];
        }
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

__global__ void __launch_bounds__(kFdThreads) fd_del2_radial_kernel(const double* __restrict__ f,
                                                                    const double* __restrict__ r,
                                                                    double* __restrict__ out, int64_t n, double g1,
                                                                    double g2, bool vec) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kFdThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kFdThreads;
    const int64_t npairs = (n + 1) >> 1;
    const double diag = -2.0 * g2;
    const double above0 = 0.0 + 2.0 * g2;
    for (int64_t pr = gtid; pr < npairs; pr += gstride) {
        const int64_t i = pr * 2;
        if (vec && i >= 2 && i + 3 < n) {
            const double fm = f[i - 1];
            const double2 fc = ld2(f + i);
            const double fp = f[i + 2];
            const double2 rc = ld2(r + i);
            st2(out + i, make_double2(del2_row(i, n, fm, fc.x, fc.y, rc.x, g1, g2, diag, above0),
                                      del2_row(i + 1, n, fc.x, fc.y, fp, rc.y, g1, g2, diag, above0)));
        } else {
            for (int64_t j = i; j < i + 2 && j < n; ++j) {
                const double fm = j > 0 ? f[j - 1] : 0.0;
                const double fp = j < n - 1 ? f[j + 1] : 0.0;
                out[j] = del2_row(j, n, fm, f[j], fp, r[j], g1, g2, diag, above0);
            }
        }
    }
}

// Generic banded apply in scipy's dia-matvec order.
struct DiaArgs {
    const double* data; const double* x; double* out; int64_t n; int ndiag; int off[5];
};

__global__ void __launch_bounds__(kFdThreads) fd_dia_kernel(const __grid_constant__ DiaArgs a) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kFdThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kFdThreads;
    for (int64_t i = gtid; i < a.n; i += gstride) {
        double acc = 0.0;
        for (int k = 0; k < a.ndiag; ++k) {
            const int64_t j = i + a.off[k];
            if (j >= 0 && j < a.n) acc = acc + a.data[static_cast<int64_t>(k) * a.n + j] * a.x[j];
        }
        a.out[i] = acc;
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

__device__ __forceinline__ double dia_const_row(const DiaConstArgs& a, int64_t i) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (k >= a.ndiag) break;
        const int64_t j = i + a.off[k];
        if (j < 0 || j >= a.n) continue;
        double c = a.coef[k];
        if (j < 4 || j >= a.n - 4)                     // overrides only live next to the boundaries
            for (int o = 0; o < a.nov; ++o)
                if (a.ov_k[o] == k && a.ov_j[o] == j) c = a.ov_v[o];
        acc = acc + c * a.x[j];
    }
    return acc;
}

__global__ void __launch_bounds__(kFdThreads) fd_dia_const_kernel(const __grid_constant__ DiaConstArgs a, bool vec) {
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kFdThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kFdThreads;
    const int64_t npairs = (a.n + 1) >> 1;
    for (int64_t pr = gtid; pr < npairs; pr += gstride) {
        const int64_t i = pr * 2;
        const double o0 = dia_const_row(a, i);
        if (i + 1 < a.n) {
            const double o1 = dia_const_row(a, i + 1);
            if (vec) st2(a.out + i, make_double2(o0, o1));
            else { a.out[i] = o0; a.out[i + 1] = o1; }
        } else {
            a.out[i] = o0;
        }
    }
}

static ConstDiv host_recip(double b) {
    ConstDiv c;
    c.b = b;
    c.rb = 1.0 / b;
    c.fast_ok = const_divisor_ok(b) ? 1 : 0;
    return c;
}

}  // namespace tp

using namespace tp;

extern "C" int tp_fd_centered_f64(const double* y, double* d, int64_t n, double two_dr, tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!y || !d))) return TP_E_ARG;
    if (n == 0) return TP_OK;
    const bool vec = aligned16(y) && aligned16(d);
    fd_centered_kernel<<<fd_blocks((n + 1) / 2), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        y, d, n, host_recip(two_dr), vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_fd_upwind_left_f64(const double* y, const double* widths, double* d, int64_t n,
                                     tp_stream_t stream) {
    const DeviceProps& dp = device_props();
    if (dp.status != TP_OK) return dp.status;
    if (n < 0 || (n > 0 && (!y || !d)) || (n > 1 && !widths)) return TP_E_ARG;
    if (n == 0) return TP_OK;
    const bool vec = aligned16(y) && aligned16(d);
    fd_upwind_kernel<<<fd_blocks((n + 1) / 2), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(y, widths, d, n,
                                                                                                   vec);
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
    fd_del2_radial_kernel<<<fd_blocks((n + 1) / 2), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        f, r, out, n, g1, g2, vec);
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
    fd_dia_kernel<<<fd_blocks(n), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
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
    fd_dia_const_kernel<<<fd_blocks((n + 1) / 2), kFdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        a, aligned16(out));
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
