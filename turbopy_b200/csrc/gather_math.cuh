// Linear 1-D field gather at one coordinate, the three roundings the reference
// has (line-by-line spec: oracle/gather.py).
//   A  Grid.create_interpolator           turbopy/core.py:711-755
//   B  interp1d (1-D y) -> numpy.interp   turbopy/computetools.py:480
//   C  interp1d (N-D y) -> scipy _call_linear
#pragma once
#include "common.cuh"

namespace tp {

// Node coordinates + up to six field components, in shared or global memory.
struct GridView {
    const double* r;          // ng node coordinates
    const double* comp[6];    // node 0 of each component (nullptr: identically 0)
    int           stride[6];  // node stride in elements
    int           ng;
    double        r_first, r_last;
    double        inv_dr;     // only an index guess, never enters a result
    double        dr;         // Grid.dr, formula A's window half-width
};

// j with r[j] <= x < r[j+1], clamped to [0, ng-2]; numpy.interp's binary search
// result for in-range x (x == r[ng-1] is handled by the callers).  A uniform
// grid guess is verified against the actual node coordinates (Grid.r comes
// from np.linspace, not r_min + j*dr, so the guess can be off by one); a grid
// that is not uniform falls through to a binary search.
__device__ __forceinline__ int find_cell(const GridView& g, double x) {
    const int last = g.ng - 2;
    double t = (x - g.r_first) * g.inv_dr;
    int j = (t >= 0.0) ? ((t < static_cast<double>(last)) ? static_cast<int>(t) : last) : 0;
    if (x < g.r[j]) {
        if (j > 0) --j;
    } else if (j < last && x >= g.r[j + 1]) {
        ++j;
    }
    if (__builtin_expect((x < g.r[j] && j > 0) || (j < last && x >= g.r[j + 1]), 0)) {
        int lo = 0, hi = last;                 // invariant: r[lo] <= x (or lo == 0)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (g.r[mid] <= x) lo = mid; else hi = mid - 1;
        }
        j = lo;
    }
    return j;
}

// Range status shared by all formulas (interp1d's bounds check raises for
// x < r[0] or x > r[-1]; create_interpolator asserts the same, core.py:726-727).
__device__ __forceinline__ unsigned range_status(const GridView& g, double x) {
    return (x < g.r_first ? TP_ST_BELOW : 0u) | (x > g.r_last ? TP_ST_ABOVE : 0u);
}

// ---- formula B -------------------------------------------------------------
// numpy.interp:  slope = (y1-y0)/(x1-x0);  out = slope*(x-x0) + y0, with the
// exact-node shortcuts and the NaN fallbacks of the C implementation.
struct CellB {
    int j; bool exact; double dx0; Recip rdx; int jn;
};

__device__ __forceinline__ CellB locate_b(const GridView& g, double x) {
    CellB c;
    c.j = find_cell(g, x);
    if (x >= g.r_last) c.j = g.ng - 1;            // x == r[-1] -> y[-1]
    const double x0 = g.r[c.j];
    c.exact = (c.j == g.ng - 1) || (x0 == x);
    c.jn = (c.j < g.ng - 1) ? c.j + 1 : c.j;
    c.dx0 = x - x0;
    c.rdx = make_recip(g.r[c.jn] - x0);
    return c;
}

__device__ __forceinline__ double eval_b(const GridView& g, const CellB& c, int k, double x) {
    const double* y = g.comp[k];
    if (y == nullptr) return 0.0;
    const double y0 = y[static_cast<int64_t>(c.j) * g.stride[k]];
    if (c.exact) return y0;
    const double y1 = y[static_cast<int64_t>(c.jn) * g.stride[k]];
    const double slope = div_shared(y1 - y0, c.rdx);
    double out = slope * c.dx0 + y0;
    if (__builtin_expect(out != out, 0)) {        // numpy: "try the other direction"
        if (x != x) return x;                     // numpy.interp passes a NaN abscissa through
        out = slope * (x - g.r[c.jn]) + y1;
        if (out != out && y0 == y1) out = y0;
    }
    return out;
}

// ---- formula C -------------------------------------------------------------
// hi = clip(searchsorted(r, x, 'left'), 1, ng-1), lo = hi-1;
// out = ((x-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-x)/(x_hi-x_lo))*y_lo
struct CellC { int lo; double w_hi, w_lo; };

__device__ __forceinline__ CellC locate_c(const GridView& g, double x) {
    CellC c;
    int j = find_cell(g, x);                      // r[j] <= x < r[j+1], j <= ng-2
    // 'left' insertion index: first i with r[i] >= x
    int hi = (x == g.r[j]) ? j : j + 1;
    hi = hi < 1 ? 1 : hi;
    c.lo = hi - 1;
    const double x_lo = g.r[c.lo], x_hi = g.r[hi];
    const Recip rw = make_recip(x_hi - x_lo);
    c.w_hi = div_shared(x - x_lo, rw);
    c.w_lo = div_shared(x_hi - x, rw);
    return c;
}

__device__ __forceinline__ double eval_c(const GridView& g, const CellC& c, int k) {
    const double* y = g.comp[k];
    if (y == nullptr) return 0.0;
    const double y_lo = y[static_cast<int64_t>(c.lo) * g.stride[k]];
    const double y_hi = y[static_cast<int64_t>(c.lo + 1) * g.stride[k]];
    return c.w_hi * y_hi + c.w_lo * y_lo;
}

// ---- formula A -------------------------------------------------------------
// nodes with  x - dr < r[i] < x + dr  (core.py:728); one node -> y[i];
// two -> y0 + ((x - r_lo)*(y1 - y0))/(r_hi - r_lo); else the :729 assert.
struct CellA { int first; int count; double dx0; Recip rdx; };

__device__ __forceinline__ CellA locate_a(const GridView& g, double x) {
    CellA c;
    const int j = find_cell(g, x);
    const double lo_edge = x - g.dr, hi_edge = x + g.dr;
    int first = -1, count = 0;
    const int i0 = j - 2 < 0 ? 0 : j - 2;
    const int i1 = j + 3 > g.ng - 1 ? g.ng - 1 : j + 3;
    for (int i = i0; i <= i1; ++i) {
        const double ri = g.r[i];
        if (lo_edge < ri && ri < hi_edge) {
            if (first < 0) first = i;
            ++count;
        }
    }
    c.first = first < 0 ? 0 : first;
    c.count = count;
    const int nxt = c.first + 1 < g.ng ? c.first + 1 : c.first;
    c.dx0 = x - g.r[c.first];
    c.rdx = make_recip(g.r[nxt] - g.r[c.first]);
    return c;
}

__device__ __forceinline__ double eval_a(const GridView& g, const CellA& c, int k) {
    const double* y = g.comp[k];
    if (y == nullptr) return 0.0;
    const double y0 = y[static_cast<int64_t>(c.first) * g.stride[k]];
    if (c.count == 1) return y0;
    const int nxt = c.first + 1 < g.ng ? c.first + 1 : c.first;
    const double y1 = y[static_cast<int64_t>(nxt) * g.stride[k]];
    return y0 + div_shared(c.dx0 * (y1 - y0), c.rdx);
}

// Gather all six components with one formula.  Returns the TP_ST_* status;
// F is only meaningful when the status is 0.
template <int FORMULA>
__device__ __forceinline__ unsigned gather6(const GridView& g, double x, double (&F)[6]) {
    unsigned st = range_status(g, x);
    if (st) return st;
    if (FORMULA == TP_INTERP_B) {
        const CellB c = locate_b(g, x);
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] = eval_b(g, c, k, x);
    } else if (FORMULA == TP_INTERP_C) {
        const CellC c = locate_c(g, x);
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] = eval_c(g, c, k);
    } else {
        const CellA c = locate_a(g, x);
        if (c.count != 1 && c.count != 2) return TP_ST_BAD_WINDOW;
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] = eval_a(g, c, k);
    }
    return 0u;
}

}  // namespace tp
