// Linear 1-D field gather at one coordinate, the three roundings the reference
// has (line-by-line spec: oracle/gather.py).
//   A  Grid.create_interpolator           turbopy/core.py:711-755
//   B  interp1d (1-D y) -> numpy.interp   turbopy/computetools.py:480
//   C  interp1d (N-D y) -> scipy _call_linear
//
// Two evaluators per formula:
//   gather6_fast  straight-line code for the generic particle: strictly inside a
//                 cell found by the uniform-grid guess (x0 < x < x1), FAST
//                 divisions recorded in a Guard.  Returns false when the
//                 particle is not generic (on a node, guess off by one,
//                 outside the grid, NaN, formula A window not the plain two
//                 nodes); the caller then uses
//   gather6_exact the complete semantics (search, on-node shortcuts, NaN
//                 fallbacks, status codes) with IEEE divisions.
// Both produce bit-identical values where both apply (tests/test_gpu_gather.py).
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
    double        qa;         // formula A: margin of the load-free window test (window_margin below); 0 = not available
};

// ---------------------------------------------------------------------------
// fast path
// ---------------------------------------------------------------------------
// The two nodes of the particle's cell (and, for formula A, the coordinates of
// their outer neighbours), fetched either from the separate arrays of a
// GridView or from packed 64-byte node records.
struct NodePair {
    int j;                      // cell index from the uniform-grid guess
    double x0, x1;              // r[j], r[j+1]
    double xm, xp;              // r[j-1], r[j+2] (formula A only; clamped at the ends)
    double y0[6], y1[6];        // field components at the two nodes
};

__device__ __forceinline__ int guess_cell(double x, double r_first, double inv_dr, int ng) {
    // F2I saturates and maps NaN to 0; the clamp keeps the loads in bounds for any x
    const int j = __double2int_rz((x - r_first) * inv_dr);
    return max(0, min(j, ng - 2));
}

// ALL: every one of the six components is on the grid (kernel-uniform, tested once by the caller): no null checks,
// and interp_nodes() below is called with a literal mask so that its per-component selects fold away.
// QUICK (formula A): the caller holds a window margin (GridView::qa > 0), the outer neighbours are not needed.
template <int FORMULA, bool ALL = false, bool QUICK = false>
__device__ __forceinline__ void fetch_nodes(const GridView& g, double x, NodePair& np) {
    const int j = guess_cell(x, g.r_first, g.inv_dr, g.ng);
    np.j = j;
    np.x0 = g.r[j];
    np.x1 = g.r[j + 1];
    if (FORMULA == TP_INTERP_A) {
        np.xm = 0.0; np.xp = 0.0;
        if (!QUICK) {
            np.xm = g.r[max(j - 1, 0)];
            np.xp = g.r[min(j + 2, g.ng - 1)];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        np.y0[k] = 0.0; np.y1[k] = 0.0;
        if (ALL || g.comp[k] != nullptr) {
            const double* y = g.comp[k] + static_cast<int64_t>(j) * g.stride[k];
            np.y0[k] = y[0];
            np.y1[k] = y[g.stride[k]];
        }
    }
}

// Packed node records for grids that live in L2: record j = {r[j], Ex, Ey, Ez,
// Bx, By, Bz, 0}[j], 64 bytes, 64-byte aligned (written by pack_nodes_kernel each
// call).  A particle fetches its two nodes with four 256-bit loads instead of
// fourteen scattered 8-byte loads -- the L1 pays per 128-byte line touched, so
// random gathers get ~3.5x cheaper (profiles/r1_kernels.json, ng = 2^20+1).
struct RecView {
    const double* rec;
    int ng;
    double r_first, inv_dr, dr;
    unsigned present;           // bit k: component k exists
    uint64_t policy;            // L2 eviction policy of the record loads
};

__device__ __forceinline__ void ld_record(const double* p, double (&v)[8], uint64_t policy) {
    asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p), "l"(policy));
    asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]) : "l"(p + 4), "l"(policy));
}

// The same on a tp_linspace grid: the node coordinates are rebuilt in registers (common.cuh: lin_r), only the
// field values are loaded -- half the loads of a one-component gather.
template <int FORMULA>
__device__ __forceinline__ void fetch_nodes_lin(const GridView& g, const LinGrid& lg, double x, NodePair& np) {
    const int j = guess_cell(x, g.r_first, g.inv_dr, g.ng);
    np.j = j;
    const double dj = static_cast<double>(j);
    np.x0 = lin_r_at(lg, dj, 0, j);
    np.x1 = lin_r_at(lg, dj, 1, j + 1);
    if (FORMULA == TP_INTERP_A) {
        np.xm = lin_r(lg, max(j - 1, 0));
        np.xp = lin_r(lg, min(j + 2, g.ng - 1));
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        np.y0[k] = 0.0; np.y1[k] = 0.0;
        if (g.comp[k] != nullptr) {
            const double* y = g.comp[k] + static_cast<int64_t>(j) * g.stride[k];
            np.y0[k] = y[0];
            np.y1[k] = y[g.stride[k]];
        }
    }
}

// 16-byte aligned variant (record pitches that are not multiples of 32 bytes): four 128-bit loads
__device__ __forceinline__ void ld_record16(const double* p, double (&v)[8], uint64_t policy) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                     : "=d"(v[2 * q]), "=d"(v[2 * q + 1]) : "l"(p + 2 * q), "l"(policy));
}

// PITCH: doubles between consecutive node records -- 8 (packed, 256-bit loads) or kNodePitchWin (the windowed
// kernel's table, padded so that records of neighbouring nodes start in different shared-memory bank groups).
constexpr int kNodePitchWin = 10;

template <int FORMULA, int PITCH = 8>
__device__ __forceinline__ void fetch_records(const RecView& rv, double x, NodePair& np) {
    const int j = guess_cell(x, rv.r_first, rv.inv_dr, rv.ng);
    np.j = j;
    double a[8], b[8];
    const double* base = rv.rec + static_cast<int64_t>(j) * PITCH;
    if (PITCH % 4 == 0) {
        ld_record(base, a, rv.policy);
        ld_record(base + PITCH, b, rv.policy);
    } else {
        ld_record16(base, a, rv.policy);
        ld_record16(base + PITCH, b, rv.policy);
    }
    np.x0 = a[0]; np.x1 = b[0];
    if (FORMULA == TP_INTERP_A) {
        np.xm = rv.rec[static_cast<int64_t>(max(j - 1, 0)) * PITCH];
        np.xp = rv.rec[static_cast<int64_t>(min(j + 2, rv.ng - 1)) * PITCH];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { np.y0[k] = a[1 + k]; np.y1[k] = b[1 + k]; }
}

// CELL records for formula B (numpy.interp: out = slope*(x - x0) + y0 with slope = (y1-y0)/(x1-x0)): the slope
// belongs to the cell, not to the particle, so pack_cells_kernel forms it once per cell and step with the plain IEEE
// division numpy uses, and a particle needs ONE 128-byte record
//     cell j = {r[j], r[j+1], y[j] (6 components), slope_j (6 components)} + 4 pad doubles
// (112 payload bytes; the two 64-byte node records it replaces are the same number of bytes) and, per component, one
// multiply and one add -- no reciprocal, no quotient, no guard.  Absent components are stored as y = slope = +0.0.
// The table's pitch is 18 doubles (144 B), not 16: in the shared-memory window, lanes that read the same 16-byte
// chunk of DIFFERENT cells then fall into different bank groups (144/16 = 9 is odd), where a 128-byte pitch puts
// all of them on the same four banks (8-way conflicts as soon as the particle order has decayed a little:
// measured 92 % -> 61 % of the HBM peak 100 steps after a sort with the 128-byte pitch).
constexpr int kCellDoubles = 18;

__device__ __forceinline__ void ld_cell(const double* p, double (&c)[16], uint64_t policy) {
#pragma unroll
    for (int q = 0; q < 7; ++q)
        asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                     : "=d"(c[2 * q]), "=d"(c[2 * q + 1]) : "l"(p + 2 * q), "l"(policy));
    c[14] = 0.0; c[15] = 0.0;
}

__device__ __forceinline__ bool interp_cell_b(const double (&c)[16], double x, unsigned present, double (&F)[6]) {
    const double x0 = c[0], x1 = c[1];
    const double dx0 = x - x0;
    if (present == 0x3fu) {                         // all six components on the grid (kernel-uniform): no selects
#pragma unroll                                      // (ncu: 24 FSEL + their predicates per pair of particles otherwise)
        for (int k = 0; k < 6; ++k) F[k] = c[8 + k] * dx0 + c[2 + k];
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            F[k] = 0.0;
            if (present & (1u << k)) F[k] = c[8 + k] * dx0 + c[2 + k];
        }
    }
    return (x0 < x) && (x < x1);                    // strictly inside the cell (false for NaN): else the exact path
}

// Interpolate the present components from a NodePair.  `present` bit k = 0 means
// the component is identically zero (F[k] = +0.0 without touching the data).
template <int FORMULA, bool QUICK = false>
__device__ __forceinline__ bool interp_nodes(const NodePair& np, double x, double dr, int ng, unsigned present,
                                             double (&F)[6], Guard& gd, double qa = 0.0) {
    const double x0 = np.x0, x1 = np.x1;
    bool ok = (x0 < x) && (x < x1);                 // strictly inside cell j (false for NaN)
    if (FORMULA == TP_INTERP_B) {
        // numpy.interp: slope = (y1-y0)/(x1-x0); out = slope*(x-x0) + y0
        const Recip<true> rdx = make_recip<true>(x1 - x0, gd);
        const double dx0 = x - x0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            F[k] = 0.0;
            if (present & (1u << k)) F[k] = quot<true>(np.y1[k] - np.y0[k], rdx, gd) * dx0 + np.y0[k];
        }
    } else if (FORMULA == TP_INTERP_C) {
        // ((x-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-x)/(x_hi-x_lo))*y_lo
        const Recip<true> rw = make_recip<true>(x1 - x0, gd);
        const double w_hi = quot<true>(x - x0, rw, gd);
        const double w_lo = quot<true>(x1 - x, rw, gd);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            F[k] = 0.0;
            if (present & (1u << k)) F[k] = w_hi * np.y1[k] + w_lo * np.y0[k];
        }
    } else {
        // create_interpolator: the window (x-dr, x+dr) must hold exactly nodes j, j+1
        const double lo_edge = x - dr, hi_edge = x + dr;
        const double dx0 = x - x0;
        if (QUICK)              // window_margin(): at least qa (> 0) away from both nodes => r[j-1], r[j+2] are outside
            ok = ok && (lo_edge < x0) && (x1 < hi_edge) && (dx0 >= qa) && (x1 - x >= qa);
        else
            // The top cell (j == ng - 2, no node j + 2) is recognised by its data: there fetch_nodes() clamps the
            // index and hands back xp == r[j+1] == x1 (grids are strictly increasing).  Written as `j == ng - 2`,
            // ptxas 12.9 takes the test from the predicate output of the VIMNMX.RELU that clamps j in guess_cell(),
            // and the r[j+2] test then passed for every cell below the top one (found by
            // tests/test_gpu_gather.py::test_formula_a_window_margin_edges on jittered grids; on turboPy's uniform
            // grids r[j+2] lies outside the window anyway, which is why no earlier test saw it).
            ok = ok && (lo_edge < x0) && (x1 < hi_edge) && (np.j == 0 || !(lo_edge < np.xm)) &&
                 (np.xp == x1 || !(np.xp < hi_edge));
        const Recip<true> rdx = make_recip<true>(x1 - x0, gd);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            F[k] = 0.0;
            if (present & (1u << k)) F[k] = np.y0[k] + quot<true>(dx0 * (np.y1[k] - np.y0[k]), rdx, gd);
        }
    }
    return ok;
}

// Formula A without the outer-neighbour loads.  The fast path must prove that the window (x-dr, x+dr) holds exactly
// the nodes j, j+1, i.e. besides lo_edge < r[j] and r[j+1] < hi_edge also r[j-1] <= fl(x-dr) and r[j+2] >= fl(x+dr)
// (core.py:728-729).  With w = the smallest cell width of the grid, r[j-1] <= r[j] - w and r[j+2] >= r[j+1] + w, so
//     x - r[j] >= (dr - w) + eps   and   r[j+1] - x >= (dr - w) + eps,     eps >= the rounding of fl(x -+ dr),
// are sufficient: two compares on values the particle already holds instead of two more random node loads (16 of
// formula A's 128 shared-memory bytes per particle).  On turboPy's linspace grids dr - w is a few ulps, so all but
// ~1e-12 of the particles pass; those (and every particle of a grid whose cells are much narrower than Grid.dr, where
// the margin is switched off) are decided by the loads / the exact path as before.  Rounding: widths are taken
// from fl(r[j+1]-r[j]) >= true width * (1 - 2^-53); fl(x -+ dr) is off by <= 2^-53 (max|r| + dr); the computed
// differences fl(x - r[j]), fl(r[j+1] - x) are off by a factor <= 1 + 2^-53; the margin below carries a factor 2
// on top of a 4x larger eps, which covers all of them.
// Called by ALL threads of the CTA (two barriers); `slot` = 8 bytes of shared memory.
__device__ __forceinline__ double window_margin(const double* r, int ng, double dr, double r_first, double r_last,
                                                unsigned long long* slot) {
    if (threadIdx.x == 0) *slot = ~0ull;
    __syncthreads();
    unsigned long long m = ~0ull;
    for (int j = threadIdx.x; j < ng - 1; j += blockDim.x) {
        const double w = r[j + 1] - r[j];
        // positive doubles order like their bit patterns; a non-positive or NaN width switches the margin off
        m = min(m, w > 0.0 ? static_cast<unsigned long long>(__double_as_longlong(w)) : 0ull);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(slot, m);
    __syncthreads();
    const unsigned long long bits = *slot;
    if (bits == 0ull || bits == ~0ull) return 0.0;
    const double wmin = __longlong_as_double(static_cast<long long>(bits));
    const double wlb = wmin - wmin * 0x1p-52;                       // <= the true smallest width
    const double gap = fmax(dr - wlb, 0.0);
    const double eps = (fmax(fabs(r_first), fabs(r_last)) + dr) * 0x1p-51;
    const double margin = 2.0 * (gap + eps);
    return (margin < dr * 0.015625 && margin > 0.0) ? margin : 0.0;  // also false for NaN / inf
}

__device__ __forceinline__ unsigned present_mask(const GridView& g) {
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) m |= (g.comp[k] != nullptr) ? (1u << k) : 0u;
    return m;
}

// PMASK != 0: the set of components on the grid is known at COMPILE time (the specialised kernels of the common
// shapes: 0x02 = E_y only, the particle-in-field example; 0x07 = E only).  Absent components then cost nothing --
// no null checks, no zeroing, no selects: ~100 of the generic kernel's 692 warp instructions per two particles in
// the pif shape (profiles/r2_window_kernel.md), which is what a power-capped GPU wants back.
template <int FORMULA, unsigned PMASK, bool QUICK = false>
__device__ __forceinline__ bool gather6_fast_mask(const GridView& g, double x, double (&F)[6], Guard& gd) {
    NodePair np;
    const int j = guess_cell(x, g.r_first, g.inv_dr, g.ng);
    np.j = j;
    np.x0 = g.r[j];
    np.x1 = g.r[j + 1];
    if (FORMULA == TP_INTERP_A) {
        np.xm = 0.0; np.xp = 0.0;
        if (!QUICK) {
            np.xm = g.r[max(j - 1, 0)];
            np.xp = g.r[min(j + 2, g.ng - 1)];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        np.y0[k] = 0.0; np.y1[k] = 0.0;
        if (PMASK & (1u << k)) {
            const double* y = g.comp[k] + static_cast<int64_t>(j) * g.stride[k];
            np.y0[k] = y[0];
            np.y1[k] = y[g.stride[k]];
        }
    }
    return interp_nodes<FORMULA, QUICK>(np, x, g.dr, g.ng, PMASK, F, gd, g.qa);
}

template <int FORMULA, bool QUICK = false>
__device__ __forceinline__ bool gather6_fast(const GridView& g, double x, double (&F)[6], Guard& gd) {
    NodePair np;
    const unsigned present = present_mask(g);
    if (present == 0x3fu) {                          // kernel-uniform
        fetch_nodes<FORMULA, true, QUICK>(g, x, np);
        return interp_nodes<FORMULA, QUICK>(np, x, g.dr, g.ng, 0x3fu, F, gd, g.qa);
    }
    fetch_nodes<FORMULA, false, QUICK>(g, x, np);
    return interp_nodes<FORMULA, QUICK>(np, x, g.dr, g.ng, present, F, gd, g.qa);
}

template <int FORMULA>
__device__ __forceinline__ bool gather6_fast_lin(const GridView& g, const LinGrid& lg, double x, double (&F)[6], Guard& gd) {
    NodePair np;
    fetch_nodes_lin<FORMULA>(g, lg, x, np);
    return interp_nodes<FORMULA>(np, x, g.dr, g.ng, present_mask(g), F, gd);
}

template <int FORMULA>
__device__ __forceinline__ bool gather6_fast_records(const RecView& rv, double x, double (&F)[6], Guard& gd) {
    NodePair np;
    fetch_records<FORMULA>(rv, x, np);
    if (rv.present == 0x3fu) return interp_nodes<FORMULA>(np, x, rv.dr, rv.ng, 0x3fu, F, gd);
    return interp_nodes<FORMULA>(np, x, rv.dr, rv.ng, rv.present, F, gd);
}

// ---------------------------------------------------------------------------
// exact path (complete semantics, IEEE divisions)
// ---------------------------------------------------------------------------

// j with r[j] <= x < r[j+1], clamped to [0, ng-2]: numpy.interp's binary
// search result for in-range x.  Grid.r comes from np.linspace, not
// r_min + j*dr, so the uniform guess can be off by one; a grid that is not
// uniform at all falls through to a binary search.
static __device__ __noinline__ int find_cell(const GridView& g, double x) {
    const int last = g.ng - 2;
    const double t = (x - g.r_first) * g.inv_dr;
    int j = (t >= 0.0) ? ((t < static_cast<double>(last)) ? static_cast<int>(t) : last) : 0;
    if (x < g.r[j]) {
        if (j > 0) --j;
    } else if (j < last && x >= g.r[j + 1]) {
        ++j;
    }
    if ((x < g.r[j] && j > 0) || (j < last && x >= g.r[j + 1])) {
        int lo = 0, hi = last;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (g.r[mid] <= x) lo = mid; else hi = mid - 1;
        }
        j = lo;
    }
    return j;
}

__device__ __forceinline__ double node_value(const GridView& g, int k, int j) {
    return g.comp[k][static_cast<int64_t>(j) * g.stride[k]];
}

// Returns the TP_ST_* status; F is only meaningful when it is 0.
template <int FORMULA>
__device__ __noinline__ unsigned gather6_exact(const GridView& g, double x, double (&F)[6]) {
    // interp1d's bounds check raises for x < r[0] or x > r[-1]; create_interpolator
    // asserts the same (core.py:726-727)
    const unsigned st = (x < g.r_first ? TP_ST_BELOW : 0u) | (x > g.r_last ? TP_ST_ABOVE : 0u);
    if (st) return st;
    const int ng = g.ng;
    if (FORMULA == TP_INTERP_B) {
        int j = find_cell(g, x);
        if (x >= g.r_last) j = ng - 1;                       // x == r[-1] -> y[-1]
        const double x0 = g.r[j];
        const bool exact = (j == ng - 1) || (x0 == x);
        const int jn = (j < ng - 1) ? j + 1 : j;
        const double dx = g.r[jn] - x0;
        for (int k = 0; k < 6; ++k) {
            if (g.comp[k] == nullptr) { F[k] = 0.0; continue; }
            const double y0 = node_value(g, k, j);
            if (exact) { F[k] = y0; continue; }
            const double y1 = node_value(g, k, jn);
            const double slope = (y1 - y0) / dx;
            double out = slope * (x - x0) + y0;
            if (out != out) {                                // numpy: "try the other direction"
                if (x != x) out = x;                         // a NaN abscissa passes through
                else {
                    out = slope * (x - g.r[jn]) + y1;
                    if (out != out && y0 == y1) out = y0;
                }
            }
            F[k] = out;
        }
    } else if (FORMULA == TP_INTERP_C) {
        const int j = find_cell(g, x);
        int hi = (x == g.r[j]) ? j : j + 1;                  // 'left' insertion index
        hi = hi < 1 ? 1 : hi;
        const int lo = hi - 1;
        const double x_lo = g.r[lo], x_hi = g.r[hi];
        const double w_hi = (x - x_lo) / (x_hi - x_lo);
        const double w_lo = (x_hi - x) / (x_hi - x_lo);
        for (int k = 0; k < 6; ++k) {
            if (g.comp[k] == nullptr) { F[k] = 0.0; continue; }
            F[k] = w_hi * node_value(g, k, hi) + w_lo * node_value(g, k, lo);
        }
    } else {
        const int j = find_cell(g, x);
        const double lo_edge = x - g.dr, hi_edge = x + g.dr;
        int first = -1, count = 0;
        const int i0 = j - 2 < 0 ? 0 : j - 2;
        const int i1 = j + 3 > ng - 1 ? ng - 1 : j + 3;
        for (int i = i0; i <= i1; ++i) {
            const double ri = g.r[i];
            if (lo_edge < ri && ri < hi_edge) {
                if (first < 0) first = i;
                ++count;
            }
        }
        if (count != 1 && count != 2) return TP_ST_BAD_WINDOW;    // core.py:729
        const int nxt = first + 1 < ng ? first + 1 : first;
        for (int k = 0; k < 6; ++k) {
            if (g.comp[k] == nullptr) { F[k] = 0.0; continue; }
            const double y0 = node_value(g, k, first);
            if (count == 1) { F[k] = y0; continue; }              // core.py:731-732
            const double y1 = node_value(g, k, nxt);
            F[k] = y0 + ((x - g.r[first]) * (y1 - y0)) / (g.r[nxt] - g.r[first]);   // :752-753
        }
    }
    return 0u;
}

}  // namespace tp
