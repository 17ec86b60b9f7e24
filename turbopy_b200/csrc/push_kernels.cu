// Kernel 1 of the hot path: BorisPush.push and the fused linear gather + push.
//
//   push_kernel         reference: turbopy/computetools.py:407-441
//   gather_push_kernel  reference: the gather of ChargedParticle.update
//                       (tests/fixtures/particle_in_field/particle_in_field.py:61-63;
//                       turbopy/core.py:711-755, turbopy/computetools.py:459-481)
//                       fused with the push above
//   interp_kernel       reference: Interpolators.interpolate1D(x, y)(x_new)
//
// HBM-bound streaming kernels: 96 B of particle state per particle-step
// (6 doubles read + 6 written), fp64 math in registers, no tensor cores.  The
// fast path reads/writes the SoA arrays with 128-bit accesses (two particles per
// thread per array); the 1-D node coordinates and field components are staged
// once per CTA into shared memory with TMA bulk copies (cp.async.bulk +
// mbarrier), overlapped with the CTA's first particle loads.
#include "boris_math.cuh"
#include "gather_math.cuh"
#include "particle_io.cuh"
#include "grid_stage.cuh"
#include "tma_ring.cuh"
#include "tma_ring_fields.cuh"
#include "tma_ring_window.cuh"
#include "moments_math.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace tp {

// ---------------------------------------------------------------------------
// push (E, B given per call)
// ---------------------------------------------------------------------------
// Recovery path of one particle: reload it from global memory (nothing has been
// stored yet) and redo the step with IEEE divisions.
__device__ __noinline__ void push_one_exact(const PushArgs& a, int64_t i) {
    const Vec3 Eu = uniform_of(a.E), Bu = uniform_of(a.B);     // rebuilt here: no references into the hot loop
    Vec3 x, p;
    load_one(a.p, i, x, p);
    const Vec3 eq = dt_field_q(a.k, field_at(a.E, Eu, i)), bq = dt_field_q(a.k, field_at(a.B, Bu, i));
    Guard g;
    for (int s = 0; s < a.nsteps; ++s) boris_step<false>(a.k, eq, bq, x, p, g);
    store_one(a.p, i, x, p);
}

// nsteps straight-line steps of one particle; false if any of them left the fast path's domain
__device__ __forceinline__ bool push_steps_fast(const PushK& k, int nsteps, const Vec3& eq, const Vec3& bq, Vec3& x, Vec3& p) {
    Guard g;
    bool fin = true;
    for (int s = 0; s < nsteps; ++s) {
        boris_step<true>(k, eq, bq, x, p, g);
        fin = fin && finite3(x.x, x.y, x.z);
    }
    return guard_ok(g) && fin && k.fast_ok;
}

__device__ __forceinline__ bool push_one_fast(const PushArgs& a, const Vec3& Eu, const Vec3& Bu, int64_t i,
                                              Vec3& x, Vec3& p) {
    return push_steps_fast(a.k, a.nsteps, dt_field_q(a.k, field_at(a.E, Eu, i)), dt_field_q(a.k, field_at(a.B, Bu, i)), x, p);
}

template <bool VEC2>
__global__ void __launch_bounds__(kThreads, TP_MIN_BLOCKS) push_kernel(const __grid_constant__ PushArgs a) {
    const Vec3 Eu = uniform_of(a.E), Bu = uniform_of(a.B);
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kThreads;
    if (VEC2) {
        const int64_t npairs = a.p.n >> 1;
        for (int64_t pr = gtid; pr < npairs; pr += gstride) {
            const int64_t i = pr * 2;
            Vec3 x[2], p[2];
            load_pair(a.p, i, x, p);
            prefetch_tile(a.p, (pr - threadIdx.x + TP_PF_DIST * gstride) * 2, npairs * 2);
            const bool ok0 = push_one_fast(a, Eu, Bu, i, x[0], p[0]);
            const bool ok1 = push_one_fast(a, Eu, Bu, i + 1, x[1], p[1]);
            if (__builtin_expect(ok0 && ok1, 1)) {
                store_pair(a.p, i, x, p);
            } else {
                if (ok0) store_one(a.p, i, x[0], p[0]); else push_one_exact(a,i);
                if (ok1) store_one(a.p, i + 1, x[1], p[1]); else push_one_exact(a,i + 1);
            }
        }
        if ((a.p.n & 1) && gtid == 0) push_one_exact(a,a.p.n - 1);     // odd tail
    } else {
        for (int64_t i = gtid; i < a.p.n; i += gstride) {
            Vec3 x, p;
            load_one(a.p, i, x, p);
            if (__builtin_expect(push_one_fast(a, Eu, Bu, i, x, p), 1)) store_one(a.p, i, x, p);
            else push_one_exact(a,i);
        }
    }
}

// Generic particle: ONE straight-line gather + push step, validity in the return value.  A wrong cell guess or
// a particle that is off the grid makes it false; the clamped index keeps the loads in bounds.  Multi-step passes
// (GatherArgs::nsteps) loop around this at the call sites, both particles of a thread inside one iteration, so
// that the two independent dependency chains stay interleaved.
template <int FORMULA, unsigned PMASK = 0u, bool QUICK = false>
__device__ __forceinline__ bool gather_step_fast(const GatherArgs& a, const GridView& g, Vec3& x, Vec3& p) {
    double F[6];
    Guard gd;
    bool ok = PMASK ? gather6_fast_mask<FORMULA, PMASK, QUICK>(g, axis_of(x, a.axis), F, gd)
                    : gather6_fast<FORMULA, QUICK>(g, axis_of(x, a.axis), F, gd);
    Vec3 E, B;
    E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
    boris_step<true>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x, p, gd);
    return ok && guard_ok(gd) && finite3(x.x, x.y, x.z) && a.k.fast_ok;
}

// The same with the two nodes fetched from packed records (grids that live in L2).
template <int FORMULA>
__device__ __forceinline__ bool gather_step_fast_records(const GatherArgs& a, const RecView& rv, Vec3& x, Vec3& p) {
    double F[6];
    Guard gd;
    bool ok = gather6_fast_records<FORMULA>(rv, axis_of(x, a.axis), F, gd);
    Vec3 E, B;
    E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
    boris_step<true>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x, p, gd);
    return ok && guard_ok(gd) && finite3(x.x, x.y, x.z) && a.k.fast_ok;
}

// nsteps steps of one particle (scalar tails)
template <int FORMULA>
__device__ __forceinline__ bool gather_steps_fast(const GatherArgs& a, const GridView& g, Vec3& x, Vec3& p) {
    bool ok = true;
    for (int s = 0; s < a.nsteps; ++s) ok = gather_step_fast<FORMULA>(a, g, x, p) && ok;
    return ok;
}

__device__ __forceinline__ RecView make_rec_view(const GatherArgs& a, int keep_in_l2) {
    RecView rv;
    rv.policy = keep_in_l2 ? l2_evict_last() : l2_evict_normal();
    rv.rec = a.records; rv.ng = a.ng; rv.r_first = a.r_first; rv.inv_dr = a.inv_dr; rv.dr = a.dr;
    rv.present = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) rv.present |= a.comp[k] ? (1u << k) : 0u;
    return rv;
}

// record j = {r[j], Ex[j], Ey[j], Ez[j], Bx[j], By[j], Bz[j], 0}; absent components are stored as +0.0
template <int PITCH>
__global__ void __launch_bounds__(256) pack_nodes_kernel(const __grid_constant__ GatherArgs a, double* __restrict__ rec) {
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < a.ng; j += gstride) {
        double v[8];
        v[0] = a.r[j];
#pragma unroll
        for (int k = 0; k < 6; ++k) v[1 + k] = a.comp[k] ? a.comp[k][j * a.stride[k]] : 0.0;
        v[7] = 0.0;
        double* dst = rec + j * PITCH;
#pragma unroll
        for (int q = 0; q < 4; ++q) st2(dst + 2 * q, make_double2(v[2 * q], v[2 * q + 1]));
    }
}

// cell j = {r[j], r[j+1], y[j][0..5], slope_j[0..5], 0, 0} with slope = (y[j+1]-y[j])/(r[j+1]-r[j]) by IEEE division,
// exactly numpy.interp's per-cell quantity (gather_math.cuh: interp_cell_b); absent components as +0.0.
// Each warp packs 32 consecutive cells (4608 contiguous bytes of the table): the records are assembled in a per-warp
// shared-memory slab and written out with fully coalesced 128-bit stores (a thread writing its own 144-byte-pitch
// record directly touches 32 half-filled sectors per store instruction: 61 us per 2^20 cells; this way ~35).
__global__ void __launch_bounds__(256) pack_cells_kernel(const __grid_constant__ GatherArgs a, double* __restrict__ rec) {
    __shared__ __align__(16) double slab[8][32 * kCellDoubles];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ncell = a.ng - 1;
    const int64_t nchunk = (ncell + 31) / 32;                       // chunks of 32 cells, one per warp and iteration
    const int64_t wstride = static_cast<int64_t>(gridDim.x) * 8;
    for (int64_t c = static_cast<int64_t>(blockIdx.x) * 8 + warp; c < nchunk; c += wstride) {
        const int64_t j = c * 32 + lane;
        double* mine = slab[warp] + lane * kCellDoubles;
        if (j < ncell) {
            const double x0 = a.r[j], x1 = a.r[j + 1];
            const double dx = x1 - x0;
            mine[0] = x0; mine[1] = x1;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double y0 = 0.0, sl = 0.0;
                if (a.comp[k]) {
                    y0 = a.comp[k][j * a.stride[k]];
                    sl = (a.comp[k][(j + 1) * a.stride[k]] - y0) / dx;
                }
                mine[2 + k] = y0; mine[8 + k] = sl;
            }
#pragma unroll
            for (int k = 14; k < kCellDoubles; ++k) mine[k] = 0.0;
        }
        __syncwarp();
        const int64_t first = c * 32;
        const int live = static_cast<int>(min(static_cast<int64_t>(32), ncell - first));
        const int nvec = live * (kCellDoubles / 2);                  // 16-byte pieces of this chunk
        double* dst = rec + first * kCellDoubles;
        for (int v = lane; v < nvec; v += 32) st2(dst + 2 * v, *reinterpret_cast<const double2*>(slab[warp] + 2 * v));
        __syncwarp();
    }
}

// Recovery path: reload the particle (nothing has been stored yet), complete
// gather semantics, IEEE divisions; out-of-grid particles are flagged and left
// untouched.  The out-of-line entry points below take only the grid-constant
// argument block and the shared-memory base and rebuild their own GridView:
// passing the hot loop's view or particle registers by reference would force
// them into local memory (LDL/STL in the hot loop, profiles/r1_v3).
template <int FORMULA>
__device__ __forceinline__ void exact_core(const GatherArgs& a, const GridView& g, int64_t i, Vec3& x, Vec3& p) {
    load_one(a.p, i, x, p);
    for (int s = 0; s < a.nsteps; ++s) {
        double F[6];
        const unsigned st = gather6_exact<FORMULA>(g, axis_of(x, a.axis), F);
        if (st != 0u) {                 // off the grid: flagged, keeps the state it had when it left (the input if s == 0)
            flag_particle(a.status, i, st);
            return;
        }
        Vec3 E, B;
        E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
        Guard gd;
        boris_step<false>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x, p, gd);
    }
}

template <int FORMULA, bool STAGED>
__device__ __noinline__ void gather_step_exact(const GatherArgs& a, unsigned char* smem, int64_t i) {
    const GridView g = make_view<STAGED>(a, smem);
    Vec3 x, p;
    exact_core<FORMULA>(a, g, i, x, p);
    store_one(a.p, i, x, p);            // a flagged particle stores back what it held (bit-identical to untouched for one step)
}

// ---------------------------------------------------------------------------
// fused gather + push
// ---------------------------------------------------------------------------
template <bool VEC2, int FORMULA, bool STAGED>
__global__ void __launch_bounds__(kThreads, TP_MIN_BLOCKS) gather_push_kernel(const __grid_constant__ GatherArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (STAGED) stage_issue(a, smem);
    const GridView g = make_view<STAGED>(a, smem);
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kThreads;
    // The loops are rotated: the first particle loads are issued before the
    // (CTA-uniform) wait for the staged grid, so they overlap the TMA copies.
    if (VEC2) {
        const int64_t npairs = a.p.n >> 1;
        int64_t pr = gtid;
        Vec3 x[2], p[2];
        if (pr < npairs) load_pair(a.p, pr * 2, x, p);
        if (STAGED) stage_wait(smem);
        while (pr < npairs) {
            const int64_t i = pr * 2;
            prefetch_tile(a.p, (pr - threadIdx.x + TP_PF_DIST * gstride) * 2, npairs * 2);
            bool ok0 = true, ok1 = true;
            for (int st = 0; st < a.nsteps; ++st) {
                ok0 = gather_step_fast<FORMULA>(a, g, x[0], p[0]) && ok0;
                ok1 = gather_step_fast<FORMULA>(a, g, x[1], p[1]) && ok1;
            }
            if (__builtin_expect(ok0 && ok1, 1)) {
                store_pair(a.p, i, x, p);
            } else {
                if (ok0) store_one(a.p, i, x[0], p[0]); else gather_step_exact<FORMULA, STAGED>(a, smem, i);
                if (ok1) store_one(a.p, i + 1, x[1], p[1]); else gather_step_exact<FORMULA, STAGED>(a, smem, i + 1);
            }
            pr += gstride;
            if (pr < npairs) load_pair(a.p, pr * 2, x, p);
        }
        if ((a.p.n & 1) && gtid == 0) gather_step_exact<FORMULA, STAGED>(a, smem, a.p.n - 1);
    } else {
        int64_t i = gtid;
        Vec3 x, p;
        if (i < a.p.n) load_one(a.p, i, x, p);
        if (STAGED) stage_wait(smem);
        while (i < a.p.n) {
            if (__builtin_expect(gather_steps_fast<FORMULA>(a, g, x, p), 1)) store_one(a.p, i, x, p);
            else gather_step_exact<FORMULA, STAGED>(a, smem, i);
            i += gstride;
            if (i < a.p.n) load_one(a.p, i, x, p);
        }
    }
}

// ---- fused gather + push ---------------------------------------------------
struct TmaArgs {
    GatherArgs g;
    RingArgs ring;
    MomConsts mom;              // MOM instances: constants of the fused moment diagnostic ...
    double* mom_partials;       // ... and where CTA b leaves its five partial sums (8 doubles apart)
};

// The moment diagnostic's share of one thread: the two momenta of a tile slot (or one particle of the ragged tail),
// fast arithmetic under a guard of its own, redone with IEEE operations if an operand left the fast path's range.
// The constants are the push's own (mass**2, c2, 1/c2 from PushK -- the host only fuses when the diagnostic's agree)
// plus the diagnostic's mass: |p|^2 and m1 = sqrt(mass**2 + |p|^2/c2) are then literally the first quantities
// boris_step forms from the same momentum, and the compiler merges them; what the diagnostic adds per particle is one
// reciprocal, one quotient and five additions.
__device__ __forceinline__ void moments_add(const Vec3* p, int count, const PushK& pk, double mass, double (&acc)[kRedVals]) {
    MomConsts k;
    k.mass = mass; k.mm = pk.mm; k.c2 = pk.c2; k.rc2 = pk.rc2; k.fast_ok = pk.fast_ok;
    double c[kRedVals] = {0.0, 0.0, 0.0, 0.0, 0.0};
    Guard g;
    for (int v = 0; v < count; ++v) accumulate<true>(p[v].x, p[v].y, p[v].z, k, g, c);
    if (__builtin_expect(!(guard_ok(g) && k.fast_ok), 0)) {
#pragma unroll
        for (int q = 0; q < kRedVals; ++q) c[q] = 0.0;
        for (int v = 0; v < count; ++v) accumulate<false>(p[v].x, p[v].y, p[v].z, k, g, c);
    }
#pragma unroll
    for (int q = 0; q < kRedVals; ++q) acc[q] += c[q];
}


template <int FORMULA, bool STAGED>
__device__ __noinline__ void gather_step_exact_slot(const GatherArgs& a, unsigned char* smem, int64_t i, double* xs,
                                                    double* ps, int pitch) {
    const GridView g = make_view<STAGED>(a, smem);
    Vec3 x, p;
    exact_core<FORMULA>(a, g, i, x, p);          // flagged: x, p hold the state the particle left the grid with
    slot_write(xs, ps, pitch, x, p);
}

// GMODE: 0 = field arrays read from global memory / L2, 1 = arrays staged in
// shared memory by the TMA, 2 = packed node records in L2 (fast path only; the
// recovery path always reads the original arrays).
// PMASK: 0 = which components are on the grid is a run-time property (any shape); otherwise the compile-time mask of
// a specialised shape (staged grids only: see gather6_fast_mask).
// MOM: also accumulate {KE, Px, Py, Pz, |p|^2} of the momenta AS LOADED, one partial per CTA in ta.mom_partials
// (see gather_push_win_kernel).
template <int FORMULA, int GMODE, bool AOS, unsigned PMASK = 0u, bool MOM = false>
__global__ void __launch_bounds__(kTmaThreads, 1) gather_push_tma_kernel(const __grid_constant__ TmaArgs ta) {
    constexpr bool STAGED = GMODE == 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const GatherArgs& a = ta.g;
    ring_init(ta.ring, smem);
    if (STAGED) stage_issue(a, smem); else __syncthreads();       // (stage_issue syncs after its own init)
    ring_prologue<AOS>(a.p, ta.ring, smem);
    GridView g = make_view<STAGED>(a, smem);
    if (STAGED) stage_wait(smem);
    if (STAGED && FORMULA == TP_INTERP_A)           // formula A without the outer-neighbour loads (gather_math.cuh)
        g.qa = window_margin(g.r, g.ng, g.dr, g.r_first, g.r_last, reinterpret_cast<unsigned long long*>(smem + kMarginSlot));
    const RecView rv = make_rec_view(a, ta.ring.stream_hint);
    auto exact = [&](int64_t i, double* xs, double* ps, int pitch) {
        gather_step_exact_slot<FORMULA, STAGED>(a, smem, i, xs, ps, pitch);
    };
    double acc[kRedVals] = {0.0, 0.0, 0.0, 0.0, 0.0};
    auto observe = [&](const Vec3 (&p)[2]) {
        if (MOM) moments_add(p, 2, a.k, ta.mom.mass, acc);
    };
    if (STAGED && FORMULA == TP_INTERP_A && g.qa > 0.0)      // (kernel-uniform) the ring instance without the neighbour loads
        tma_ring<AOS>(a.p, ta.ring, smem, [&](Vec3& x, Vec3& p) { return gather_step_fast<FORMULA, PMASK, true>(a, g, x, p); }, exact,
                      observe);
    else
        tma_ring<AOS>(
            a.p, ta.ring, smem,
            [&](Vec3& x, Vec3& p) {
                return GMODE == 2 ? gather_step_fast_records<FORMULA>(a, rv, x, p) : gather_step_fast<FORMULA, PMASK>(a, g, x, p);
            },
            exact, observe);
    // ragged tail (< kTile particles) on the last CTA, scalar path
    if (blockIdx.x == gridDim.x - 1) {
        for (int64_t i = ta.ring.ntiles * kTile + threadIdx.x; i < a.p.n; i += kTmaThreads) {   // all 16 warps
            Vec3 x, p;
            load_one(a.p, i, x, p);
            if (MOM) moments_add(&p, 1, a.k, ta.mom.mass, acc);
            if (gather_step_fast<FORMULA>(a, g, x, p)) store_one(a.p, i, x, p);
            else gather_step_exact<FORMULA, STAGED>(a, smem, i);
        }
    }
    if (MOM) block_sum<kTmaThreads>(acc, ta.mom_partials + static_cast<int64_t>(blockIdx.x) * 8);   // all 16 warps
}

// ---- fused gather + push, node records staged per tile (large grids, cell-ordered particles) ----------
struct TmaWinArgs {
    GatherArgs g;
    RingArgs ring;
    WindowArgs win;
    MomConsts mom;              // MOM instances: constants of the fused moment diagnostic ...
    double* mom_partials;       // ... and where CTA b leaves its five partial sums (8 doubles apart)
};

template <int FORMULA>
__device__ __forceinline__ void gather_pair_fast_window(const GatherArgs& a, const RecView& rv, const WinView& w,
                                                        Vec3 (&x)[2], Vec3 (&p)[2], bool& ok0, bool& ok1) {
    NodePair n0, n1;
    const double xa = axis_of(x[0], a.axis), xb = axis_of(x[1], a.axis);
    fetch_pair_window<FORMULA>(rv, w, xa, xb, n0, n1);
    {
        double F[6];
        Guard gd;
        const bool ok = rv.present == 0x3fu ? interp_nodes<FORMULA>(n0, xa, rv.dr, rv.ng, 0x3fu, F, gd)
                                            : interp_nodes<FORMULA>(n0, xa, rv.dr, rv.ng, rv.present, F, gd);
        Vec3 E, B;
        E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
        boris_step<true>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x[0], p[0], gd);
        ok0 = ok && guard_ok(gd) && finite3(x[0].x, x[0].y, x[0].z) && a.k.fast_ok;
    }
    {
        double F[6];
        Guard gd;
        const bool ok = rv.present == 0x3fu ? interp_nodes<FORMULA>(n1, xb, rv.dr, rv.ng, 0x3fu, F, gd)
                                            : interp_nodes<FORMULA>(n1, xb, rv.dr, rv.ng, rv.present, F, gd);
        Vec3 E, B;
        E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
        boris_step<true>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x[1], p[1], gd);
        ok1 = ok && guard_ok(gd) && finite3(x[1].x, x[1].y, x[1].z) && a.k.fast_ok;
    }
}

// formula B on cell records: no division in the gather at all
__device__ __forceinline__ void gather_pair_fast_cells(const GatherArgs& a, const RecView& rv, const WinView& w,
                                                       Vec3 (&x)[2], Vec3 (&p)[2], bool& ok0, bool& ok1) {
    double c0[16], c1[16];
    const double xa = axis_of(x[0], a.axis), xb = axis_of(x[1], a.axis);
    fetch_pair_cells(rv, w, xa, xb, c0, c1);
    {
        double F[6];
        Guard gd;
        const bool ok = interp_cell_b(c0, xa, rv.present, F);
        Vec3 E, B;
        E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
        boris_step<true>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x[0], p[0], gd);
        ok0 = ok && guard_ok(gd) && finite3(x[0].x, x[0].y, x[0].z) && a.k.fast_ok;
    }
    {
        double F[6];
        Guard gd;
        const bool ok = interp_cell_b(c1, xb, rv.present, F);
        Vec3 E, B;
        E.x = F[0]; E.y = F[1]; E.z = F[2]; B.x = F[3]; B.y = F[4]; B.z = F[5];
        boris_step<true>(a.k, dt_field_q(a.k, E), dt_field_q(a.k, B), x[1], p[1], gd);
        ok1 = ok && guard_ok(gd) && finite3(x[1].x, x[1].y, x[1].z) && a.k.fast_ok;
    }
}

// MOM: the kernel also accumulates {KE, Px, Py, Pz, |p|^2} of the momenta AS LOADED (i.e. before this push) and
// leaves one partial per CTA in ta.mom_partials -- the energy / moment diagnostic of the step without a second pass
// over the momentum arrays (24 B per particle saved).  Deterministic: static tile -> CTA -> thread assignment,
// per-thread serial sums, shuffle trees, fixed-order fold of the CTA partials (moments_final_kernel).
template <int FORMULA, bool AOS, bool MOM = false>
__global__ void __launch_bounds__(kTmaThreads, 1) gather_push_win_kernel(const __grid_constant__ TmaWinArgs ta) {
    constexpr bool CELLS = FORMULA == TP_INTERP_B;          // (the host packs the matching record table)
    extern __shared__ __align__(128) unsigned char smem[];
    const GatherArgs& a = ta.g;
    ring_init(ta.ring, smem);
    __syncthreads();
    const RecView rv = make_rec_view(a, ta.ring.stream_hint);
    double acc[kRedVals] = {0.0, 0.0, 0.0, 0.0, 0.0};
    tma_ring_window<AOS, CELLS>(
        a.p, ta.ring, ta.win, rv, a.axis, smem,
        [&](Vec3 (&x)[2], Vec3 (&p)[2], const WinView& w, bool& ok0, bool& ok1) {
            if (CELLS) gather_pair_fast_cells(a, rv, w, x, p, ok0, ok1);
            else gather_pair_fast_window<FORMULA>(a, rv, w, x, p, ok0, ok1);
        },
        [&](int64_t i, double* xs, double* ps, int pitch) {
            gather_step_exact_slot<FORMULA, false>(a, smem, i, xs, ps, pitch);
        },
        [&](const Vec3 (&p)[2]) {
            if (MOM) moments_add(p, 2, a.k, ta.mom.mass, acc);
        });
    // ragged tail (< kTile particles) on the last CTA, scalar path over the original arrays
    if (blockIdx.x == gridDim.x - 1) {
        const GridView g = make_view<false>(a, smem);
        for (int64_t i = ta.ring.ntiles * kTile + threadIdx.x; i < a.p.n; i += kTmaThreads) {
            Vec3 x, p;
            load_one(a.p, i, x, p);
            if (MOM) moments_add(&p, 1, a.k, ta.mom.mass, acc);
            if (gather_step_fast<FORMULA>(a, g, x, p)) store_one(a.p, i, x, p);
            else gather_step_exact<FORMULA, false>(a, smem, i);
        }
    }
    if (MOM) block_sum<kTmaThreads>(acc, ta.mom_partials + static_cast<int64_t>(blockIdx.x) * 8);   // all 16 warps
}

// Cell range of every full tile from the particles' CURRENT coordinates: primes the per-tile windows on the
// first call for an ensemble and after a re-ordering (tp_sort_by_cell).  Reads 8 B per particle.
__global__ void __launch_bounds__(256) tile_bounds_kernel(const double* __restrict__ coord, int64_t stride, int64_t ntiles,
                                                          double r_first, double inv_dr, int ng, int2* __restrict__ bounds) {
    __shared__ int smin[8], smax[8];
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int lo = 0x7fffffff, hi = static_cast<int>(0x80000000u);
        for (int i = threadIdx.x; i < kTile; i += 256) {
            const double x = coord[(t * kTile + i) * stride];
            const bool in = x >= r_first && (x - r_first) * inv_dr < static_cast<double>(ng);     // false for NaN
            if (in) {
                const int j = guess_cell(x, r_first, inv_dr, ng);
                lo = min(lo, j); hi = max(hi, j);
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) { lo = min(lo, smin[w]); hi = max(hi, smax[w]); }
            bounds[t] = hi >= lo ? make_int2(lo, hi - lo + 1) : make_int2(0, 0);
        }
        __syncthreads();
    }
}

// ---- BorisPush.push with uniform E, B (configs 1 and 4 call push() directly) --
struct PushTmaArgs {
    PushArgs a;
    RingArgs ring;
    double mom_mass;            // MOM instance: the diagnostic's mass ...
    double* mom_partials;       // ... and where CTA b leaves its five partial sums (8 doubles apart)
};

__device__ __noinline__ void push_one_exact_slot(const PushArgs& a, int64_t i, double* xs, double* ps, int pitch) {
    const Vec3 Eu = uniform_of(a.E), Bu = uniform_of(a.B);
    Vec3 x, p;
    load_one(a.p, i, x, p);                    // the tile has not been stored yet: global memory holds the input
    const Vec3 eq = dt_field_q(a.k, Eu), bq = dt_field_q(a.k, Bu);
    Guard g;
    for (int s = 0; s < a.nsteps; ++s) boris_step<false>(a.k, eq, bq, x, p, g);
    slot_write(xs, ps, pitch, x, p);
}

// ---- BorisPush.push with per-particle (n,3) E and B ------------------------------
__device__ __noinline__ void push_fields_exact_slot(const PushArgs& a, int64_t i, double* xs, double* ps, int pitch) {
    const Vec3 none = {0.0, 0.0, 0.0};
    Vec3 x, p;
    load_one(a.p, i, x, p);                    // global memory still holds the input
    const Vec3 eq = dt_field_q(a.k, field_at(a.E, none, i)), bq = dt_field_q(a.k, field_at(a.B, none, i));
    Guard g;
    for (int s = 0; s < a.nsteps; ++s) boris_step<false>(a.k, eq, bq, x, p, g);
    slot_write(xs, ps, pitch, x, p);
}

template <bool PA, bool FA>
__global__ void __launch_bounds__(kTmaThreads, 1) push_fields_tma_kernel(const __grid_constant__ PushTmaArgs ta) {
    extern __shared__ __align__(128) unsigned char smem[];
    const PushArgs& a = ta.a;
    ring_init(ta.ring, smem);
    __syncthreads();
    tma_ring_fields<PA, FA, true>(
        a, ta.ring, smem,
        [&](const Vec3& E, const Vec3& B, Vec3& x, Vec3& p) {
            return push_steps_fast(a.k, a.nsteps, dt_field_q(a.k, E), dt_field_q(a.k, B), x, p);
        },
        [&](int64_t i, double* xs, double* ps, int pitch) { push_fields_exact_slot(a, i, xs, ps, pitch); });
    if (blockIdx.x == gridDim.x - 1) {                       // ragged tail, scalar path
        const Vec3 none = {0.0, 0.0, 0.0};
        for (int64_t i = ta.ring.ntiles * kTileF + threadIdx.x; i < a.p.n; i += kTmaThreads) {
            Vec3 x, p;
            load_one(a.p, i, x, p);
            if (push_one_fast(a, none, none, i, x, p)) store_one(a.p, i, x, p);
            else push_one_exact(a, i);
        }
    }
}

// Uniform E, B on the small-tile ring: one particle per consumer thread, 480-particle tiles, 22.5 KB stages,
// kPushStages of them.  Measured (16 Mi particles): this ring 3 / 4 / 5 / 6 stages = 91.7 / 98.0 / 92.1 / 88.5 % of the
// HBM peak, the 960-particle ring of the fused kernel 2 / 3 / 4 stages = 92.9 / 94.1 / 89.6 %.  (The fused gather+push
// is the other way round: 98 % on the 960-particle ring, 83-90 % here.)
// MOM: also accumulate {KE, Px, Py, Pz, |p|^2} of the momenta AS LOADED (see gather_push_win_kernel).
constexpr int kPushStages = 4;
template <bool PA, bool MOM = false>
__global__ void __launch_bounds__(kTmaThreads, 1) push_tma_kernel(const __grid_constant__ PushTmaArgs ta) {
    extern __shared__ __align__(128) unsigned char smem[];
    const PushArgs& a = ta.a;
    ring_init(ta.ring, smem);
    __syncthreads();
    const Vec3 equ = dt_field_q(a.k, uniform_of(a.E)), bqu = dt_field_q(a.k, uniform_of(a.B));
    double acc[kRedVals] = {0.0, 0.0, 0.0, 0.0, 0.0};
    tma_ring_fields<PA, false, false>(
        a, ta.ring, smem,
        [&](const Vec3&, const Vec3&, Vec3& x, Vec3& p) { return push_steps_fast(a.k, a.nsteps, equ, bqu, x, p); },
        [&](int64_t i, double* xs, double* ps, int pitch) { push_one_exact_slot(a, i, xs, ps, pitch); },
        [&](const Vec3& p) {
            if (MOM) moments_add(&p, 1, a.k, ta.mom_mass, acc);
        });
    if (blockIdx.x == gridDim.x - 1) {                       // ragged tail, scalar path
        const Vec3 Eu = uniform_of(a.E), Bu = uniform_of(a.B);
        for (int64_t i = ta.ring.ntiles * kTileF + threadIdx.x; i < a.p.n; i += kTmaThreads) {
            Vec3 x, p;
            load_one(a.p, i, x, p);
            if (MOM) moments_add(&p, 1, a.k, ta.mom_mass, acc);
            if (push_one_fast(a, Eu, Bu, i, x, p)) store_one(a.p, i, x, p);
            else push_one_exact(a, i);
        }
    }
    if (MOM) block_sum<kTmaThreads>(acc, ta.mom_partials + static_cast<int64_t>(blockIdx.x) * 8);   // all 16 warps
}

// ---------------------------------------------------------------------------
// gather only (Interpolators.interpolate1D(x, y)(x_new))
// ---------------------------------------------------------------------------
struct InterpArgs {
    GatherArgs g;               // particle part unused
    const double* xq; int64_t xq_stride; int64_t n;
    double* out; int64_t out_stride_c;
    int ncomp;
    LinGrid lin;                // LIN kernels: the node coordinates as linspace parameters
    int ncpass;                 // interp_cells_nc_kernel: components per pass (as many planes as fit shared memory, <= 6)
};

// The view of six consecutive components c0 .. c0+5 of the y block.
__device__ __forceinline__ GridView interp_view(const InterpArgs& a, const GridView& g, int c0) {
    GridView gv = g;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const bool on = c0 + c < a.ncomp;
        gv.comp[c] = on ? g.comp[0] + static_cast<int64_t>(c0 + c) * a.g.stride[1] : nullptr;
        gv.stride[c] = g.stride[0];
    }
    return gv;
}

// Recovery path of one point: complete semantics, IEEE divisions, writes its own
// outputs.  It takes only the grid-constant block and the shared-memory base:
// handing it the hot loop's view or result array by reference would force them
// into local memory (that made the first version of this kernel 10x slower).
template <int FORMULA, bool STAGED>
__device__ __noinline__ void interp_point_exact(const InterpArgs& a, unsigned char* smem, int64_t i, int c0) {
    const GridView gv = interp_view(a, make_view<STAGED>(a.g, smem), c0);
    const double x = a.xq[i * a.xq_stride];
    double F[6];
    const unsigned st = gather6_exact<FORMULA>(gv, x, F);
    if (st != 0u && c0 == 0) flag_particle(a.g.status, i, st);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int c = 0; c < 6; ++c)
        if (c0 + c < a.ncomp) a.out[static_cast<int64_t>(c0 + c) * a.out_stride_c + i] = st ? nan : F[c];
}

// Four consecutive points per thread and iteration: two 128-bit loads of x, two
// 128-bit stores per component (the kernel moves only 8 B in + 8 B out per
// point and component, so it lives on bytes in flight).
// LIN: the abscissas are a tp_linspace grid -- the fast path rebuilds x[j], x[j+1] instead of loading them (the
// kernel is bound by its random shared-memory loads: four per point for a 1-D y, two with LIN).
template <int FORMULA, bool STAGED, int NC, bool LIN = false>   // NC = components per pass: 1 (1-D y, the common case) or 6
__global__ void __launch_bounds__(kThreads, NC == 1 ? 3 : 2) interp_kernel(const __grid_constant__ InterpArgs a, bool vec) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (STAGED) { stage_issue(a.g, smem); stage_wait(smem); }
    const GridView g = make_view<STAGED>(a.g, smem);
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kThreads;
    const int64_t nquads = (a.n + 3) >> 2;
    for (int c0 = 0; c0 < a.ncomp; c0 += NC) {               // NC components per pass
        GridView gv = interp_view(a, g, c0);
        if (NC == 1) {
#pragma unroll
            for (int c = 1; c < 6; ++c) gv.comp[c] = nullptr;     // lets the compiler drop the unused lanes
        }
        for (int64_t qd = gtid; qd < nquads; qd += gstride) {
            const int64_t i = qd * 4;
            const bool full = vec && i + 4 <= a.n;
            double x[4];
            if (full) {
                const double2 u = ld2(a.xq + i), v = ld2(a.xq + i + 2);
                x[0] = u.x; x[1] = u.y; x[2] = v.x; x[3] = v.y;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) x[q] = i + q < a.n ? a.xq[(i + q) * a.xq_stride] : 0.0;
            }
            double o[NC][4];
            unsigned bad = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double F[6];
                Guard gd;
                bool ok = (LIN ? gather6_fast_lin<FORMULA>(gv, a.lin, x[q], F, gd) : gather6_fast<FORMULA>(gv, x[q], F, gd)) &&
                          guard_ok(gd);
#pragma unroll
                for (int c = 0; c < NC; ++c) { ok = ok && (F[c] == F[c]); o[c][q] = F[c]; }     // NaN -> numpy's fallbacks
                if (!ok && i + q < a.n) bad |= 1u << q;
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c0 + c >= a.ncomp) continue;
                double* dst = a.out + static_cast<int64_t>(c0 + c) * a.out_stride_c + i;
                if (full) {
                    st2(dst, make_double2(o[c][0], o[c][1]));
                    st2(dst + 2, make_double2(o[c][2], o[c][3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (i + q < a.n) dst[q] = o[c][q];
                }
            }
            if (__builtin_expect(bad != 0u, 0))              // recovery overwrites the point's outputs
                for (int q = 0; q < 4; ++q)
                    if (bad & (1u << q)) interp_point_exact<FORMULA, STAGED>(a, smem, i + q, c0);
        }
    }
}

// Cell tables of the 1-D-y kernels below.  {y0, slope} per cell always; abscissas that are not a tp_linspace grid
// (interp1d takes ANY increasing x: stretched grids, hand-made node lists) also get {x0, x1} per cell and a BUCKET
// index: 2 (ng-1) uniform buckets over [x[0], x[-1]], bucket b -> the last cell whose left node lies in an earlier
// bucket.  A point then starts at its bucket's cell and walks right over the (usually 0-2) nodes of its own bucket
// -- no binary search, no dependence on x being uniform (the uniform-spacing guess sent every point of a
// stretched grid to the out-of-line exact path: 19 G points/s instead of 300).  bucket_of() is monotone in x, so
// a node in an earlier bucket is strictly left of the point whatever the rounding.
struct CellTables {
    double2* ys;        // {y[j], slope_j}
    double2* xs;        // {x[j], x[j+1]}           (not LIN)
    int* bucket;        // first cell to try        (not LIN)
    int ncell, nbucket;
    double inv_bw;
};

template <bool LIN>
__device__ __forceinline__ CellTables cell_tables_at(unsigned char* base, const InterpArgs& a) {
    CellTables t;
    t.ncell = a.g.ng - 1;
    t.nbucket = 2 * t.ncell;
    t.ys = reinterpret_cast<double2*>(base);
    t.xs = t.ys + t.ncell;
    t.bucket = reinterpret_cast<int*>(t.xs + t.ncell);
    t.inv_bw = a.g.inv_dr * 2.0;
    return t;
}

__device__ __forceinline__ int bucket_of(const CellTables& t, double x, double x_first) {
    const int b = __double2int_rz((x - x_first) * t.inv_bw);            // saturates, NaN -> 0
    return max(0, min(b, t.nbucket - 1));
}

// Called by all threads of the CTA; the caller syncs afterwards.
template <bool LIN>
__device__ __forceinline__ void cell_tables_build(const CellTables& t, const InterpArgs& a, int nthreads) {
    const double* y = a.g.comp[0];
    const int sn = a.g.stride[0];
    for (int j = threadIdx.x; j < t.ncell; j += nthreads) {
        const double x0 = LIN ? lin_r(a.lin, j) : a.g.r[j], x1 = LIN ? lin_r(a.lin, j + 1) : a.g.r[j + 1];
        const double y0 = y[static_cast<int64_t>(j) * sn], y1 = y[static_cast<int64_t>(j + 1) * sn];
        t.ys[j] = make_double2(y0, (y1 - y0) / (x1 - x0));
        if (!LIN) {
            t.xs[j] = make_double2(x0, x1);
            // buckets (gb(x[j]), gb(x[j+1])] start at cell j; the top cell takes everything up to the last bucket
            const int b0 = j == 0 ? -1 : bucket_of(t, x0, a.g.r_first);
            const int b1 = j == t.ncell - 1 ? t.nbucket - 1 : bucket_of(t, x1, a.g.r_first);
            for (int b = b0 + 1; b <= b1; ++b) t.bucket[b] = j;
        }
    }
}

// The cell to try for x and its two abscissas: the uniform-spacing guess on a tp_linspace grid, the bucket walk otherwise.
template <bool LIN>
__device__ __forceinline__ int cell_find(const CellTables& t, const InterpArgs& a, double xv, double& x0, double& x1) {
    int j;
    if (LIN) {
        j = guess_cell(xv, a.g.r_first, a.g.inv_dr, a.g.ng);
        const double dj = static_cast<double>(j);
        x0 = lin_r_at(a.lin, dj, 0, j); x1 = lin_r_at(a.lin, dj, 1, j + 1);
    } else {
        constexpr int kWalk = 6;
        j = t.bucket[bucket_of(t, xv, a.g.r_first)];
        double2 e = t.xs[j];
#pragma unroll 1
        for (int w = 0; w < kWalk && xv >= e.y && j < t.ncell - 1; ++w) e = t.xs[++j];
        x0 = e.x; x1 = e.y;
    }
    return j;
}

// y(x) by the per-cell table; false when x is not strictly inside the cell that was found (on a node, at the ends,
// outside, NaN, more than kWalk nodes inside one bucket) or the value is NaN: the caller takes the exact path.
template <bool LIN>
__device__ __forceinline__ bool cell_value(const CellTables& t, const InterpArgs& a, double xv, double& out) {
    double x0, x1;
    const int j = cell_find<LIN>(t, a, xv, x0, x1);
    const double2 tb = t.ys[j];
    out = tb.y * (xv - x0) + tb.x;
    return (x0 < xv) && (xv < x1) && (out == out);
}

// 1-D y, numpy.interp rounding (formula B), the common call (Interpolators.interpolate1D(grid.r, field)(x)):
// every CTA builds a per-CELL table {y[j], slope_j} in shared memory (slope = (y[j+1]-y[j])/(x[j+1]-x[j]) by IEEE
// division: numpy.interp's per-cell quantity, ng/256 divisions per thread, once per launch) and a point then costs
// ONE 128-bit shared load, a multiply and an add -- no reciprocal, no quotient, no guard, half the random shared
// loads of interp_kernel.  LIN: the abscissas are a tp_linspace grid, x[j] / x[j+1] are rebuilt in
// registers; otherwise the table also holds {x[j], x[j+1]} (two 128-bit loads per point).  Points that are not
// strictly inside the guessed cell (on a node, at the ends, outside, NaN, non-uniform x) take interp_point_exact.
template <bool LIN>
__global__ void __launch_bounds__(kThreads, 3) interp_cells_kernel(const __grid_constant__ InterpArgs a, bool vec) {
    extern __shared__ __align__(128) unsigned char smem[];
    const CellTables tb = cell_tables_at<LIN>(smem, a);
    cell_tables_build<LIN>(tb, a, kThreads);
    __syncthreads();
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kThreads;
    const int64_t nquads = (a.n + 3) >> 2;
    // kUnroll quads (4 points each) per thread and iteration, all their x loads issued before the first use: the
    // kernel moves only 16 B per point, so it needs ~45 KB of loads in flight per SM to cover the HBM latency
#ifndef TP_INTERP_UNROLL
#define TP_INTERP_UNROLL 2
#endif
    constexpr int kUnroll = TP_INTERP_UNROLL;
    for (int64_t qd0 = gtid; qd0 < nquads; qd0 += kUnroll * gstride) {
        double x[kUnroll][4];
        bool full[kUnroll], live[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t qd = qd0 + u * gstride;
            const int64_t i = qd * 4;
            live[u] = qd < nquads;
            full[u] = live[u] && vec && i + 4 <= a.n;
            if (full[u]) {
                const double2 p = ld2(a.xq + i), v = ld2(a.xq + i + 2);
                x[u][0] = p.x; x[u][1] = p.y; x[u][2] = v.x; x[u][3] = v.y;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) x[u][q] = (live[u] && i + q < a.n) ? a.xq[(i + q) * a.xq_stride] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!live[u]) continue;
            const int64_t i = (qd0 + u * gstride) * 4;
            double o[4];
            unsigned bad = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = cell_value<LIN>(tb, a, x[u][q], o[q]);          // false: numpy's special cases (exact path)
                if (!ok && i + q < a.n) bad |= 1u << q;
            }
            double* dst = a.out + i;
            if (full[u]) {
                st2(dst, make_double2(o[0], o[1]));
                st2(dst + 2, make_double2(o[2], o[3]));
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (i + q < a.n) dst[q] = o[q];
            }
            if (__builtin_expect(bad != 0u, 0))              // recovery overwrites the point's output (reads the global arrays)
                for (int q = 0; q < 4; ++q)
                    if (bad & (1u << q)) interp_point_exact<TP_INTERP_B, false>(a, smem, i + q, 0);
        }
    }
}

// N-D y, scipy interp1d._call_linear rounding (formula C, scipy 1.18: y_new = ((x - x_lo)/(x_hi - x_lo)) * y_hi +
// ((x_hi - x)/(x_hi - x_lo)) * y_lo -- two weights per POINT, shared by the components, so there is no per-cell slope
// to tabulate as for numpy.interp).  What can be tabulated is the node PAIR (second half): up to six
// components per pass get {y[j], y[j+1]} planes in shared memory (plane c at ys + c * ncell), so that a point reads one
// 16-byte-aligned 128-bit entry per component at the same cell index instead of two 64-bit node values (half the
// shared-memory instructions of interp_kernel<C, staged, 6>; the bytes are the same).  One 512-thread CTA per SM (the
// table of a 1025-node grid with six components is 96 KB), four consecutive points per thread, 128-bit loads of x and
// stores per component; the two quotients go through the guarded fast path like everywhere else.  Points that are
// not strictly inside the cell that was found (on a node -- searchsorted's left cell --, at the ends, outside, NaN)
// take interp_point_exact<C>, which writes all components of the pass itself.
constexpr int kCellsNcThreads = 512;

template <bool LIN>
__global__ void __launch_bounds__(kCellsNcThreads, 1) interp_cells_nc_kernel(const __grid_constant__ InterpArgs a, bool vec) {
    extern __shared__ __align__(128) unsigned char smem[];
    CellTables tb;
    tb.ncell = a.g.ng - 1;
    tb.nbucket = 2 * tb.ncell;
    tb.inv_bw = a.g.inv_dr * 2.0;
    const int ncmax = a.ncpass;
    tb.ys = reinterpret_cast<double2*>(smem);
    tb.xs = tb.ys + static_cast<size_t>(ncmax) * tb.ncell;
    tb.bucket = reinterpret_cast<int*>(tb.xs + tb.ncell);
    if (!LIN) {
        for (int j = threadIdx.x; j < tb.ncell; j += kCellsNcThreads) {
            const double x0 = a.g.r[j], x1 = a.g.r[j + 1];
            tb.xs[j] = make_double2(x0, x1);
            const int b0 = j == 0 ? -1 : bucket_of(tb, x0, a.g.r_first);
            const int b1 = j == tb.ncell - 1 ? tb.nbucket - 1 : bucket_of(tb, x1, a.g.r_first);
            for (int b = b0 + 1; b <= b1; ++b) tb.bucket[b] = j;
        }
    }
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kCellsNcThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kCellsNcThreads;
    const int64_t nquads = (a.n + 3) >> 2;
    const int sn = a.g.stride[0];
    const int64_t sc = a.g.stride[1];
    for (int c0 = 0; c0 < a.ncomp; c0 += ncmax) {
        const int nc = min(ncmax, a.ncomp - c0);
        __syncthreads();                                     // the previous pass has finished reading the planes
        for (int idx = threadIdx.x; idx < nc * tb.ncell; idx += kCellsNcThreads) {
            const int c = idx / tb.ncell, j = idx - c * tb.ncell;
            const double* y = a.g.comp[0] + static_cast<int64_t>(c0 + c) * sc;
            const double y0 = y[static_cast<int64_t>(j) * sn], y1 = y[static_cast<int64_t>(j + 1) * sn];
            tb.ys[idx] = make_double2(y0, y1);
        }
        __syncthreads();
        for (int64_t qd = gtid; qd < nquads; qd += gstride) {
            const int64_t i = qd * 4;
            const bool full = vec && i + 4 <= a.n;
            double x[4];
            if (full) {
                const double2 u = ld2(a.xq + i), v = ld2(a.xq + i + 2);
                x[0] = u.x; x[1] = u.y; x[2] = v.x; x[3] = v.y;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) x[q] = i + q < a.n ? a.xq[(i + q) * a.xq_stride] : 0.0;
            }
            double o[6][4];
            unsigned bad = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double x0, x1;
                const int j = cell_find<LIN>(tb, a, x[q], x0, x1);
                Guard gd;
                const Recip<true> rw = make_recip<true>(x1 - x0, gd);
                const double w_hi = quot<true>(x[q] - x0, rw, gd), w_lo = quot<true>(x1 - x[q], rw, gd);
                bool ok = (x0 < x[q]) && (x[q] < x1) && guard_ok(gd);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    if (c < nc) {
                        const double2 e = tb.ys[c * tb.ncell + j];           // {y_lo, y_hi}
                        o[c][q] = w_hi * e.y + w_lo * e.x;
                        ok = ok && (o[c][q] == o[c][q]);
                    }
                }
                if (!ok && i + q < a.n) bad |= 1u << q;
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                if (c >= nc) continue;
                double* dst = a.out + static_cast<int64_t>(c0 + c) * a.out_stride_c + i;
                if (full) {
                    st2(dst, make_double2(o[c][0], o[c][1]));
                    st2(dst + 2, make_double2(o[c][2], o[c][3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (i + q < a.n) dst[q] = o[c][q];
                }
            }
            if (__builtin_expect(bad != 0u, 0))              // recovery overwrites the point's outputs (reads the global arrays)
                for (int q = 0; q < 4; ++q)
                    if (bad & (1u << q)) interp_point_exact<TP_INTERP_C, false>(a, smem, i + q, c0);
        }
    }
}

// The same on the TMA ring (round 2): the points stream through shared memory like the particles of
// the push kernels -- one persistent CTA per SM, a producer lane that fills 15 KB stages of 1920 abscissas with
// cp.async.bulk and drains the results with bulk stores, 15 consumer warps that turn their four points into values IN
// PLACE (x -> y in the same slots).  No thread waits on a global load; the register-path kernel above needs ~48 KB of
// its own loads in flight per SM and reaches 70-78 % of the 16 B/point bound, this one is bound by the copy itself.
// Stage layout: thread t owns points {2t, 2t+1} and {960+2t, 960+2t+1} of the tile (two conflict-free LDS.128).
constexpr int kTileI = 4 * kConsumers;                            // points per tile
constexpr uint32_t kStageI = kTileI * sizeof(double);             // 15 KB

struct InterpRingArgs {
    InterpArgs a;
    int nstages;
    uint32_t table_off, tiles_off;
    int64_t ntiles;
};

// value of one point by the complete semantics (on a node, at the ends, outside the grid, NaN, non-uniform x)
__device__ __noinline__ double interp_point_exact_value(const InterpArgs& a, int64_t i) {
    GridView gv = interp_view(a, make_view<false>(a.g, nullptr), 0);
#pragma unroll
    for (int c = 1; c < 6; ++c) gv.comp[c] = nullptr;
    double F[6];
    const unsigned st = gather6_exact<TP_INTERP_B>(gv, a.xq[i], F);
    if (st != 0u) flag_particle(a.g.status, i, st);
    return st ? __longlong_as_double(0x7ff8000000000000ll) : F[0];
}

template <bool LIN>
__global__ void __launch_bounds__(kTmaThreads, 1) interp_cells_ring_kernel(const __grid_constant__ InterpRingArgs ra) {
    extern __shared__ __align__(128) unsigned char smem[];
    const InterpArgs& a = ra.a;
    uint64_t* full = ring_full(smem);
    uint64_t* done = ring_done(smem);
    unsigned char* tiles = smem + ra.tiles_off;
    const int nst = ra.nstages;
    const int64_t t0 = blockIdx.x, tstep = gridDim.x;
    if (threadIdx.x == 0)
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kConsumerWarps); }
    __syncthreads();
    if (threadIdx.x == kConsumers)                                   // first loads fly while the table is built
        for (int s = 0; s < nst; ++s) {
            const int64_t t = t0 + s * tstep;
            if (t >= ra.ntiles) break;
            mbar_expect_tx(&full[s], kStageI);
            tma_bulk_g2s(tiles + s * kStageI, a.xq + t * kTileI, kStageI, &full[s]);
        }
    const CellTables tb = cell_tables_at<LIN>(smem + ra.table_off, a);
    cell_tables_build<LIN>(tb, a, kTmaThreads);
    __syncthreads();
    int s = 0;
    uint32_t parity = 0;
    if (threadIdx.x >= kConsumers) {
        if (threadIdx.x == kConsumers) {                             // the producer lane
            for (int64_t t = t0; t < ra.ntiles; t += tstep) {
                mbar_wait(&done[s], parity);
                tma_bulk_s2g(a.out + t * kTileI, tiles + s * kStageI, kStageI);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                const int64_t tn = t + nst * tstep;
                if (tn < ra.ntiles) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_expect_tx(&full[s], kStageI);
                    tma_bulk_g2s(tiles + s * kStageI, a.xq + tn * kTileI, kStageI, &full[s]);
                }
                if (++s == nst) { s = 0; parity ^= 1u; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        for (int64_t t = t0; t < ra.ntiles; t += tstep) {
            mbar_wait(&full[s], parity);
            double* st = reinterpret_cast<double*>(tiles + s * kStageI);
            double* mine[2] = {st + 2 * threadIdx.x, st + 2 * kConsumers + 2 * threadIdx.x};
            const double2 u = ld2(mine[0]), v = ld2(mine[1]);
            const double x[4] = {u.x, u.y, v.x, v.y};
            double o[4];
            unsigned bad = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (!cell_value<LIN>(tb, a, x[q], o[q])) bad |= 1u << q;        // numpy's special cases: exact path
            if (__builtin_expect(bad != 0u, 0)) {
                const int64_t first = t * kTileI;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (bad & (1u << q))
                        o[q] = interp_point_exact_value(a, first + (q >> 1) * 2 * kConsumers + 2 * threadIdx.x + (q & 1));
            }
            st2(mine[0], make_double2(o[0], o[1]));
            st2(mine[1], make_double2(o[2], o[3]));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
            if (++s == nst) { s = 0; parity ^= 1u; }
        }
    }
    // ragged tail (< kTileI points) on the last CTA; it writes plain global stores to a range no bulk store touches
    if (blockIdx.x == gridDim.x - 1)
        for (int64_t i = ra.ntiles * kTileI + threadIdx.x; i < a.n; i += kTmaThreads) a.out[i] = interp_point_exact_value(a, i);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int fill_particles(const tp_particles* p, ParticleArgs& a, bool& vec2, bool* aos16 = nullptr) {
    if (!p || p->n < 0) return TP_E_ARG;
    if (p->n > 0 && (!p->pos || !p->mom)) return TP_E_ARG;
    if (p->pos_stride_n <= 0 || p->pos_stride_c <= 0 || p->mom_stride_n <= 0 || p->mom_stride_c <= 0)
        return TP_E_LAYOUT;
    a.pos = p->pos; a.pos_sn = p->pos_stride_n; a.pos_sc = p->pos_stride_c;
    a.mom = p->mom; a.mom_sn = p->mom_stride_n; a.mom_sc = p->mom_stride_c;
    a.n = p->n;
    vec2 = layout_of(a.pos_sn, a.pos_sc, a.n) == kLayoutSoA && layout_of(a.mom_sn, a.mom_sc, a.n) == kLayoutSoA &&
           aligned16(a.pos) && aligned16(a.mom) && (a.pos_sc % 2 == 0) && (a.mom_sc % 2 == 0) && a.n >= 2;
    if (aos16)
        *aos16 = layout_of(a.pos_sn, a.pos_sc, a.n) == kLayoutAoS && layout_of(a.mom_sn, a.mom_sc, a.n) == kLayoutAoS &&
                 aligned16(a.pos) && aligned16(a.mom);
    return TP_OK;
}

static int fill_consts(const tp_push_consts* k, PushK& o) {
    if (!k) return TP_E_ARG;
    o.dt = k->dt; o.q = k->charge; o.mm = k->mass_sq;
    o.c2 = k->c2;
    o.rc2 = 1.0 / k->c2;                        // IEEE on the host: RN(1/c2)
    o.fast_ok = const_divisor_ok(k->c2) ? 1 : 0;
    return TP_OK;
}

static int grid_blocks(const void* kernel, size_t smem, int64_t work_items) {
    const DeviceProps& d = device_props();
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int64_t full = static_cast<int64_t>(d.sm_count) * per_sm;     // one resident wave, persistent loop
    const int64_t need = (work_items + kThreads - 1) / kThreads;
    return static_cast<int>(std::max<int64_t>(1, std::min(full, need)));
}

static int fill_field(const double* ptr, int mode, int64_t sn, int64_t sc, FieldArg& f) {
    f.ptr = ptr; f.mode = mode; f.sn = sn; f.sc = sc; f.u[0] = f.u[1] = f.u[2] = 0.0;
    if (!ptr) return TP_E_ARG;
    if (mode == TP_FIELD_UNIFORM_HOST) { f.u[0] = ptr[0]; f.u[1] = ptr[1]; f.u[2] = ptr[2]; f.ptr = nullptr; }
    else if (mode == TP_FIELD_PER_PARTICLE) { if (sn <= 0 || sc <= 0) return TP_E_LAYOUT; }
    else if (mode != TP_FIELD_UNIFORM_DEV) return TP_E_MODE;
    return TP_OK;
}

// Build the global view and (if it fits) the shared-memory staging plan.
static int fill_grid(const tp_grid_fields* g, GatherArgs& a, size_t& smem_bytes, bool& staged,
                     int64_t comp0_extent = 0) {
    if (!g || !g->r) return TP_E_ARG;
    if (g->ng < 2 || g->ng > (1ll << 31) - 8 || !(g->dr > 0.0)) return TP_E_GRID;
    a.r = g->r; a.ng = static_cast<int>(g->ng); a.dr = g->dr; a.inv_dr = 1.0 / g->dr;
    a.r_first = g->r_first; a.r_last = g->r_last;
    struct Range { uintptr_t lo, hi; };
    Range rng[kMaxSeg]; int nr = 0;
    rng[nr++] = {reinterpret_cast<uintptr_t>(g->r), reinterpret_cast<uintptr_t>(g->r + g->ng)};
    bool stride_ok = true;
    for (int c = 0; c < 6; ++c) {
        a.comp[c] = g->comp[c];
        a.stride[c] = 1;
        if (!g->comp[c]) continue;
        if (g->comp_stride[c] <= 0 || g->comp_stride[c] > (1 << 20)) return TP_E_LAYOUT;
        a.stride[c] = static_cast<int>(g->comp_stride[c]);
        const uintptr_t lo = reinterpret_cast<uintptr_t>(g->comp[c]);
        const int64_t extent = (c == 0 && comp0_extent > 0) ? comp0_extent : (g->ng - 1) * g->comp_stride[c] + 1;
        rng[nr++] = {lo, lo + extent * sizeof(double)};
        if (g->comp_stride[c] > 8) stride_ok = false;       // too sparse to be worth staging
    }
    // merge overlapping / touching ranges (columns of one interleaved (ng,3) array)
    std::sort(rng, rng + nr, [](const Range& x, const Range& y) { return x.lo < y.lo; });
    Range seg[kMaxSeg]; int ns = 0;
    for (int i = 0; i < nr; ++i) {
        if (ns && rng[i].lo <= seg[ns - 1].hi) seg[ns - 1].hi = std::max(seg[ns - 1].hi, rng[i].hi);
        else seg[ns++] = rng[i];
    }
    uint32_t off = kSmemHeader;
    uint64_t total = kSmemHeader;
    a.nseg = ns;
    for (int s = 0; s < ns; ++s) {
        const uint64_t bytes = seg[s].hi - seg[s].lo;
        total += (bytes + 15) & ~15ull;
        if (total > (1u << 20)) { total = ~0ull >> 1; break; }
        a.seg[s].src = reinterpret_cast<const double*>(seg[s].lo);
        a.seg[s].smem_off = off;
        a.seg[s].n_doubles = static_cast<uint32_t>(bytes / sizeof(double));
        a.seg[s].tma_bytes = (seg[s].lo & 15u) ? 0u : static_cast<uint32_t>(bytes & ~15ull);
        off += static_cast<uint32_t>((bytes + 15) & ~15ull);
    }
    const DeviceProps& d = device_props();
    staged = stride_ok && total <= static_cast<uint64_t>(d.smem_optin) - 1024;
    smem_bytes = staged ? static_cast<size_t>(total) : 0;
    if (staged) {
        auto smem_of = [&](const double* ptr) -> uint32_t {
            const uintptr_t u = reinterpret_cast<uintptr_t>(ptr);
            for (int s = 0; s < ns; ++s)
                if (u >= seg[s].lo && u < seg[s].hi) return a.seg[s].smem_off + static_cast<uint32_t>(u - seg[s].lo);
            return 0;
        };
        a.r_smem_off = smem_of(g->r);
        for (int c = 0; c < 6; ++c) a.comp_smem_off[c] = g->comp[c] ? smem_of(g->comp[c]) : 0;
    } else {
        a.nseg = 0;
    }
    return TP_OK;
}

template <typename K>
static int set_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024)
        TP_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    return TP_OK;
}

template <bool VEC2, int FORMULA, bool STAGED>
static int launch_gather_push(const GatherArgs& a, size_t smem, cudaStream_t s) {
    auto kern = gather_push_kernel<VEC2, FORMULA, STAGED>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    const int64_t items = VEC2 ? (a.p.n + 1) / 2 : a.p.n;
    const int blocks = grid_blocks(reinterpret_cast<const void*>(kern), smem, items);
    kern<<<blocks, kThreads, smem, s>>>(a);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

template <bool VEC2, int FORMULA>
static int launch_gather_push_s(const GatherArgs& a, size_t smem, bool staged, cudaStream_t s) {
    return staged ? launch_gather_push<VEC2, FORMULA, true>(a, smem, s)
                  : launch_gather_push<VEC2, FORMULA, false>(a, 0, s);
}

template <bool VEC2>
static int launch_gather_push_f(const GatherArgs& a, int formula, size_t smem, bool staged, cudaStream_t s) {
    switch (formula) {
        case TP_INTERP_A: return launch_gather_push_s<VEC2, TP_INTERP_A>(a, smem, staged, s);
        case TP_INTERP_B: return launch_gather_push_s<VEC2, TP_INTERP_B>(a, smem, staged, s);
        case TP_INTERP_C: return launch_gather_push_s<VEC2, TP_INTERP_C>(a, smem, staged, s);
        default: return TP_E_MODE;
    }
}

// TMA-pipelined variant: SoA or AoS, 16-byte aligned arrays, at least one full tile per SM.
template <int FORMULA, int GMODE, bool AOS, unsigned PMASK = 0u>
static int launch_gather_push_tma(const TmaArgs& ta, size_t smem, int blocks, cudaStream_t s) {
    auto kern = gather_push_tma_kernel<FORMULA, GMODE, AOS, PMASK>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    kern<<<blocks, kTmaThreads, smem, s>>>(ta);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

static bool mask_kernels_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TURBOPY_B200_MASK_KERNELS");     // tuning / A-B switch, default on
        return !(e && e[0] == '0');
    }();
    return on;
}

template <int FORMULA, int GMODE, bool AOS, unsigned PMASK = 0u>
static int launch_gather_push_tma_mom(const TmaArgs& ta, size_t smem, int blocks, cudaStream_t s) {
    auto kern = gather_push_tma_kernel<FORMULA, GMODE, AOS, PMASK, true>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    kern<<<blocks, kTmaThreads, smem, s>>>(ta);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

// `mom` (staged grid, formula B, SoA -- the caller checks): the instances that also accumulate the moment sums.
template <int FORMULA>
static int launch_gather_push_tma_v(const TmaArgs& ta, size_t smem, int blocks, int gmode, bool aos, cudaStream_t s,
                                    bool mom = false) {
    if (gmode == 1 && mask_kernels_enabled()) {              // staged grid: the specialised shapes
        unsigned mask = 0;
        for (int k = 0; k < 6; ++k) mask |= ta.g.comp[k] ? (1u << k) : 0u;
        // Pacing.  The specialised consumers finish a tile in about half the time the memory system needs to deliver
        // the next one; releasing the stage that early keeps all three stages of all 148 rings in flight at once,
        // which over-subscribes HBM (the same effect as a fourth stage: measured 262 us per launch instead of 254).
        // They therefore sleep ~0.8 us per tile before touching it (no issue slots, no power): 254 us per launch,
        // 97 % of the HBM peak over 8000 steps (profiles/r2_window_kernel.md).
        static const int pace = [] { const char* e = std::getenv("TURBOPY_B200_CONSUMER_DELAY_NS"); return e ? std::atoi(e) : 800; }();
        TmaArgs tb = ta;
        tb.ring.consumer_delay_ns = pace;
        if (FORMULA == TP_INTERP_B && mom && mask == 0x02u)
            return launch_gather_push_tma_mom<TP_INTERP_B, 1, false, 0x02u>(tb, smem, blocks, s);
        if (mask == 0x02u)
            return aos ? launch_gather_push_tma<FORMULA, 1, true, 0x02u>(tb, smem, blocks, s)
                       : launch_gather_push_tma<FORMULA, 1, false, 0x02u>(tb, smem, blocks, s);
        if (mask == 0x07u && !mom)
            return aos ? launch_gather_push_tma<FORMULA, 1, true, 0x07u>(tb, smem, blocks, s)
                       : launch_gather_push_tma<FORMULA, 1, false, 0x07u>(tb, smem, blocks, s);
    }
    if (FORMULA == TP_INTERP_B && mom) return launch_gather_push_tma_mom<TP_INTERP_B, 1, false>(ta, smem, blocks, s);
    switch (gmode) {
        case 1: return aos ? launch_gather_push_tma<FORMULA, 1, true>(ta, smem, blocks, s)
                           : launch_gather_push_tma<FORMULA, 1, false>(ta, smem, blocks, s);
        case 2: return aos ? launch_gather_push_tma<FORMULA, 2, true>(ta, smem, blocks, s)
                           : launch_gather_push_tma<FORMULA, 2, false>(ta, smem, blocks, s);
        default: return aos ? launch_gather_push_tma<FORMULA, 0, true>(ta, smem, blocks, s)
                            : launch_gather_push_tma<FORMULA, 0, false>(ta, smem, blocks, s);
    }
}

// Library-owned scratch for the packed node records when the caller passes none (tp_gather_push_f64 on a grid
// too large for shared memory; tp_gather_push_windows_f64 takes caller-owned scratch, which is what the Python
// tool uses): one buffer per stream (calls on one
// stream are ordered, so the buffer is free again when the next call runs).
// Allocation is synchronous: a call that would have to (re)allocate while its
// stream is being captured into a graph fails with TP_E_GRAPH -- run one eager
// call first.  A buffer that is outgrown or whose slot is recycled is RETIRED,
// not freed: a CUDA graph captured earlier may still point at it (64 B per grid
// node; released with the process).
static int records_scratch(int64_t ng, cudaStream_t s, double** out) {
    struct Slot { cudaStream_t stream; double* buf; int64_t cap; int device; };
    static Slot slots[16];
    static int used = 0;
    static std::mutex guard;                                // callers may come from several host threads
    std::lock_guard<std::mutex> lock(guard);
    int dev = 0;
    TP_CUDA_TRY(cudaGetDevice(&dev));
    Slot* hit = nullptr;
    for (int i = 0; i < used; ++i)
        if (slots[i].stream == s && slots[i].device == dev) hit = &slots[i];
    if (hit && hit->cap >= ng) { *out = hit->buf; return TP_OK; }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return TP_E_GRAPH;
    if (!hit) {
        static int next = 0;
        if (used < 16) hit = &slots[used++];
        else {                                              // recycle the oldest slot (its stream is most likely gone)
            hit = &slots[next];
            next = (next + 1) % 16;
            hit->buf = nullptr;                             // retired (see above)
        }
        hit->stream = s; hit->buf = nullptr; hit->cap = 0; hit->device = dev;
    }
    const int64_t grown = std::max<int64_t>(std::max<int64_t>(ng, 1024), 2 * hit->cap);   // geometric: few retirements
    hit->buf = nullptr; hit->cap = 0;                       // an outgrown buffer is retired (see above)
    TP_CUDA_TRY(cudaMalloc(&hit->buf, static_cast<size_t>(grown) * 64));
    hit->cap = grown;
    *out = hit->buf;
    return TP_OK;
}

static bool records_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TURBOPY_B200_RECORDS");      // tuning / A-B switch, default on
        return !(e && e[0] == '0');
    }();
    return on;
}

static bool l2_hints_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TURBOPY_B200_L2HINT");       // tuning / A-B switch, default on
        return !(e && e[0] == '0');
    }();
    return on;
}

static bool tma_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TURBOPY_B200_TMA");          // tuning / A-B switch, default on
        return !(e && e[0] == '0');
    }();
    return on;
}

// Ring geometry for `n` particles behind `front` bytes of other shared memory; false if the ring is not worth it.
static bool plan_ring(int64_t n, size_t front, RingArgs& ring, size_t& smem, int& blocks, int tile = kTile) {
    const DeviceProps& d = device_props();
    ring.ntiles = n / tile;
    ring.stream_hint = 0;
    ring.consumer_delay_ns = 0;
    if (!tma_enabled() || ring.ntiles < d.sm_count) return false;
    ring.tiles_off = static_cast<uint32_t>((front + 127) & ~size_t(127));
    const int64_t room = d.smem_optin - 1024 - static_cast<int64_t>(ring.tiles_off);
    ring.nstages = static_cast<int>(std::min<int64_t>(kMaxStages, room / kTileBytes));
    if (const char* cap = std::getenv("TURBOPY_B200_STAGES"))       // tuning knob: ring depth (2..4, as far as it fits)
        ring.nstages = static_cast<int>(std::min<int64_t>(room / kTileBytes, std::min(4, std::max(2, std::atoi(cap)))));
    if (ring.nstages < 2) return false;
    smem = ring.tiles_off + static_cast<size_t>(ring.nstages) * kTileBytes;
    blocks = static_cast<int>(std::min<int64_t>(d.sm_count, ring.ntiles));
    return true;
}

// kind: 0 = packed 64-byte node records (the L2 kernel), 1 = 80-byte-pitch node records (windowed, formulas A / C),
// 2 = 144-byte-pitch cell records (windowed, formula B)
static int pack_records(const GatherArgs& a, cudaStream_t s, double* caller_scratch, double** rec, int kind = 0) {
    int rc;
    if (caller_scratch) *rec = caller_scratch;               // TP_RECORDS_DOUBLES_PER_NODE * ng doubles owned by the caller
    else if ((rc = records_scratch(kind == 0 ? a.ng : 3 * static_cast<int64_t>(a.ng), s, rec))) return rc;
    const DeviceProps& d = device_props();
    const int pblocks = static_cast<int>(std::min<int64_t>((a.ng + 255) / 256, static_cast<int64_t>(d.sm_count) * 8));
    if (kind == 2) pack_cells_kernel<<<pblocks, 256, 0, s>>>(a, *rec);               // (36 KB of static shared memory)
    else if (kind == 1) pack_nodes_kernel<kNodePitchWin><<<pblocks, 256, 0, s>>>(a, *rec);
    else pack_nodes_kernel<8><<<pblocks, 256, 0, s>>>(a, *rec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

// A moment diagnostic riding on a gather+push call: {KE, Px, Py, Pz, |p|^2, n, 0, 0} of the momenta BEFORE the push,
// i.e. what tp_reduce_moments_f64 called just before the push would return (to summation-order rounding).
struct MomRequest {
    double mass, mass_sq, c2;
    double* out8;
    double* scratch;            // tp_reduce_scratch_doubles() doubles
};

static int moments_standalone(const ParticleArgs& p, const MomRequest& mr, cudaStream_t s) {
    return tp_reduce_moments_f64(p.mom, p.mom_sn, p.mom_sc, p.n, mr.mass, mr.mass_sq, mr.c2, mr.out8, mr.scratch, s);
}
static int moments_standalone(const GatherArgs& a, const MomRequest& mr, cudaStream_t s) {
    return moments_standalone(a.p, mr, s);
}

static bool fused_moments_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TURBOPY_B200_FUSED_MOMENTS");      // A-B switch, default on
        return !(e && e[0] == '0');
    }();
    return on;
}

// Returns TP_OK and sets `launched` when the TMA kernel took the call.
static int try_gather_push_tma(const GatherArgs& a, int formula, size_t grid_smem, bool staged, bool aos,
                               double* rec_scratch, const MomRequest* mr, cudaStream_t s, bool& launched) {
    launched = false;
    TmaArgs ta;
    ta.g = a;
    ta.mom = MomConsts{};
    ta.mom_partials = nullptr;
    size_t smem = 0; int blocks = 0;
    if (!plan_ring(a.p.n, staged ? grid_smem : kSmemHeader, ta.ring, smem, blocks)) return TP_OK;
    int rc;
    // the moment diagnostic rides along on the staged formula-B SoA instances; otherwise the standalone reduction first
    const bool mom = mr && staged && formula == TP_INTERP_B && !aos && blocks <= kRedMaxBlocks && fused_moments_enabled() &&
                     mr->mass_sq == a.k.mm && mr->c2 == a.k.c2;
    if (mr && !mom && (rc = moments_standalone(a, *mr, s))) return rc;
    if (mom) {
        ta.mom = make_mom_consts(mr->mass, mr->mass_sq, mr->c2);
        ta.mom_partials = mr->scratch;
        rc = launch_gather_push_tma_v<TP_INTERP_B>(ta, smem, blocks, 1, false, s, true);
        if (rc == TP_OK) rc = launch_moments_final(mr->scratch, blocks, a.p.n, mr->out8, s);
        launched = (rc == TP_OK);
        return rc;
    }
    int gmode = staged ? 1 : 0;
    if (!staged && records_enabled()) {
        // the grid lives in L2: pack {r, E, B} into 64-byte node records first (ng * 64 B, a few us)
        double* rec = nullptr;
        if ((rc = pack_records(a, s, rec_scratch, &rec))) return rc;
        ta.g.records = rec;
        gmode = 2;
        ta.ring.stream_hint = l2_hints_enabled() ? 1 : 0;
    }
    switch (formula) {
        case TP_INTERP_A: rc = launch_gather_push_tma_v<TP_INTERP_A>(ta, smem, blocks, gmode, aos, s); break;
        case TP_INTERP_B: rc = launch_gather_push_tma_v<TP_INTERP_B>(ta, smem, blocks, gmode, aos, s); break;
        case TP_INTERP_C: rc = launch_gather_push_tma_v<TP_INTERP_C>(ta, smem, blocks, gmode, aos, s); break;
        default: return TP_E_MODE;
    }
    launched = (rc == TP_OK);
    return rc;
}

static bool windows_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TURBOPY_B200_WINDOWS");      // tuning / A-B switch, default on
        return !(e && e[0] == '0');
    }();
    return on;
}

// Ring + per-stage record windows for `n` particles; false if the windowed kernel is not applicable.
static bool plan_window_ring(int64_t n, RingArgs& ring, WindowArgs& win, size_t& smem, int& blocks) {
    const DeviceProps& d = device_props();
    if (!windows_enabled() || !records_enabled()) return false;
    if (!plan_ring(n, kWinFront, ring, smem, blocks)) return false;
    // Default: what is left of the 164 KB shared-memory carve-out the 3-stage ring needs anyway (the next
    // configurable sizes are 196 and 228 KB, and every step up takes 32 KB from the L1 that the random-order
    // fallback lives on: measured 31 -> 19 G pushes/s in random order with 228 KB).  ~9 KB per stage = 148 node
    // records = 145 cells per tile; a cell-ordered 960-particle tile of config 3 spans 14-67 cells.
    const int64_t carve = std::min<int64_t>(164 * 1024, d.smem_optin);
    const int64_t left = carve - 1024 - static_cast<int64_t>(smem);
    int64_t per_stage = (left / ring.nstages) & ~int64_t(127);
    if (const char* cap = std::getenv("TURBOPY_B200_WINDOW_BYTES")) {          // tuning knob
        const int64_t most = ((d.smem_optin - 1024 - static_cast<int64_t>(smem)) / ring.nstages) & ~int64_t(127);
        per_stage = std::min<int64_t>(most, std::max(0, std::atoi(cap)) & ~127);
    }
    if (per_stage < 16 * kCellBytes) return false;
    win.win_off = static_cast<uint32_t>(smem);                 // tiles_off and kTileBytes are multiples of 128
    win.win_bytes = static_cast<uint32_t>(per_stage);
    smem += static_cast<size_t>(ring.nstages) * per_stage;
    return true;
}

template <int FORMULA>
static int launch_gather_push_win(const TmaWinArgs& ta, size_t smem, int blocks, bool aos, cudaStream_t s) {
    auto go = [&](auto kern) {
        int rc = set_smem(kern, smem);
        if (rc) return rc;
        kern<<<blocks, kTmaThreads, smem, s>>>(ta);
        count_launch();
        return static_cast<int>(cudaGetLastError());
    };
    return aos ? go(gather_push_win_kernel<FORMULA, true>) : go(gather_push_win_kernel<FORMULA, false>);
}

// Returns TP_OK and sets `launched` when the windowed kernel took the call.  `mr` (optional): when the call is
// taken, the moments are delivered too -- inside the kernel for the formula-B SoA instance, by the standalone
// reduction in front of it otherwise.
static int try_gather_push_windows(const GatherArgs& a, int formula, bool aos, int32_t* bounds, int64_t bounds_len,
                                   double* rec_scratch, const MomRequest* mr, cudaStream_t s, bool& launched) {
    launched = false;
    if (!bounds) return TP_OK;
    TmaWinArgs ta;
    ta.g = a;
    ta.mom = MomConsts{};
    ta.mom_partials = nullptr;
    size_t smem = 0; int blocks = 0;
    if (!plan_window_ring(a.p.n, ta.ring, ta.win, smem, blocks)) return TP_OK;
    if (bounds_len < 2 * ta.ring.ntiles) return TP_E_ARG;
    ta.win.bounds = reinterpret_cast<int2*>(bounds);
    int rc;
    double* rec = nullptr;
    if ((rc = pack_records(a, s, rec_scratch, &rec, formula == TP_INTERP_B ? 2 : 1))) return rc;
    ta.g.records = rec;
    ta.ring.stream_hint = l2_hints_enabled() ? 1 : 0;
    if (mr && formula == TP_INTERP_B && !aos && blocks <= kRedMaxBlocks && fused_moments_enabled() &&
        mr->mass_sq == a.k.mm && mr->c2 == a.k.c2) {       // (the kernel takes mass**2 and c2 from the push's constants)
        ta.mom = make_mom_consts(mr->mass, mr->mass_sq, mr->c2);
        ta.mom_partials = mr->scratch;
        auto kern = gather_push_win_kernel<TP_INTERP_B, false, true>;
        if ((rc = set_smem(kern, smem))) return rc;
        kern<<<blocks, kTmaThreads, smem, s>>>(ta);
        count_launch();
        if ((rc = static_cast<int>(cudaGetLastError()))) return rc;
        rc = launch_moments_final(mr->scratch, blocks, a.p.n, mr->out8, s);
        launched = (rc == TP_OK);
        return rc;
    }
    if (mr && (rc = moments_standalone(a, *mr, s))) return rc;
    switch (formula) {
        case TP_INTERP_A: rc = launch_gather_push_win<TP_INTERP_A>(ta, smem, blocks, aos, s); break;
        case TP_INTERP_B: rc = launch_gather_push_win<TP_INTERP_B>(ta, smem, blocks, aos, s); break;
        case TP_INTERP_C: rc = launch_gather_push_win<TP_INTERP_C>(ta, smem, blocks, aos, s); break;
        default: return TP_E_MODE;
    }
    launched = (rc == TP_OK);
    return rc;
}

template <int FORMULA, bool STAGED, int NC, bool LIN = false>
static int launch_interp_nc(const InterpArgs& a, size_t smem, cudaStream_t s);

template <int FORMULA, bool STAGED, bool LIN = false>
static int launch_interp(const InterpArgs& a, size_t smem, cudaStream_t s) {
    return a.ncomp == 1 ? launch_interp_nc<FORMULA, STAGED, 1, LIN>(a, smem, s)
                        : launch_interp_nc<FORMULA, STAGED, 6, LIN>(a, smem, s);
}

template <int FORMULA, bool STAGED, int NC, bool LIN>
static int launch_interp_nc(const InterpArgs& a, size_t smem, cudaStream_t s) {
    auto kern = interp_kernel<FORMULA, STAGED, NC, LIN>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    const int blocks = grid_blocks(reinterpret_cast<const void*>(kern), smem, (a.n + 3) / 4);
    const bool vec = a.xq_stride == 1 && aligned16(a.xq) && aligned16(a.out) && a.out_stride_c % 2 == 0;
    kern<<<blocks, kThreads, smem, s>>>(a, vec);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

// Entry point shared with host_path.cu
int gather_push_device(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g, int axis,
                       int formula, uint64_t* status, cudaStream_t s, int nsteps, int32_t* bounds, int64_t bounds_len,
                       double* rec_scratch, const MomRequest* mr) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (nsteps < 0) return TP_E_ARG;
    if (nsteps == 0) return TP_OK;
    GatherArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nsteps = nsteps;
    bool vec2 = false, aos16 = false;
    int rc = fill_particles(p, a.p, vec2, &aos16);
    if (rc) return rc;
    if ((rc = fill_consts(k, a.k))) return rc;
    if (axis < 0 || axis > 2) return TP_E_MODE;
    if (!status) return TP_E_ARG;
    size_t smem = 0; bool staged = false;
    if ((rc = fill_grid(g, a, smem, staged))) return rc;
    a.axis = axis;
    a.status = reinterpret_cast<unsigned long long*>(status);
    if (a.p.n == 0) return TP_OK;
    if (formula < TP_INTERP_A || formula > TP_INTERP_C) return TP_E_MODE;
    // Multi-step passes are fp64-bound, not HBM-bound: they run on the register-path kernel (whose step loop
    // sits around the pair of particles) and leave the single-step TMA kernel free of a runtime loop, which
    // costs it 3 % (measured: 66.6 -> 64.2 G pushes/s on the headline configuration).
    if ((vec2 || aos16) && a.nsteps == 1) {
        bool launched = false;
        if (!staged) {          // grid too large for shared memory: per-tile record windows (cell-ordered particles)
            if ((rc = try_gather_push_windows(a, formula, aos16, bounds, bounds_len, rec_scratch, mr, s, launched))) return rc;
            if (launched) return TP_OK;
        }
        if ((rc = try_gather_push_tma(a, formula, smem, staged, aos16, rec_scratch, mr, s, launched))) return rc;
        if (launched) return TP_OK;
    }
    if (mr && (rc = moments_standalone(a, *mr, s))) return rc;
    return vec2 ? launch_gather_push_f<true>(a, formula, smem, staged, s)
                : launch_gather_push_f<false>(a, formula, smem, staged, s);
}

}  // namespace tp

using namespace tp;

static int boris_push_device(const tp_particles* p, const tp_push_consts* k, const double* E, int e_mode,
                             int64_t e_sn, int64_t e_sc, const double* B, int b_mode, int64_t b_sn,
                             int64_t b_sc, int nsteps, tp_stream_t stream, const MomRequest* mr = nullptr) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (nsteps < 0) return TP_E_ARG;
    if (nsteps == 0) return TP_OK;
    PushArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nsteps = nsteps;
    bool vec2 = false, aos16 = false;
    int rc = fill_particles(p, a.p, vec2, &aos16);
    if (rc) return rc;
    if ((rc = fill_consts(k, a.k))) return rc;
    if ((rc = fill_field(E, e_mode, e_sn, e_sc, a.E))) return rc;
    if ((rc = fill_field(B, b_mode, b_sn, b_sc, a.B))) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // a moment diagnostic riding on this push (tp_boris_push_moments_f64): delivered by the fused instance of the
    // uniform-field ring kernel (SoA, one step) or by the standalone reduction in front of whatever kernel runs
    auto standalone = [&]() {
        if (!mr) return static_cast<int>(TP_OK);
        const MomRequest* r = mr;
        mr = nullptr;
        return moments_standalone(a.p, *r, s);
    };
    if (a.p.n == 0) return standalone();
    const bool uniform = a.E.mode != TP_FIELD_PER_PARTICLE && a.B.mode != TP_FIELD_PER_PARTICLE;
    if ((vec2 || aos16) && uniform) {
        PushTmaArgs ta;
        ta.a = a;
        ta.mom_mass = 0.0;
        ta.mom_partials = nullptr;
        size_t smem = 0; int blocks = 0;
        if (plan_ring(a.p.n, kSmemHeader, ta.ring, smem, blocks, kTileF)) {
            const bool mom = mr && !aos16 && nsteps == 1 && blocks <= kRedMaxBlocks && fused_moments_enabled() &&
                             mr->mass_sq == a.k.mm && mr->c2 == a.k.c2;
            if (!mom && (rc = standalone())) return rc;
            int nst = kPushStages;
            if (const char* cap = std::getenv("TURBOPY_B200_STAGES")) nst = std::min(kRingMaxStages, std::max(2, std::atoi(cap)));
            ta.ring.nstages = nst;
            static const int push_pace = [] { const char* e = std::getenv("TURBOPY_B200_PUSH_DELAY_NS"); return e ? std::atoi(e) : 0; }();
            ta.ring.consumer_delay_ns = push_pace;
            smem = ta.ring.tiles_off + static_cast<size_t>(nst) * 2 * kArrayF;
            if (mom) {
                ta.mom_mass = mr->mass;
                ta.mom_partials = mr->scratch;
                if ((rc = set_smem(push_tma_kernel<false, true>, smem))) return rc;
                push_tma_kernel<false, true><<<blocks, kTmaThreads, smem, s>>>(ta);
                count_launch();
                if ((rc = static_cast<int>(cudaGetLastError()))) return rc;
                return launch_moments_final(mr->scratch, blocks, a.p.n, mr->out8, s);
            }
            if (aos16) {
                if ((rc = set_smem(push_tma_kernel<true>, smem))) return rc;
                push_tma_kernel<true><<<blocks, kTmaThreads, smem, s>>>(ta);
            } else {
                if ((rc = set_smem(push_tma_kernel<false>, smem))) return rc;
                push_tma_kernel<false><<<blocks, kTmaThreads, smem, s>>>(ta);
            }
            count_launch();
            return static_cast<int>(cudaGetLastError());
        }
    }
    if ((rc = standalone())) return rc;
    if ((vec2 || aos16) && a.E.mode == TP_FIELD_PER_PARTICLE && a.B.mode == TP_FIELD_PER_PARTICLE) {
        // both fields in one bulk-copyable layout: three component slices, or one C-order (n,3) block
        auto bulk_layout = [&](const FieldArg& f) {
            if (!aligned16(f.ptr)) return -1;
            const int l = layout_of(f.sn, f.sc, a.p.n);
            return l == kLayoutSoA ? (f.sc % 2 == 0 ? 0 : -1) : l == kLayoutAoS ? 1 : -1;
        };
        const int le = bulk_layout(a.E), lb = bulk_layout(a.B);
        PushTmaArgs ta;
        ta.a = a;
        size_t smem = 0; int blocks = 0;
        if (le >= 0 && le == lb && plan_ring(a.p.n, kSmemHeader, ta.ring, smem, blocks, kTileF)) {
            auto launch = [&](auto kern) {
                int r = set_smem(kern, smem);
                if (r) return r;
                kern<<<blocks, kTmaThreads, smem, s>>>(ta);
                count_launch();
                return static_cast<int>(cudaGetLastError());
            };
            if (aos16) return le ? launch(push_fields_tma_kernel<true, true>) : launch(push_fields_tma_kernel<true, false>);
            return le ? launch(push_fields_tma_kernel<false, true>) : launch(push_fields_tma_kernel<false, false>);
        }
    }
    if (vec2) {
        const int blocks = grid_blocks(reinterpret_cast<const void*>(push_kernel<true>), 0, (a.p.n + 1) / 2);
        push_kernel<true><<<blocks, kThreads, 0, s>>>(a);
    } else {
        const int blocks = grid_blocks(reinterpret_cast<const void*>(push_kernel<false>), 0, a.p.n);
        push_kernel<false><<<blocks, kThreads, 0, s>>>(a);
    }
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_boris_push_f64(const tp_particles* p, const tp_push_consts* k, const double* E, int e_mode,
                                 int64_t e_sn, int64_t e_sc, const double* B, int b_mode, int64_t b_sn,
                                 int64_t b_sc, tp_stream_t stream) {
    return boris_push_device(p, k, E, e_mode, e_sn, e_sc, B, b_mode, b_sn, b_sc, 1, stream);
}

extern "C" int tp_boris_push_moments_f64(const tp_particles* p, const tp_push_consts* k, const double* E, int e_mode,
                                         int64_t e_sn, int64_t e_sc, const double* B, int b_mode, int64_t b_sn,
                                         int64_t b_sc, double mass, double mass_sq, double c2, double* out8,
                                         double* reduce_scratch, tp_stream_t stream) {
    if (!out8 || !reduce_scratch) return TP_E_ARG;
    MomRequest mr;
    mr.mass = mass; mr.mass_sq = mass_sq; mr.c2 = c2; mr.out8 = out8; mr.scratch = reduce_scratch;
    return boris_push_device(p, k, E, e_mode, e_sn, e_sc, B, b_mode, b_sn, b_sc, 1, stream, &mr);
}

extern "C" int tp_boris_push_steps_f64(const tp_particles* p, const tp_push_consts* k, const double* E, int e_mode,
                                       int64_t e_sn, int64_t e_sc, const double* B, int b_mode, int64_t b_sn,
                                       int64_t b_sc, int nsteps, tp_stream_t stream) {
    return boris_push_device(p, k, E, e_mode, e_sn, e_sc, B, b_mode, b_sn, b_sc, nsteps, stream);
}

extern "C" int tp_gather_push_f64(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g, int axis,
                                  int formula, uint64_t* status, tp_stream_t stream) {
    return gather_push_device(p, k, g, axis, formula, status, static_cast<cudaStream_t>(stream), 1, nullptr, 0, nullptr, nullptr);
}

extern "C" int tp_gather_push_steps_f64(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g, int axis,
                                        int formula, int nsteps, uint64_t* status, tp_stream_t stream) {
    return gather_push_device(p, k, g, axis, formula, status, static_cast<cudaStream_t>(stream), nsteps, nullptr, 0, nullptr, nullptr);
}

extern "C" int tp_gather_push_windows_f64(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g,
                                          int axis, int formula, int32_t* tile_bounds, int64_t tile_bounds_len,
                                          double* records_scratch, uint64_t* status, tp_stream_t stream) {
    return gather_push_device(p, k, g, axis, formula, status, static_cast<cudaStream_t>(stream), 1, tile_bounds,
                              tile_bounds_len, records_scratch, nullptr);
}

extern "C" int tp_gather_push_moments_f64(const tp_particles* p, const tp_push_consts* k,
                                                  const tp_grid_fields* g, int axis, int formula, int32_t* tile_bounds,
                                                  int64_t tile_bounds_len, double* records_scratch, double mass,
                                                  double mass_sq, double c2, double* out8, double* reduce_scratch,
                                                  uint64_t* status, tp_stream_t stream) {
    if (!out8 || !reduce_scratch) return TP_E_ARG;
    MomRequest mr;
    mr.mass = mass; mr.mass_sq = mass_sq; mr.c2 = c2; mr.out8 = out8; mr.scratch = reduce_scratch;
    if (p && p->n == 0)             // nothing to push: the diagnostic row of an empty ensemble
        return tp_reduce_moments_f64(p->mom, 1, 1, 0, mass, mass_sq, c2, out8, reduce_scratch, stream);
    return gather_push_device(p, k, g, axis, formula, status, static_cast<cudaStream_t>(stream), 1, tile_bounds,
                              tile_bounds_len, records_scratch, &mr);
}

// int32 words of tile bounds tp_gather_push_windows_f64 wants for this ensemble and grid; 0 = the call would not use
// per-tile windows (grid staged whole in shared memory, layout not streamable, too few particles, switched off).
extern "C" int64_t tp_gather_windows_len(const tp_particles* p, const tp_grid_fields* g) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return 0;
    GatherArgs a;
    std::memset(&a, 0, sizeof(a));
    bool vec2 = false, aos16 = false;
    if (fill_particles(p, a.p, vec2, &aos16) || !(vec2 || aos16)) return 0;
    size_t smem = 0; bool staged = false;
    if (fill_grid(g, a, smem, staged) || staged) return 0;
    RingArgs ring; WindowArgs win; int blocks = 0;
    if (!plan_window_ring(a.p.n, ring, win, smem, blocks)) return 0;
    return 2 * ring.ntiles;
}

// Cells per tile the staged window can hold for this ensemble / grid / formula (0: the windowed kernel would not run).
// The Python tool compares it with the primed bounds to decide whether windows pay at all (particles in random order:
// they do not, and the packed 64-byte node records of tp_gather_push_f64 are the better table for the L2 path).
extern "C" int64_t tp_gather_windows_capacity(const tp_particles* p, const tp_grid_fields* g, int formula) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return 0;
    GatherArgs a;
    std::memset(&a, 0, sizeof(a));
    bool vec2 = false, aos16 = false;
    if (fill_particles(p, a.p, vec2, &aos16) || !(vec2 || aos16)) return 0;
    size_t smem = 0; bool staged = false;
    if (fill_grid(g, a, smem, staged) || staged) return 0;
    RingArgs ring; WindowArgs win; int blocks = 0;
    if (!plan_window_ring(a.p.n, ring, win, smem, blocks)) return 0;
    return formula == TP_INTERP_B ? win.win_bytes / kCellBytes : static_cast<int64_t>(win.win_bytes / kRecBytes) - 3;
}

// Prime the tile bounds from the particles' current coordinates along `axis` (first call for an ensemble, or after
// the caller re-ordered it); a zeroed buffer is also valid (the first step then gathers through L2 and primes it).
extern "C" int tp_gather_windows_init(const tp_particles* p, const tp_grid_fields* g, int axis, int32_t* tile_bounds,
                                      int64_t tile_bounds_len, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!p || !g || !tile_bounds || axis < 0 || axis > 2) return TP_E_ARG;
    if (g->ng < 2 || g->ng > (1ll << 31) - 8 || !(g->dr > 0.0)) return TP_E_GRID;
    const int64_t ntiles = p->n / kTile;
    if (tile_bounds_len < 2 * ntiles) return TP_E_ARG;
    if (ntiles == 0) return TP_OK;
    if (p->pos_stride_n <= 0 || p->pos_stride_c <= 0) return TP_E_LAYOUT;
    const int blocks = static_cast<int>(std::min<int64_t>(ntiles, static_cast<int64_t>(d.sm_count) * 8));
    tile_bounds_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p->pos + axis * p->pos_stride_c, p->pos_stride_n, ntiles, g->r_first, 1.0 / g->dr, static_cast<int>(g->ng),
        reinterpret_cast<int2*>(tile_bounds));
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_status_reset(uint64_t* status, tp_stream_t stream) {
    if (!status) return TP_E_ARG;
    static const uint64_t init[4] = {0ull, ~0ull, 0ull, 0ull};
    return static_cast<int>(cudaMemcpyAsync(status, init, sizeof(init), cudaMemcpyHostToDevice,
                                            static_cast<cudaStream_t>(stream)));
}

extern "C" int tp_gather_smem_bytes(const tp_grid_fields* g, int64_t* bytes, int* staged) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    GatherArgs a;
    std::memset(&a, 0, sizeof(a));
    size_t smem = 0; bool st = false;
    int rc = fill_grid(g, a, smem, st);
    if (rc) return rc;
    if (bytes) *bytes = static_cast<int64_t>(smem);
    if (staged) *staged = st ? 1 : 0;
    return TP_OK;
}

static int interp1d_impl(const double* r, int64_t ng, double dr, double r_first, double r_last,
                         const double* y, int ncomp, int64_t y_stride_c, int64_t y_stride_n,
                         const double* xq, int64_t xq_stride, int64_t n, double* out,
                         int64_t out_stride_c, int formula, const tp_linspace* lin, uint64_t* status, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!r || !y || !xq || !out || !status || n < 0 || ncomp < 1) return TP_E_ARG;
    if (y_stride_n <= 0 || y_stride_c <= 0 || xq_stride <= 0 || out_stride_c < n) return TP_E_LAYOUT;
    if (y_stride_c > 0x7fffffff) return TP_E_LAYOUT;
    if (formula < TP_INTERP_A || formula > TP_INTERP_C) return TP_E_MODE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // The whole y block is described as "component 0" (with its full extent) so
    // that the staging plan copies it once; the kernel walks the components.
    tp_grid_fields g;
    std::memset(&g, 0, sizeof(g));
    g.r = r; g.ng = ng; g.dr = dr; g.r_first = r_first; g.r_last = r_last;
    g.comp[0] = y; g.comp_stride[0] = y_stride_n;
    InterpArgs a;
    std::memset(&a, 0, sizeof(a));
    size_t smem = 0; bool staged = false;
    const int64_t extent = (ncomp - 1) * y_stride_c + (ng - 1) * y_stride_n + 1;
    int rc = fill_grid(&g, a.g, smem, staged, extent);
    if (rc) return rc;
    a.g.stride[1] = static_cast<int>(y_stride_c);      // component stride rides in slot 1
    a.g.status = reinterpret_cast<unsigned long long*>(status);
    a.xq = xq; a.xq_stride = xq_stride; a.n = n; a.out = out; a.out_stride_c = out_stride_c; a.ncomp = ncomp;
    if (n == 0) return TP_OK;
    static const bool cells_on = [] {
        const char* e = std::getenv("TURBOPY_B200_INTERP_CELLS");     // tuning / A-B switch, default on
        return !(e && e[0] == '0');
    }();
    if (cells_on && ncomp == 1 && formula == TP_INTERP_B && n >= 8 * ng) {
        // per-cell {y0, slope} table in shared memory (worth building when there are many more points than cells)
        const size_t need = static_cast<size_t>(ng - 1) * (lin ? 16 : 40);       // CellTables: {y0, slope} (+ {x0, x1} + 2 buckets)
        const bool fits3 = need <= static_cast<size_t>(d.smem_optin) / 3 - 1024;    // three CTAs per SM (register path)
        const bool fits1 = need + 3 * kStageI + 2 * kSmemHeader <= static_cast<size_t>(d.smem_optin) - 1024;   // the ring
        if ((fits3 || fits1) && (!lin || lin->n == ng)) {
            InterpArgs ac = a;                            // the cell kernels' view: global arrays, no staging plan
            ac.g.r = r; ac.g.comp[0] = y; ac.g.stride[0] = static_cast<int>(y_stride_n);
            ac.g.nseg = 0;
            const bool vec = xq_stride == 1 && aligned16(xq) && aligned16(out);
            // enough whole tiles for every SM and bulk-copyable arrays: the TMA ring
            static const bool ring_on = [] {
                const char* e = std::getenv("TURBOPY_B200_INTERP_RING");      // tuning / A-B switch, default on
                return !(e && e[0] == '0');
            }();
            InterpRingArgs ra;
            ra.ntiles = n / kTileI;
            ra.table_off = kSmemHeader;
            ra.tiles_off = static_cast<uint32_t>((kSmemHeader + need + 127) & ~size_t(127));
            const int64_t room = static_cast<int64_t>(d.smem_optin) - 1024 - ra.tiles_off;
            static const int ring_depth = [] { const char* e = std::getenv("TURBOPY_B200_INTERP_STAGES"); return e ? std::atoi(e) : 6; }();
            ra.nstages = static_cast<int>(std::min<int64_t>(std::min(kRingMaxStages, std::max(2, ring_depth)), room / kStageI));
            if (ring_on && tma_enabled() && vec && ra.nstages >= 3 && ra.ntiles >= 4ll * d.sm_count) {
                ra.a = ac;
                const size_t bytes = ra.tiles_off + static_cast<size_t>(ra.nstages) * kStageI;
                const int blocks = static_cast<int>(std::min<int64_t>(d.sm_count, ra.ntiles));
                auto ring = [&](auto kern) {
                    int rc2 = set_smem(kern, bytes);
                    if (rc2) return rc2;
                    kern<<<blocks, kTmaThreads, bytes, s>>>(ra);
                    count_launch();
                    return static_cast<int>(cudaGetLastError());
                };
                if (lin) {
                    ra.a.lin = LinGrid{lin->start, lin->span, lin->step, lin->n};
                    return ring(interp_cells_ring_kernel<true>);
                }
                return ring(interp_cells_ring_kernel<false>);
            }
            auto go = [&](auto kern) {
                int rc2 = set_smem(kern, need);
                if (rc2) return rc2;
                const int blocks = grid_blocks(reinterpret_cast<const void*>(kern), need, (n + 3) / 4);
                kern<<<blocks, kThreads, need, s>>>(ac, vec);
                count_launch();
                return static_cast<int>(cudaGetLastError());
            };
            if (fits3) {                                  // (else the table only fits the one-CTA-per-SM ring)
                if (lin) {
                    ac.lin = LinGrid{lin->start, lin->span, lin->step, lin->n};
                    return go(interp_cells_kernel<true>);
                }
                return go(interp_cells_kernel<false>);
            }
        }
    }
    if (cells_on && formula == TP_INTERP_C && n >= 8 * ng && (!lin || lin->n == ng)) {
        // {y[j], y[j+1]} planes for as many components per pass as fit (at most six; + {x0, x1} and the bucket index
        // when x is not a linspace).  Fewer planes than components means more passes over the points (8 B/point
        // each): still worth it on a linspace grid while a pass carries >= 2 components, and always on other
        // abscissas, where the uniform-spacing guess of interp_kernel sends every point to the exact path.
        const size_t room = static_cast<size_t>(d.smem_optin) - 1024;
        const size_t fixed = static_cast<size_t>(ng - 1) * (lin ? 0 : 24);
        const int fit = room > fixed ? static_cast<int>(std::min<size_t>(6, (room - fixed) / (static_cast<size_t>(ng - 1) * 16))) : 0;
        const int ncpass = std::min(ncomp, fit);
        const size_t need = fixed + static_cast<size_t>(ng - 1) * 16 * ncpass;
        if (ncpass >= 1 && (ncpass == ncomp || ncpass >= 2 || !lin)) {
            InterpArgs ac = a;                                // global arrays, no staging plan
            ac.g.r = r; ac.g.comp[0] = y; ac.g.stride[0] = static_cast<int>(y_stride_n);
            ac.g.nseg = 0;
            ac.ncpass = ncpass;
            const bool vec = xq_stride == 1 && aligned16(xq) && aligned16(out) && out_stride_c % 2 == 0;
            auto go = [&](auto kern) {
                int rc2 = set_smem(kern, need);
                if (rc2) return rc2;
                const int64_t want = ((n + 3) / 4 + kCellsNcThreads - 1) / kCellsNcThreads;
                const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(d.sm_count, want)));
                kern<<<blocks, kCellsNcThreads, need, s>>>(ac, vec);
                count_launch();
                return static_cast<int>(cudaGetLastError());
            };
            if (lin) {
                ac.lin = LinGrid{lin->start, lin->span, lin->step, lin->n};
                return go(interp_cells_nc_kernel<true>);
            }
            return go(interp_cells_nc_kernel<false>);
        }
    }
    if (lin) {
        if (lin->n != ng) return TP_E_ARG;
        a.lin = LinGrid{lin->start, lin->span, lin->step, lin->n};
        if (staged && formula == TP_INTERP_B) return launch_interp<TP_INTERP_B, true, true>(a, smem, s);
        if (staged && formula == TP_INTERP_C) return launch_interp<TP_INTERP_C, true, true>(a, smem, s);
    }                                                    // (formula A / grids gathered through L2: the generic kernels)
    if (staged) {
        switch (formula) {
            case TP_INTERP_A: return launch_interp<TP_INTERP_A, true>(a, smem, s);
            case TP_INTERP_B: return launch_interp<TP_INTERP_B, true>(a, smem, s);
            default: return launch_interp<TP_INTERP_C, true>(a, smem, s);
        }
    }
    switch (formula) {
        case TP_INTERP_A: return launch_interp<TP_INTERP_A, false>(a, 0, s);
        case TP_INTERP_B: return launch_interp<TP_INTERP_B, false>(a, 0, s);
        default: return launch_interp<TP_INTERP_C, false>(a, 0, s);
    }
}

extern "C" int tp_interp1d_f64(const double* r, int64_t ng, double dr, double r_first, double r_last,
                               const double* y, int ncomp, int64_t y_stride_c, int64_t y_stride_n,
                               const double* xq, int64_t xq_stride, int64_t n, double* out,
                               int64_t out_stride_c, int formula, uint64_t* status, tp_stream_t stream) {
    return interp1d_impl(r, ng, dr, r_first, r_last, y, ncomp, y_stride_c, y_stride_n, xq, xq_stride, n, out, out_stride_c,
                         formula, nullptr, status, stream);
}

extern "C" int tp_interp1d_lin_f64(const double* r, int64_t ng, double dr, double r_first, double r_last,
                                   const double* y, int ncomp, int64_t y_stride_c, int64_t y_stride_n,
                                   const double* xq, int64_t xq_stride, int64_t n, double* out,
                                   int64_t out_stride_c, int formula, const tp_linspace* grid, uint64_t* status,
                                   tp_stream_t stream) {
    if (!grid) return TP_E_ARG;
    return interp1d_impl(r, ng, dr, r_first, r_last, y, ncomp, y_stride_c, y_stride_n, xq, xq_stride, n, out, out_stride_c,
                         formula, grid, status, stream);
}
