// Kernel parameter blocks and particle I/O shared by the push / gather kernels:
// 128-bit SoA pair accessors, generic strided single-particle accessors, the L2
// bulk prefetch of the register path, and the E / B arguments of the un-fused push.
#pragma once
#include "boris_math.cuh"

namespace tp {

#ifndef TP_MIN_BLOCKS
#define TP_MIN_BLOCKS 2
#endif
constexpr int kThreads = 256;           // register-path kernels

// ---------------------------------------------------------------------------
// kernel parameter blocks
// ---------------------------------------------------------------------------
struct ParticleArgs {
    double* pos; int64_t pos_sn, pos_sc;
    double* mom; int64_t mom_sn, mom_sc;
    int64_t n;
};

struct FieldArg {           // E or B of the un-fused push
    const double* ptr;      // device pointer (modes 1, 2)
    int mode;               // TP_FIELD_*
    int64_t sn, sc;         // per-particle strides
    double u[3];            // uniform value (mode 0)
};

struct PushArgs {
    ParticleArgs p;
    PushK k;
    FieldArg E, B;
    int nsteps;             // consecutive pushes with the same E, B per visit of a particle (>= 1)
};

constexpr int kMaxSeg = 7;  // node coordinates + up to six separately stored components

struct StageSeg {
    const double* src;      // global source, first element
    uint32_t smem_off;      // byte offset of that element in dynamic shared memory
    uint32_t tma_bytes;     // leading part moved by one cp.async.bulk (multiple of 16, src 16-B aligned)
    uint32_t n_doubles;     // total elements of the segment
};

struct GatherArgs {
    ParticleArgs p;
    PushK k;
    // global view (also the source of the staging copies)
    const double* r;
    const double* comp[6];
    int stride[6];
    int ng;
    double r_first, r_last, inv_dr, dr;
    int axis;
    // staging plan (staged variant only)
    int nseg;
    StageSeg seg[kMaxSeg];
    uint32_t r_smem_off;
    uint32_t comp_smem_off[6];
    unsigned long long* status;
    const double* records;      // packed 64-byte node records (gather mode 2), else nullptr
    int nsteps;                 // consecutive gather+push steps on unchanged grid fields per visit of a particle (>= 1)
};

// ---------------------------------------------------------------------------
// particle I/O
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_pair(const ParticleArgs& a, int64_t i, Vec3 (&x)[2], Vec3 (&p)[2]) {
    const double2 x0 = ld2(a.pos + i);
    const double2 x1 = ld2(a.pos + a.pos_sc + i);
    const double2 x2 = ld2(a.pos + 2 * a.pos_sc + i);
    const double2 p0 = ld2(a.mom + i);
    const double2 p1 = ld2(a.mom + a.mom_sc + i);
    const double2 p2 = ld2(a.mom + 2 * a.mom_sc + i);
    x[0].x = x0.x; x[1].x = x0.y; x[0].y = x1.x; x[1].y = x1.y; x[0].z = x2.x; x[1].z = x2.y;
    p[0].x = p0.x; p[1].x = p0.y; p[0].y = p1.x; p[1].y = p1.y; p[0].z = p2.x; p[1].z = p2.y;
}

__device__ __forceinline__ void store_pair(const ParticleArgs& a, int64_t i, const Vec3 (&x)[2], const Vec3 (&p)[2]) {
    st2(a.pos + i, make_double2(x[0].x, x[1].x));
    st2(a.pos + a.pos_sc + i, make_double2(x[0].y, x[1].y));
    st2(a.pos + 2 * a.pos_sc + i, make_double2(x[0].z, x[1].z));
    st2(a.mom + i, make_double2(p[0].x, p[1].x));
    st2(a.mom + a.mom_sc + i, make_double2(p[0].y, p[1].y));
    st2(a.mom + 2 * a.mom_sc + i, make_double2(p[0].z, p[1].z));
}

// One thread per CTA asks the TMA to pull the six 4 KB component slices of the
// CTA's tile `first .. first + 2*kThreads` (a later loop iteration) into L2, so
// that the loads of that iteration hit L2 instead of paying the DRAM latency
// with only 16 resident warps per SM.  SoA fast path only (16-byte aligned).
#ifndef TP_PF_DIST
#define TP_PF_DIST 2      /* iterations ahead; 0/1/2/3 measured: 46.3 / 51.0 / 51.3 / 50.6 G pushes/s */
#endif
__device__ __forceinline__ void prefetch_tile(const ParticleArgs& a, int64_t first, int64_t n_even) {
#if TP_PF_DIST > 0
    if (threadIdx.x == 0 && first < n_even) {
        const int64_t rest = n_even - first;
        const uint32_t bytes = static_cast<uint32_t>((rest < 2 * kThreads ? rest : 2 * kThreads) * sizeof(double));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pos + c * a.pos_sc + first), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.mom + c * a.mom_sc + first), "r"(bytes)
                         : "memory");
        }
    }
#endif
}

__device__ __forceinline__ void load_one(const ParticleArgs& a, int64_t i, Vec3& x, Vec3& p) {
    const double* xs = a.pos + i * a.pos_sn;
    const double* ps = a.mom + i * a.mom_sn;
    x.x = xs[0]; x.y = xs[a.pos_sc]; x.z = xs[2 * a.pos_sc];
    p.x = ps[0]; p.y = ps[a.mom_sc]; p.z = ps[2 * a.mom_sc];
}

__device__ __forceinline__ void store_one(const ParticleArgs& a, int64_t i, const Vec3& x, const Vec3& p) {
    double* xs = a.pos + i * a.pos_sn;
    double* ps = a.mom + i * a.mom_sn;
    xs[0] = x.x; xs[a.pos_sc] = x.y; xs[2 * a.pos_sc] = x.z;
    ps[0] = p.x; ps[a.mom_sc] = p.y; ps[2 * a.mom_sc] = p.z;
}

__device__ __forceinline__ Vec3 field_at(const FieldArg& f, const Vec3& uni, int64_t i) {
    if (f.mode != TP_FIELD_PER_PARTICLE) return uni;
    const double* s = f.ptr + i * f.sn;
    Vec3 v; v.x = s[0]; v.y = s[f.sc]; v.z = s[2 * f.sc];
    return v;
}

__device__ __forceinline__ Vec3 uniform_of(const FieldArg& f) {
    Vec3 v; v.x = f.u[0]; v.y = f.u[1]; v.z = f.u[2];
    if (f.mode == TP_FIELD_UNIFORM_DEV) { v.x = f.ptr[0]; v.y = f.ptr[1]; v.z = f.ptr[2]; }
    return v;
}

}  // namespace tp
