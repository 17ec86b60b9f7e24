// Per-particle arithmetic and block reductions of the energy / moment diagnostic (kernel 3), shared by the
// standalone reduction (reduce_kernels.cu) and by the fused gather+push kernel that accumulates the moments of the
// momenta it has just loaded (push_kernels.cu: gather_push_win_kernel<.., MOM = true>).
#pragma once
#include "common.cuh"

namespace tp {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 2048;
constexpr int kRedVals = 5;               // KE, Px, Py, Pz, |p|^2

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic CTA sum of kRedVals values per thread: warp shuffle tree, then warp 0 folds the per-warp sums.
// THREADS = blockDim.x (a multiple of 32, at most 1024); every thread of the CTA must call it.
template <int THREADS = kRedThreads>
__device__ __forceinline__ void block_sum(double (&v)[kRedVals], double* out) {
    static_assert(THREADS % 32 == 0 && THREADS <= 1024, "whole warps");
    __shared__ double sh[kRedVals][THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kRedVals; ++k) {
        const double w = warp_sum(v[k]);
        if (lane == 0) sh[k][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < kRedVals; ++k) {
            double w = lane < THREADS / 32 ? sh[k][lane] : 0.0;
            w = warp_sum(w);
            if (lane == 0) out[k] = w;
        }
    }
}

struct MomConsts { double mass, mm, c2, rc2; int fast_ok; };

inline MomConsts make_mom_consts(double mass, double mass_sq, double c2) {
    MomConsts k;
    k.mass = mass; k.mm = mass_sq; k.c2 = c2; k.rc2 = 1.0 / c2;
    k.fast_ok = const_divisor_ok(c2) ? 1 : 0;
    return k;
}

// FAST: shared-reciprocal quotients and the branch-free sqrt of common.cuh -- the same bits as the IEEE
// operations whenever the guard stays clean (the caller checks it and redoes its share with FAST = false).
template <bool FAST>
__device__ __forceinline__ void accumulate(double px, double py, double pz, const MomConsts& k, Guard& g,
                                           double (&acc)[kRedVals]) {
    const double p2 = (px * px + py * py) + pz * pz;
    const double m1 = root<FAST>(k.mm + quot<FAST>(p2, const_recip<FAST>(k.c2, k.rc2), g), g);
    acc[0] += quot<FAST>(p2, make_recip<FAST>(m1 + k.mass, g), g);
    acc[1] += px; acc[2] += py; acc[3] += pz; acc[4] += p2;
}

// Folds `nblocks` per-CTA partials (8 doubles apart in `scratch`) into out8 = {KE, Px, Py, Pz, |p|^2, n, 0, 0}
// in a fixed order: one CTA, no atomics (reduce_kernels.cu).
int launch_moments_final(const double* scratch, int nblocks, int64_t n, double* out8, cudaStream_t s);

}  // namespace tp
