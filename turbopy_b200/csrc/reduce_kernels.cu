// Kernel 3 of the hot path: kinetic-energy / momentum reduction for the
// energy / moment Diagnostics.  The reference has no such function (parity
// unpinned; definition and tolerance in oracle/moments.py): per particle
//   m1 = sqrt(mass_sq + |p|^2/c2)         (gamma*m of turbopy/computetools.py:430-431)
//   ke = |p|^2 / (m1 + mass)
// and the sums {KE, Px, Py, Pz, sum|p|^2, n}.  24 B read per particle.
// Deterministic: per-thread serial sums -> warp shuffle tree -> shared-memory
// tree -> per-CTA partials; a second single-CTA pass folds the partials in a
// fixed order (no atomics), so two runs on the same data agree bit for bit.
#include "moments_math.cuh"

#include <algorithm>
#include <cstdlib>

#include <cmath>

namespace tp {

template <bool FAST>
__device__ __forceinline__ bool moments_thread(const double* __restrict__ mom, int64_t sn, int64_t sc, int64_t n,
                                               const MomConsts& k, bool vec, double (&acc)[kRedVals]) {
#pragma unroll
    for (int v = 0; v < kRedVals; ++v) acc[v] = 0.0;
    Guard g;
    const int64_t gtid = static_cast<int64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * kRedThreads;
    if (vec) {                                            // SoA, 128-bit loads
        const int64_t npairs = n >> 1;
        for (int64_t pr = gtid; pr < npairs; pr += gstride) {
            const double2 a = ld2(mom + 2 * pr), b = ld2(mom + sc + 2 * pr), c = ld2(mom + 2 * sc + 2 * pr);
            accumulate<FAST>(a.x, b.x, c.x, k, g, acc);
            accumulate<FAST>(a.y, b.y, c.y, k, g, acc);
        }
        if ((n & 1) && gtid == 0) accumulate<FAST>(mom[n - 1], mom[sc + n - 1], mom[2 * sc + n - 1], k, g, acc);
    } else {
        for (int64_t i = gtid; i < n; i += gstride) {
            const double* p = mom + i * sn;
            accumulate<FAST>(p[0], p[sc], p[2 * sc], k, g, acc);
        }
    }
    return !FAST || (guard_ok(g) && k.fast_ok);
}

__global__ void __launch_bounds__(kRedThreads, 4) moments_partial_kernel(const double* __restrict__ mom, int64_t sn,
                                                                      int64_t sc, int64_t n, const MomConsts k,
                                                                      bool vec, double* scratch) {
    double acc[kRedVals];
    if (!moments_thread<true>(mom, sn, sc, n, k, vec, acc))      // some operand left the fast path's range:
        moments_thread<false>(mom, sn, sc, n, k, vec, acc);      // redo this thread's share with IEEE operations
    block_sum(acc, scratch + static_cast<int64_t>(blockIdx.x) * 8);
}

__global__ void __launch_bounds__(kRedThreads) moments_final_kernel(const double* scratch, int nblocks, int64_t n,
                                                                    double* out8) {
    double acc[kRedVals] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += kRedThreads)
#pragma unroll
        for (int k = 0; k < kRedVals; ++k) acc[k] += scratch[static_cast<int64_t>(b) * 8 + k];
    __shared__ double res[8];
    block_sum(acc, res);
    __syncthreads();
    if (threadIdx.x == 0) {
        out8[0] = res[0]; out8[1] = res[1]; out8[2] = res[2]; out8[3] = res[3]; out8[4] = res[4];
        out8[5] = static_cast<double>(n); out8[6] = 0.0; out8[7] = 0.0;
    }
}

int launch_moments_final(const double* scratch, int nblocks, int64_t n, double* out8, cudaStream_t s) {
    moments_final_kernel<<<1, kRedThreads, 0, s>>>(scratch, nblocks, n, out8);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_reduce_scratch_doubles(void) { return static_cast<int64_t>(kRedMaxBlocks) * 8; }

extern "C" int tp_reduce_moments_f64(const double* mom, int64_t stride_n, int64_t stride_c, int64_t n, double mass,
                                     double mass_sq, double c2, double* out8, double* scratch, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (n < 0 || !out8 || !scratch || (n > 0 && !mom)) return TP_E_ARG;
    if (stride_n <= 0 || stride_c <= 0) return TP_E_LAYOUT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool vec = layout_of(stride_n, stride_c, n) == kLayoutSoA && aligned16(mom) && (stride_c % 2 == 0) && n >= 2;
    const int64_t items = vec ? (n + 1) / 2 : n;
    int64_t blocks = (items + kRedThreads - 1) / kRedThreads;
    static int per_sm = 0;                                   // one resident wave of persistent CTAs
    if (per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, moments_partial_kernel, kRedThreads, 0) != cudaSuccess || v < 1) v = 4;
        if (const char* e = std::getenv("TURBOPY_B200_RED_BLOCKS")) v = std::max(1, std::atoi(e));   // tuning knob
        per_sm = v;
    }
    const int64_t full = static_cast<int64_t>(d.sm_count) * per_sm;
    if (blocks > full) blocks = full;
    if (blocks > kRedMaxBlocks) blocks = kRedMaxBlocks;
    if (blocks < 1) blocks = 1;
    const MomConsts k = make_mom_consts(mass, mass_sq, c2);
    moments_partial_kernel<<<static_cast<int>(blocks), kRedThreads, 0, s>>>(mom, stride_n, stride_c, n, k, vec, scratch);
    count_launch();
    const int rc = static_cast<int>(cudaGetLastError());
    return rc ? rc : launch_moments_final(scratch, static_cast<int>(blocks), n, out8, s);
}
