// TMA / mbarrier PTX wrappers and the staging of the 1-D grid (node coordinates
// + field components) into shared memory with cp.async.bulk; the GridView the
// gather math reads (shared or global memory); the device status block.
#pragma once
#include "gather_math.cuh"
#include "particle_io.cuh"

namespace tp {

// ---------------------------------------------------------------------------
// TMA bulk staging of the grid into shared memory
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// L2 eviction policies: the particle stream is touched once per step (evict_first) so that it does not push
// a large grid's node records out of L2 (evict_last on their loads).
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
// instead of re-issuing try_wait + branch every few hundred cycles (ncu, round 2: ~33 of 620 warp instructions per
// tile iteration were such spins -- issue slots and power that a power-capped kernel wants back).
#ifndef TP_MBAR_SUSPEND_NS
#define TP_MBAR_SUSPEND_NS 0x989680
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(static_cast<uint32_t>(TP_MBAR_SUSPEND_NS))
        : "memory");
}

constexpr uint32_t kSmemHeader = 128;   // mbarrier lives in the first 8 bytes

// Issue the staging copies (one elected thread drives the TMA; pieces the TMA
// cannot take -- a source that is not 16-byte aligned, an 8-byte tail -- are
// copied by the CTA's threads).  Completion: stage_wait().
__device__ __forceinline__ void stage_issue(const GatherArgs& a, unsigned char* smem) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int s = 0; s < a.nseg; ++s) total += a.seg[s].tma_bytes;
        mbar_expect_tx(bar, total);
        for (int s = 0; s < a.nseg; ++s)
            if (a.seg[s].tma_bytes)
                tma_bulk_g2s(smem + a.seg[s].smem_off, a.seg[s].src, a.seg[s].tma_bytes, bar);
    }
    for (int s = 0; s < a.nseg; ++s) {
        const uint32_t first = a.seg[s].tma_bytes >> 3;
        double* dst = reinterpret_cast<double*>(smem + a.seg[s].smem_off);
        for (uint32_t e = first + threadIdx.x; e < a.seg[s].n_doubles; e += blockDim.x) dst[e] = a.seg[s].src[e];
    }
}

__device__ __forceinline__ void stage_wait(unsigned char* smem) {
    mbar_wait(reinterpret_cast<uint64_t*>(smem), 0);
    __syncthreads();                 // the thread-copied pieces
}

template <bool STAGED>
__device__ __forceinline__ GridView make_view(const GatherArgs& a, unsigned char* smem) {
    GridView g;
    g.ng = a.ng; g.r_first = a.r_first; g.r_last = a.r_last; g.inv_dr = a.inv_dr; g.dr = a.dr; g.qa = 0.0;
    if (STAGED) {
        g.r = reinterpret_cast<const double*>(smem + a.r_smem_off);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            g.comp[c] = a.comp[c] ? reinterpret_cast<const double*>(smem + a.comp_smem_off[c]) : nullptr;
            g.stride[c] = a.stride[c];
        }
    } else {
        g.r = a.r;
#pragma unroll
        for (int c = 0; c < 6; ++c) { g.comp[c] = a.comp[c]; g.stride[c] = a.stride[c]; }
    }
    return g;
}

__device__ __forceinline__ void flag_particle(unsigned long long* status, int64_t i, unsigned st) {
    atomicAdd(&status[0], 1ull);
    atomicMin(&status[1], static_cast<unsigned long long>(i));
    atomicOr(&status[2], static_cast<unsigned long long>(st));
}

__device__ __forceinline__ double axis_of(const Vec3& x, int axis) {
    return axis == 0 ? x.x : (axis == 1 ? x.y : x.z);
}

}  // namespace tp
