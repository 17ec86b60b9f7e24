// The small-tile ring: ONE particle per consumer thread, 480-particle tiles.
// Same machinery as tma_ring.cuh (producer lane + 15 consumer warps, full / done
// mbarrier pairs, bulk loads and bulk stores); two users:
//
// * BorisPush.push with per-particle E and B -- the reference signature called
//   with (n,3) field arrays (turbopy/computetools.py:407-441): 144 B per
//   particle-step (x, p read and written, E and B read).  FIELDS = true: a stage
//   keeps the particle ring's 45 KB and carries the fields next to the state,
//       [ x y z px py pz | E | B ]   480 particles, 12 x 3840 B
//   where each of the four (n,3) arrays arrives either as three component slices
//   (SoA) or as one contiguous 11.25 KB block (C-order (n,3), torch's default).
//   Only the first 22.5 KB (the state) is stored back.
// * BorisPush.push with uniform E, B (FIELDS = false): 22.5 KB stages holding only
//   the state, four of them -- for this short per-tile computation the smaller,
//   deeper ring beats the 960-particle one (98 % vs 94 % of the HBM peak).
#pragma once
#include "tma_ring.cuh"

namespace tp {

constexpr int kTileF = kConsumers;                              // particles per tile
constexpr uint32_t kSliceF = kTileF * sizeof(double);           // 3840 B per component
constexpr uint32_t kArrayF = 3 * kSliceF;                       // one (tile,3) array
static_assert(4 * kArrayF == kTileBytes, "field ring stages are sized like particle ring stages");

template <bool AOS>
__device__ __forceinline__ void array_load(const double* base, int64_t sc, int64_t first, unsigned char* dst, uint64_t* bar) {
    if (AOS) {
        tma_bulk_g2s(dst, base + 3 * first, kArrayF, bar);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_bulk_g2s(dst + c * kSliceF, base + c * sc + first, kSliceF, bar);
    }
}

template <bool AOS>
__device__ __forceinline__ void array_store(double* base, int64_t sc, int64_t first, const unsigned char* src) {
    if (AOS) {
        tma_bulk_s2g(base + 3 * first, src, kArrayF);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_bulk_s2g(base + c * sc + first, src + c * kSliceF, kSliceF);
    }
}

// the calling thread's element of a staged (tile,3) array
template <bool AOS>
__device__ __forceinline__ double* slot_of(unsigned char* array) {
    return reinterpret_cast<double*>(array) + (AOS ? 3 * threadIdx.x : threadIdx.x);
}
template <bool AOS>
__device__ __forceinline__ Vec3 slot_read(unsigned char* array) {
    constexpr int pitch = AOS ? 1 : kTileF;
    const double* s = slot_of<AOS>(array);
    Vec3 v; v.x = s[0]; v.y = s[pitch]; v.z = s[2 * pitch];
    return v;
}

// PA / FA: layout of the particle arrays / of the field arrays (true = C-order (n,3)).
// FIELDS = false: the stage holds only the state (22.5 KB; uniform E, B come from the closure), which
// allows a deeper ring of smaller tiles.  fast(E, B, x, p) -> bool; recover(i, xs, ps, pitch) as in tma_ring.
// observe(p) sees the thread's momentum AS LOADED, before `fast` updates it (the fused moment diagnostic).
template <bool PA, bool FA, bool FIELDS, typename Fast, typename Recover, typename Observe>
__device__ __forceinline__ void tma_ring_fields(const PushArgs& a, const RingArgs& ra, unsigned char* smem, Fast fast,
                                                Recover recover, Observe observe) {
    constexpr uint32_t kStage = FIELDS ? kTileBytes : 2 * kArrayF;
    uint64_t* full = ring_full(smem);
    uint64_t* done = ring_done(smem);
    unsigned char* tiles = smem + ra.tiles_off;
    const int nst = ra.nstages;
    const int64_t t0 = blockIdx.x, tstep = gridDim.x;
    auto load = [&](int64_t t, int s) {
        unsigned char* st = tiles + s * kStage;
        const int64_t first = t * kTileF;
        mbar_expect_tx(&full[s], kStage);
        array_load<PA>(a.p.pos, a.p.pos_sc, first, st, &full[s]);
        array_load<PA>(a.p.mom, a.p.mom_sc, first, st + kArrayF, &full[s]);
        if (FIELDS) {
            array_load<FA>(a.E.ptr, a.E.sc, first, st + 2 * kArrayF, &full[s]);
            array_load<FA>(a.B.ptr, a.B.sc, first, st + 3 * kArrayF, &full[s]);
        }
    };
    int s = 0;
    uint32_t parity = 0;
    if (threadIdx.x >= kConsumers) {
        if (threadIdx.x == kConsumers) {                            // the producer lane
            for (int i = 0; i < nst; ++i) {
                const int64_t t = t0 + i * tstep;
                if (t < ra.ntiles) load(t, i);
            }
            for (int64_t t = t0; t < ra.ntiles; t += tstep) {
                mbar_wait(&done[s], parity);
                const unsigned char* st = tiles + s * kStage;
                array_store<PA>(a.p.pos, a.p.pos_sc, t * kTileF, st);
                array_store<PA>(a.p.mom, a.p.mom_sc, t * kTileF, st + kArrayF);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                const int64_t tn = t + nst * tstep;
                if (tn < ra.ntiles) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    load(tn, s);
                }
                if (++s == nst) { s = 0; parity ^= 1u; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    constexpr int pitch = PA ? 1 : kTileF;
    for (int64_t t = t0; t < ra.ntiles; t += tstep) {
        mbar_wait(&full[s], parity);
        if (ra.consumer_delay_ns) __nanosleep(ra.consumer_delay_ns);          // pacing of fast consumers (tma_ring.cuh)
        unsigned char* st = tiles + s * kStage;
        Vec3 x = slot_read<PA>(st), p = slot_read<PA>(st + kArrayF);
        Vec3 E = {0.0, 0.0, 0.0}, B = {0.0, 0.0, 0.0};
        if (FIELDS) { E = slot_read<FA>(st + 2 * kArrayF); B = slot_read<FA>(st + 3 * kArrayF); }
        double* xs = slot_of<PA>(st);
        double* ps = slot_of<PA>(st + kArrayF);
        observe(p);
        if (__builtin_expect(fast(E, B, x, p), 1)) slot_write(xs, ps, pitch, x, p);
        else recover(t * kTileF + threadIdx.x, xs, ps, pitch);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
        if (++s == nst) { s = 0; parity ^= 1u; }
    }
}

template <bool PA, bool FA, bool FIELDS, typename Fast, typename Recover>
__device__ __forceinline__ void tma_ring_fields(const PushArgs& a, const RingArgs& ra, unsigned char* smem, Fast fast,
                                                Recover recover) {
    tma_ring_fields<PA, FA, FIELDS>(a, ra, smem, fast, recover, [](const Vec3&) {});
}

}  // namespace tp
