// The warp-specialised shared-memory ring that streams the particle arrays
// through the TMA (producer warp + consumer warps, mbarrier full / done pairs).
#pragma once
#include "grid_stage.cuh"

namespace tp {

// ---------------------------------------------------------------------------
// particles streamed by the TMA: the shared-memory ring
// ---------------------------------------------------------------------------
// The register-path kernels (push_kernel / gather_push_kernel) keep only 16 warps
// per SM resident (128 registers per thread) and their loads sit at the end of
// each iteration, so the HBM latency is exposed (profiles/r1_v1_gather_push.md:
// long_scoreboard is the top stall).
// Here the particle arrays themselves go through the TMA: one persistent CTA
// per SM owns a ring of NST shared-memory stages of 45 KB; a stage holds one
// tile of kTile = 960 particles, either as six contiguous 7.5 KB component
// slices (SoA) or as two contiguous 22.5 KB (n,3) C-order blocks (AoS, numpy's
// layout).  The CTA is warp-specialised:
//   producer (warp 15, one elected lane) -- fills a stage with cp.async.bulk
//     loads that complete on the stage's `full` mbarrier (expect_tx = 45 KB);
//     when the consumers have released a stage (`done` mbarrier) it drains the
//     updated tile with cp.async.bulk stores (bulk async-group), waits until
//     those stores have finished READING shared memory (wait_group.read 0 --
//     only the producer stalls) and immediately re-fills the stage with the
//     tile NST iterations ahead;
//   consumers (warps 0-14, 480 threads) -- wait on `full`, read their two
//     particles with 128-bit LDS (conflict-free in both layouts), run the same
//     fast / exact arithmetic as the register path, write the result back with
//     128-bit STS, fence.proxy.async, and arrive on `done` once per warp.  There
//     is no CTA-wide barrier in the loop: the warps drift apart and overlap
//     each other's LDS / fp64 / STS phases.
// No thread ever waits on a global load, and NST-1 tiles of loads are in flight
// behind the tile being computed (a 3-stage ring now behaves like the earlier
// thread-0-driven 4-stage one: 88 % -> 95 % of the HBM peak at ng = 4097).
constexpr int kTmaThreads = 512;                               // 15 consumer warps + 1 producer warp
constexpr int kConsumers = 480;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kTile = 2 * kConsumers;                          // particles per tile
constexpr uint32_t kSliceBytes = kTile * sizeof(double);       // 7.5 KB per component
constexpr uint32_t kTileBytes = 6 * kSliceBytes;               // 45 KB per stage
constexpr int kMaxStages = 3;                                  // measured: 2 / 3 / 4 stages = 84 / 97 / 90 % of peak

struct RingArgs {
    int nstages;
    uint32_t tiles_off;        // byte offset of stage 0 in dynamic shared memory (128-B aligned)
    int64_t ntiles;            // full tiles; the remaining n - ntiles*kTile particles use the scalar path
    int stream_hint;           // 1: particle loads / stores carry the L2 evict_first policy
    int consumer_delay_ns;     // pacing of fast consumers: sleep this long per tile before touching it (0 = none)
};

__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void tma_bulk_s2g_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes), "l"(policy)
                 : "memory");
}

// SoA tiles with the evict_first policy (the large-grid kernels: keeps the node records in L2)
__device__ __forceinline__ void tile_load_hint(const ParticleArgs& p, int64_t tile, unsigned char* stage, uint64_t* bar,
                                               uint64_t policy) {
    const int64_t first = tile * kTile;
    mbar_expect_tx(bar, kTileBytes);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        tma_bulk_g2s_hint(stage + c * kSliceBytes, p.pos + c * p.pos_sc + first, kSliceBytes, bar, policy);
        tma_bulk_g2s_hint(stage + (3 + c) * kSliceBytes, p.mom + c * p.mom_sc + first, kSliceBytes, bar, policy);
    }
}
__device__ __forceinline__ void tile_store_hint(const ParticleArgs& p, int64_t tile, const unsigned char* stage, uint64_t policy) {
    const int64_t first = tile * kTile;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        tma_bulk_s2g_hint(p.pos + c * p.pos_sc + first, stage + c * kSliceBytes, kSliceBytes, policy);
        tma_bulk_s2g_hint(p.mom + c * p.mom_sc + first, stage + (3 + c) * kSliceBytes, kSliceBytes, policy);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <bool AOS>
__device__ __forceinline__ void tile_load(const ParticleArgs& p, int64_t tile, unsigned char* stage, uint64_t* bar) {
    const int64_t first = tile * kTile;
    mbar_expect_tx(bar, kTileBytes);
    if (AOS) {
        tma_bulk_g2s(stage, p.pos + 3 * first, 3 * kSliceBytes, bar);
        tma_bulk_g2s(stage + 3 * kSliceBytes, p.mom + 3 * first, 3 * kSliceBytes, bar);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            tma_bulk_g2s(stage + c * kSliceBytes, p.pos + c * p.pos_sc + first, kSliceBytes, bar);
            tma_bulk_g2s(stage + (3 + c) * kSliceBytes, p.mom + c * p.mom_sc + first, kSliceBytes, bar);
        }
    }
}

template <bool AOS>
__device__ __forceinline__ void tile_store(const ParticleArgs& p, int64_t tile, const unsigned char* stage) {
    const int64_t first = tile * kTile;
    if (AOS) {
        tma_bulk_s2g(p.pos + 3 * first, stage, 3 * kSliceBytes);
        tma_bulk_s2g(p.mom + 3 * first, stage + 3 * kSliceBytes, 3 * kSliceBytes);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            tma_bulk_s2g(p.pos + c * p.pos_sc + first, stage + c * kSliceBytes, kSliceBytes);
            tma_bulk_s2g(p.mom + c * p.mom_sc + first, stage + (3 + c) * kSliceBytes, kSliceBytes);
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Where the calling thread's particle v (0 or 1) lives inside a stage: pointers
// to its x and px components; y/z (py/pz) follow at multiples of pitch doubles.
template <bool AOS>
struct Slot {
    static constexpr int pitch = AOS ? 1 : kTile;
    __device__ static double* x(double* stage, int v) { return AOS ? stage + 6 * threadIdx.x + 3 * v : stage + 2 * threadIdx.x + v; }
    __device__ static double* p(double* stage, int v) { return x(stage, v) + 3 * kTile; }
};

template <bool AOS>
__device__ __forceinline__ void stage_read(const double* stage, Vec3 (&x)[2], Vec3 (&p)[2]) {
    if (AOS) {          // six contiguous doubles per array: x0 y0 z0 x1 y1 z1
        const double* b = stage + 6 * threadIdx.x;
        const double2 a0 = ld2(b), a1 = ld2(b + 2), a2 = ld2(b + 4);
        const double2 m0 = ld2(b + 3 * kTile), m1 = ld2(b + 3 * kTile + 2), m2 = ld2(b + 3 * kTile + 4);
        x[0].x = a0.x; x[0].y = a0.y; x[0].z = a1.x; x[1].x = a1.y; x[1].y = a2.x; x[1].z = a2.y;
        p[0].x = m0.x; p[0].y = m0.y; p[0].z = m1.x; p[1].x = m1.y; p[1].y = m2.x; p[1].z = m2.y;
    } else {
        const double* b = stage + 2 * threadIdx.x;
        const double2 x0 = ld2(b), x1 = ld2(b + kTile), x2 = ld2(b + 2 * kTile);
        const double2 p0 = ld2(b + 3 * kTile), p1 = ld2(b + 4 * kTile), p2 = ld2(b + 5 * kTile);
        x[0].x = x0.x; x[1].x = x0.y; x[0].y = x1.x; x[1].y = x1.y; x[0].z = x2.x; x[1].z = x2.y;
        p[0].x = p0.x; p[1].x = p0.y; p[0].y = p1.x; p[1].y = p1.y; p[0].z = p2.x; p[1].z = p2.y;
    }
}

template <bool AOS>
__device__ __forceinline__ void stage_write(double* stage, const Vec3 (&x)[2], const Vec3 (&p)[2]) {
    if (AOS) {
        double* b = stage + 6 * threadIdx.x;
        st2(b, make_double2(x[0].x, x[0].y)); st2(b + 2, make_double2(x[0].z, x[1].x)); st2(b + 4, make_double2(x[1].y, x[1].z));
        b += 3 * kTile;
        st2(b, make_double2(p[0].x, p[0].y)); st2(b + 2, make_double2(p[0].z, p[1].x)); st2(b + 4, make_double2(p[1].y, p[1].z));
    } else {
        double* b = stage + 2 * threadIdx.x;
        st2(b, make_double2(x[0].x, x[1].x));
        st2(b + kTile, make_double2(x[0].y, x[1].y));
        st2(b + 2 * kTile, make_double2(x[0].z, x[1].z));
        st2(b + 3 * kTile, make_double2(p[0].x, p[1].x));
        st2(b + 4 * kTile, make_double2(p[0].y, p[1].y));
        st2(b + 5 * kTile, make_double2(p[0].z, p[1].z));
    }
}

__device__ __forceinline__ void slot_write(double* xs, double* ps, int pitch, const Vec3& x, const Vec3& p) {
    xs[0] = x.x; xs[pitch] = x.y; xs[2 * pitch] = x.z;
    ps[0] = p.x; ps[pitch] = p.y; ps[2 * pitch] = p.z;
}

// The ring driver.  fast(x, p) -> bool updates one particle in registers;
// recover(i, xs, ps, pitch) recomputes particle i from global memory (the tile
// has not been stored yet) and writes it into its stage slot.
// Shared-memory header (kSmemHeader = 128 B): grid barrier at byte 0, full[s] at 16 + 8 s, done[s] at 64 + 8 s, s < 6.
constexpr int kRingMaxStages = 6;
__device__ __forceinline__ uint64_t* ring_full(unsigned char* smem) { return reinterpret_cast<uint64_t*>(smem + 16); }
__device__ __forceinline__ uint64_t* ring_done(unsigned char* smem) { return reinterpret_cast<uint64_t*>(smem + 64); }
constexpr uint32_t kMarginSlot = 112;      // 8 bytes of the header: window_margin()'s block reduction (gather_math.cuh)
static_assert(64 + 8 * kRingMaxStages <= kMarginSlot && kMarginSlot + 8 <= kSmemHeader, "mbarriers must fit the shared-memory header");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// observe(p[2]) sees the thread's two momenta AS LOADED, before `fast` updates them (the fused moment diagnostic).
template <bool AOS, typename Fast, typename Recover, typename Observe>
__device__ __forceinline__ void tma_ring(const ParticleArgs& P, const RingArgs& ra, unsigned char* smem, Fast fast,
                                         Recover recover, Observe observe) {
    uint64_t* full = ring_full(smem);
    uint64_t* done = ring_done(smem);
    unsigned char* tiles = smem + ra.tiles_off;
    const int nst = ra.nstages;
    const int64_t t0 = blockIdx.x, tstep = gridDim.x;
    int s = 0;
    uint32_t parity = 0;
    if (threadIdx.x >= kConsumers) {
        // ---------------- producer warp: one elected lane drives the TMA ----------------
        if (threadIdx.x == kConsumers) {
            const bool hint = !AOS && ra.stream_hint;
            const uint64_t policy = l2_evict_first();
            for (int64_t t = t0; t < ra.ntiles; t += tstep) {
                mbar_wait(&done[s], parity);                    // every consumer warp released the stage
                if (hint) tile_store_hint(P, t, tiles + s * kTileBytes, policy);
                else tile_store<AOS>(P, t, tiles + s * kTileBytes);
                const int64_t tn = t + nst * tstep;             // the tile that reuses this stage
                if (tn < ra.ntiles) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores done READING smem
                    if (hint) tile_load_hint(P, tn, tiles + s * kTileBytes, &full[s], policy);
                    else tile_load<AOS>(P, tn, tiles + s * kTileBytes, &full[s]);
                }
                if (++s == nst) { s = 0; parity ^= 1u; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    // -------------------- consumer warps --------------------
    for (int64_t t = t0; t < ra.ntiles; t += tstep) {
        mbar_wait(&full[s], parity);
        if (ra.consumer_delay_ns) __nanosleep(ra.consumer_delay_ns);
        double* stage = reinterpret_cast<double*>(tiles + s * kTileBytes);
        Vec3 x[2], p[2];
        stage_read<AOS>(stage, x, p);
        observe(p);
        const bool ok0 = fast(x[0], p[0]);
        const bool ok1 = fast(x[1], p[1]);
        stage_write<AOS>(stage, x, p);
        if (__builtin_expect(!(ok0 && ok1), 0)) {               // recovery overwrites the particle's slot
            const int64_t i = t * kTile + 2 * threadIdx.x;
            if (!ok0) recover(i, Slot<AOS>::x(stage, 0), Slot<AOS>::p(stage, 0), Slot<AOS>::pitch);
            if (!ok1) recover(i + 1, Slot<AOS>::x(stage, 1), Slot<AOS>::p(stage, 1), Slot<AOS>::pitch);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);     // one arrival per consumer warp
        if (++s == nst) { s = 0; parity ^= 1u; }
    }
}

template <bool AOS, typename Fast, typename Recover>
__device__ __forceinline__ void tma_ring(const ParticleArgs& P, const RingArgs& ra, unsigned char* smem, Fast fast,
                                         Recover recover) {
    tma_ring<AOS>(P, ra, smem, fast, recover, [](const Vec3 (&)[2]) {});
}

// barrier init (before the CTA-wide sync) and the prologue fills (after it)
__device__ __forceinline__ void ring_init(const RingArgs& ra, unsigned char* smem) {
    if (threadIdx.x == 0)
        for (int s = 0; s < ra.nstages; ++s) {
            mbar_init(&ring_full(smem)[s], 1);
            mbar_init(&ring_done(smem)[s], kConsumerWarps);
        }
}

template <bool AOS>
__device__ __forceinline__ void ring_prologue(const ParticleArgs& P, const RingArgs& ra, unsigned char* smem) {
    if (threadIdx.x == kConsumers) {
        for (int s = 0; s < ra.nstages; ++s) {
            const int64_t t = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(s) * gridDim.x;
            if (t >= ra.ntiles) continue;
            if (!AOS && ra.stream_hint) tile_load_hint(P, t, smem + ra.tiles_off + s * kTileBytes, &ring_full(smem)[s], l2_evict_first());
            else tile_load<AOS>(P, t, smem + ra.tiles_off + s * kTileBytes, &ring_full(smem)[s]);
        }
    }
}

}  // namespace tp
