// The particle ring of tma_ring.cuh with a per-tile WINDOW of node records next to
// each tile: the fused gather for grids too large for shared memory (BASELINE
// configs[2]: 2^20+1 nodes, 64 MB of records) when the particles are ordered by
// cell (tp.sort_by_cell).
//
// A tile of 960 cell-ordered particles touches a handful of CONSECUTIVE cells, so
// its field data is one short contiguous run of 64-byte node records
// {r, Ex, Ey, Ez, Bx, By, Bz, 0}.  Round 1 gathered those records per particle with
// four 256-bit global loads (L1 wavefront / latency bound: 72-77 % of the HBM peak,
// 1.66x DRAM read traffic in random order).  Here the producer lane fetches the
// run ONCE per tile with one cp.async.bulk into the stage, completing on the same
// `full` mbarrier as the particle slices, and the consumers gather from shared
// memory (lanes of a warp sit in the same or neighbouring cells: broadcast loads).
//
// Which cells a tile needs is known one step ahead: after pushing, every consumer
// warp reduces the cell range its particles moved INTO (redux.sync min / max, one
// shared atomic pair per warp), and the producer stores the tile's range to a
// per-tile bounds array in global memory (8 B per 960 particles).  The next call
// reads it before issuing the tile's loads.  The bounds are a HINT, never trusted:
// a particle whose cell is outside the staged window (stale bounds after the caller
// re-ordered or re-initialised the particles, a window larger than the stage, the
// very first call with zeroed bounds) reads its records from global memory / L2
// like the round-1 kernel -- same arithmetic, same bits -- and the bounds written
// by that step are exact again.  Particles in random order therefore run at the L2
// path's speed and cost nothing extra but the empty-window bookkeeping.
#pragma once
#include "tma_ring.cuh"

namespace tp {

constexpr uint32_t kWinMetaOff = kSmemHeader;          // per stage: {jlo, ncells, next_min, next_max}, 16 B each
constexpr uint32_t kWinFront = 2 * kSmemHeader;        // barriers + meta
constexpr uint32_t kRecBytes = kNodePitchWin * 8;      // node records {r, E, B, 0} + pad, 80 B: formulas A and C
constexpr uint32_t kCellBytes = kCellDoubles * 8;      // cell records {x0, x1, y0[6], slope[6]} + pad, 144 B: formula B
static_assert((kRecBytes / 16) % 2 == 1 && (kCellBytes / 16) % 2 == 1, "odd pitch in 16-byte chunks: bank-conflict-free");

struct WindowArgs {
    int2* bounds;               // per tile {first cell, number of cells} the tile's particles sit in; {*, 0} = unknown
    uint32_t win_off;           // byte offset of stage 0's window in dynamic shared memory (128-B aligned)
    uint32_t win_bytes;         // window capacity per stage (multiple of 128)
};

// the staged window of the tile being computed, as the consumers see it
struct WinView {
    uint32_t base;              // shared-space address of record `wlo`
    int jlo, ncells;            // cells [jlo, jlo + ncells) are covered (ncells == 0: nothing staged)
    int wlo;                    // first staged record
};

__device__ __forceinline__ void lds_record(uint32_t addr, double (&v)[8]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(addr + 16));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[4]), "=d"(v[5]) : "r"(addr + 32));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[6]), "=d"(v[7]) : "r"(addr + 48));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// The two nodes of the cells of BOTH particles of a thread: from the staged window when the window covers every
// particle of the warp (the cell-ordered case: one uniform branch per warp and tile, broadcast-friendly LDS), else
// from the global record table with all eight 256-bit loads of the thread issued back to back (the round-1 access
// pattern: the loads of the two particles overlap each other's latency).  Same values either way.
template <int FORMULA>
__device__ __forceinline__ void fetch_records_smem(const RecView& rv, const WinView& w, int j, NodePair& np) {
    np.j = j;
    const uint32_t a0 = w.base + static_cast<uint32_t>(j - w.wlo) * kRecBytes;
    double a[8], b[8];
    lds_record(a0, a);
    lds_record(a0 + kRecBytes, b);
    np.x0 = a[0]; np.x1 = b[0];
    if (FORMULA == TP_INTERP_A) {           // the window holds records jlo-1 .. jhi+2 (clamped to the grid)
        np.xm = lds_f64(w.base + static_cast<uint32_t>(max(j - 1, 0) - w.wlo) * kRecBytes);
        np.xp = lds_f64(w.base + static_cast<uint32_t>(min(j + 2, rv.ng - 1) - w.wlo) * kRecBytes);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { np.y0[k] = a[1 + k]; np.y1[k] = b[1 + k]; }
}

template <int FORMULA>
__device__ __forceinline__ void fetch_pair_window(const RecView& rv, const WinView& w, double xa, double xb,
                                                  NodePair& na, NodePair& nb) {
    const int ja = guess_cell(xa, rv.r_first, rv.inv_dr, rv.ng);
    const int jb = guess_cell(xb, rv.r_first, rv.inv_dr, rv.ng);
    const bool in = static_cast<unsigned>(ja - w.jlo) < static_cast<unsigned>(w.ncells) &&
                    static_cast<unsigned>(jb - w.jlo) < static_cast<unsigned>(w.ncells);
    if (__all_sync(0xffffffffu, in)) {
        fetch_records_smem<FORMULA>(rv, w, ja, na);
        fetch_records_smem<FORMULA>(rv, w, jb, nb);
    } else {
        fetch_records<FORMULA, kNodePitchWin>(rv, xa, na);
        fetch_records<FORMULA, kNodePitchWin>(rv, xb, nb);
    }
}

// The same for cell records (formula B): one 128-byte record per particle, seven LDS.128 from the window or four
// 256-bit loads of one cache line from the global table.
__device__ __forceinline__ void lds_cell(uint32_t addr, double (&c)[16]) {
#pragma unroll
    for (int q = 0; q < 7; ++q)
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(c[2 * q]), "=d"(c[2 * q + 1]) : "r"(addr + 16 * q));
    c[14] = 0.0; c[15] = 0.0;
}

__device__ __forceinline__ void fetch_pair_cells(const RecView& rv, const WinView& w, double xa, double xb,
                                                 double (&ca)[16], double (&cb)[16]) {
    const int ja = guess_cell(xa, rv.r_first, rv.inv_dr, rv.ng);
    const int jb = guess_cell(xb, rv.r_first, rv.inv_dr, rv.ng);
    const bool in = static_cast<unsigned>(ja - w.jlo) < static_cast<unsigned>(w.ncells) &&
                    static_cast<unsigned>(jb - w.jlo) < static_cast<unsigned>(w.ncells);
    if (__all_sync(0xffffffffu, in)) {
        lds_cell(w.base + static_cast<uint32_t>(ja - w.wlo) * kCellBytes, ca);
        lds_cell(w.base + static_cast<uint32_t>(jb - w.wlo) * kCellBytes, cb);
    } else {
        ld_cell(rv.rec + static_cast<int64_t>(ja) * kCellDoubles, ca, rv.policy);
        ld_cell(rv.rec + static_cast<int64_t>(jb) * kCellDoubles, cb, rv.policy);
    }
}

__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Ring driver.  fast(x[2], p[2], win, ok0, ok1) updates the thread's two particles in registers; recover as in tma_ring.
// CELLS: the record table holds 128-byte cell records (cells jlo .. jhi are staged) instead of 64-byte node records
// (nodes jlo-1 .. jhi+2, clamped: formula A also looks at the outer neighbours).
// observe(p[2]) sees the thread's two momenta AS LOADED, before `fast` updates them (the fused moment diagnostic
// accumulates there; the other instances pass a no-op).
template <bool AOS, bool CELLS, typename Fast, typename Recover, typename Observe>
__device__ __forceinline__ void tma_ring_window(const ParticleArgs& P, const RingArgs& ra, const WindowArgs& wa,
                                                const RecView& rv, int axis, unsigned char* smem, Fast fast,
                                                Recover recover, Observe observe) {
    uint64_t* full = ring_full(smem);
    uint64_t* done = ring_done(smem);
    int4* meta = reinterpret_cast<int4*>(smem + kWinMetaOff);
    unsigned char* tiles = smem + ra.tiles_off;
    unsigned char* wins = smem + wa.win_off;
    const int nst = ra.nstages;
    constexpr uint32_t kRec = CELLS ? kCellBytes : kRecBytes;
    const int cap_records = static_cast<int>(wa.win_bytes / kRec);
    const int64_t t0 = blockIdx.x, tstep = gridDim.x;
    int s = 0;
    uint32_t parity = 0;
    if (threadIdx.x >= kConsumers) {
        // ---------------- producer warp: one elected lane drives the TMA ----------------
        if (threadIdx.x == kConsumers) {
            const uint64_t pol_stream = l2_evict_first(), pol_keep = l2_evict_last();
            const bool hint = !AOS && ra.stream_hint;
            // load tile t and (if its bounds are known and fit) its record window into stage st
            auto fill = [&](int64_t t, int st, int2 b) {
                int wlo = 0, nrec = 0;
                const int64_t jhi = static_cast<int64_t>(b.x) + b.y - 1;       // (64-bit: the words may be garbage)
                if (b.y > 0 && b.x >= 0 && jhi <= rv.ng - 2) {
                    wlo = CELLS ? b.x : max(b.x - 1, 0);
                    const int whi = CELLS ? static_cast<int>(jhi) : min(static_cast<int>(jhi) + 2, rv.ng - 1);
                    nrec = whi - wlo + 1;
                    if (nrec > cap_records) nrec = 0;                           // too wide for the stage
                }
                meta[st] = make_int4(b.x, nrec ? b.y : 0, 0x7fffffff, static_cast<int>(0x80000000u));
                if (nrec) {
                    mbar_expect_tx_only(&full[st], static_cast<uint32_t>(nrec) * kRec);
                    tma_bulk_g2s_hint(wins + static_cast<size_t>(st) * wa.win_bytes,
                                      rv.rec + static_cast<int64_t>(wlo) * (kRec / 8),
                                      static_cast<uint32_t>(nrec) * kRec, &full[st], pol_keep);
                }
                if (hint) tile_load_hint(P, t, tiles + st * kTileBytes, &full[st], pol_stream);
                else tile_load<AOS>(P, t, tiles + st * kTileBytes, &full[st]);
            };
            for (int i = 0; i < nst; ++i) {
                const int64_t t = t0 + i * tstep;
                if (t < ra.ntiles) fill(t, i, wa.bounds[t]);
            }
            for (int64_t t = t0; t < ra.ntiles; t += tstep) {
                const int64_t tn = t + nst * tstep;             // the tile that reuses this stage
                int2 bn = make_int2(0, 0);
                if (tn < ra.ntiles) bn = wa.bounds[tn];         // in flight while we wait for the consumers
                mbar_wait(&done[s], parity);                    // every consumer warp released the stage
                const int4 m = meta[s];
                wa.bounds[t] = (m.w >= m.z) ? make_int2(m.z, m.w - m.z + 1) : make_int2(0, 0);   // for the next call
                if (hint) tile_store_hint(P, t, tiles + s * kTileBytes, pol_stream);
                else tile_store<AOS>(P, t, tiles + s * kTileBytes);
                if (tn < ra.ntiles) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores done READING smem
                    fill(tn, s, bn);
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
        double* stage = reinterpret_cast<double*>(tiles + s * kTileBytes);
        const int4 m = meta[s];
        WinView w;
        w.jlo = m.x; w.ncells = m.y; w.wlo = CELLS ? m.x : max(m.x - 1, 0);
        w.base = smem_u32(wins + static_cast<size_t>(s) * wa.win_bytes);
        Vec3 x[2], p[2];
        stage_read<AOS>(stage, x, p);
        observe(p);
        bool ok0, ok1;
        fast(x, p, w, ok0, ok1);
        stage_write<AOS>(stage, x, p);
        // the cells this warp's particles moved into: next call's window (particles that left the fast path are
        // served from global memory next time, whatever the window)
        const int j0 = guess_cell(axis_of(x[0], axis), rv.r_first, rv.inv_dr, rv.ng);
        const int j1 = guess_cell(axis_of(x[1], axis), rv.r_first, rv.inv_dr, rv.ng);
        int lo = min(ok0 ? j0 : 0x7fffffff, ok1 ? j1 : 0x7fffffff);
        int hi = max(ok0 ? j0 : static_cast<int>(0x80000000u), ok1 ? j1 : static_cast<int>(0x80000000u));
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&meta[s].z, lo);
            atomicMax(&meta[s].w, hi);
        }
        if (__builtin_expect(!(ok0 && ok1), 0)) {               // recovery overwrites the particle's slot
            const int64_t i = t * kTile + 2 * threadIdx.x;
            if (!ok0) recover(i, Slot<AOS>::x(stage, 0), Slot<AOS>::p(stage, 0), Slot<AOS>::pitch);
            if (!ok1) recover(i + 1, Slot<AOS>::x(stage, 1), Slot<AOS>::p(stage, 1), Slot<AOS>::pitch);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);     // one arrival per consumer warp (release: the atomics too)
        if (++s == nst) { s = 0; parity ^= 1u; }
    }
}

}  // namespace tp
