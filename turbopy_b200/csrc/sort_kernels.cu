// Particle reordering by grid cell (additive; the reference never reorders).
//
// The fused gather reads two node records per particle.  On grids too large
// for shared memory (config 3: 2^20+1 nodes, 64 MB of records) particles in
// random order make every lane of a record load touch its own 128-byte line:
// the kernel becomes bound by L1 wavefronts (one per lane and load, ~2 cycles
// each), not by HBM or L2 (profiles/r1_large_grid.md).  With the particles
// ordered by cell the lanes of a warp share lines again.  This file builds the
// STABLE permutation that orders the particles by
//
//   key(x) = clamp(trunc((x - r_first) * inv_dr), 0, ncells-1) >> shift      (two rounded fp64 ops, no FMA)
//
// -- ties keep their original order, so the result is reproducible and equals
// numpy.argsort(key, kind="stable") (oracle/sort.py) -- and applies it to the
// particle arrays in place.
//
// LSD radix sort over 12-bit digits of the key; each pass is a counting sort
// with one warp per contiguous chunk of the current order:
//   K1  per-warp histogram            counts[digit][warp]
//   K2  row scans + digit bases       exclusive scan in (digit, warp) order
//   K3  stable scatter of indices     out[dst] = in[i]   (warp match_any ranking, in order)
// then  K4  per component: tmp[i] = a[perm[i]];  K5  a[i] = tmp[i].
// Keys are recomputed from x in every kernel instead of being stored.  Measured alternatives (16 Mi particles):
// 10-bit digits with 6 CTAs per SM are slower (1.41 vs 1.00 ms for the scatters: the strided per-warp rows of the
// count matrix grow), as is gathering the three components of an array in one kernel (see K4 below).
#include "common.cuh"

#include <algorithm>

namespace tp {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortDigitBits = 12;
constexpr int kSortMaxBins = 1 << kSortDigitBits;
constexpr int kSortCtasPerSm = 1;

struct SortArgs {
    const double* x;            // the coordinate the gather runs along
    int64_t sn;                 // its element stride
    int64_t n;
    double r_first, inv_dr;
    int64_t ncells;
    int shift, nbins;           // key = cell >> shift; nbins = digits of this pass
    int digit_shift;            // this pass sorts by (key >> digit_shift) & (kSortMaxBins - 1)
    const int64_t* in_perm;     // current order (nullptr = identity)
    int64_t nwarps;             // chunks
    int64_t chunk;              // particles per chunk (multiple of 32)
    uint32_t* counts;           // [nbins][nwarps]
    uint32_t* bin_base;         // [nbins]
    int64_t* perm;
};

__device__ __forceinline__ int bin_of(const SortArgs& a, double x) {
    const double t = (x - a.r_first) * a.inv_dr;
    int64_t c = t >= 0.0 ? static_cast<int64_t>(fmin(t, 9.0e18)) : 0;       // NaN and negatives -> cell 0; trunc == floor for t >= 0
    c = min(c, a.ncells - 1);
    return static_cast<int>(((c >> a.shift) >> a.digit_shift) & (kSortMaxBins - 1));
}
__device__ __forceinline__ int64_t source_of(const SortArgs& a, int64_t i) { return a.in_perm ? a.in_perm[i] : i; }

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const __grid_constant__ SortArgs a) {
    extern __shared__ uint32_t cnt[];                     // [kSortWarps][nbins]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* mine = cnt + warp * a.nbins;
    for (int b = lane; b < a.nbins; b += 32) mine[b] = 0;
    __syncwarp();
    const int64_t w = static_cast<int64_t>(blockIdx.x) * kSortWarps + warp;
    if (w >= a.nwarps) return;
    const int64_t lo = w * a.chunk, hi = min(a.n, lo + a.chunk);
    for (int64_t i0 = lo + lane; i0 < hi; i0 += 32 * 4) {        // four independent (perm ->) x loads in flight per lane
        int64_t src[4];
        double xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) src[u] = i0 + 32 * u < hi ? source_of(a, i0 + 32 * u) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = a.x[src[u] * a.sn];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + 32 * u < hi) atomicAdd(&mine[bin_of(a, xv[u])], 1u);
    }
    __syncwarp();
    for (int b = lane; b < a.nbins; b += 32) a.counts[static_cast<int64_t>(b) * a.nwarps + w] = mine[b];
}

// one CTA per bin: exclusive scan of its row of per-warp counts, in place; row total -> bin_base[bin]
__global__ void __launch_bounds__(kSortThreads) sort_rows_kernel(const __grid_constant__ SortArgs a) {
    __shared__ uint32_t wsum[kSortWarps];
    __shared__ uint32_t carry;
    uint32_t* row = a.counts + static_cast<int64_t>(blockIdx.x) * a.nwarps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < a.nwarps; base += kSortThreads) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < a.nwarps ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t before = carry;
        for (int k = 0; k < warp; ++k) before += wsum[k];
        if (i < a.nwarps) row[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == kSortThreads - 1) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.bin_base[blockIdx.x] = carry;
}

// one CTA: exclusive scan of the bin totals, in place
__global__ void __launch_bounds__(kSortThreads) sort_bins_kernel(const __grid_constant__ SortArgs a) {
    __shared__ uint32_t wsum[kSortWarps];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < a.nbins; base += kSortThreads) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < a.nbins ? a.bin_base[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t before = carry;
        for (int k = 0; k < warp; ++k) before += wsum[k];
        if (i < a.nbins) a.bin_base[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == kSortThreads - 1) carry = before + inc;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const __grid_constant__ SortArgs a) {
    extern __shared__ uint32_t cnt[];                     // [kSortWarps][nbins]: next free slot of each bin for this warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w = static_cast<int64_t>(blockIdx.x) * kSortWarps + warp;
    if (w >= a.nwarps) return;
    uint32_t* mine = cnt + warp * a.nbins;
    for (int b = lane; b < a.nbins; b += 32) mine[b] = a.bin_base[b] + a.counts[static_cast<int64_t>(b) * a.nwarps + w];
    __syncwarp();
    const int64_t lo = w * a.chunk, hi = min(a.n, lo + a.chunk);
    constexpr int kAhead = 4;                                    // iterations whose (dependent) loads are issued together
    for (int64_t group = lo; group < hi; group += 32 * kAhead) {  // chunk is a multiple of 32: whole warp iterates together
        int64_t src[kAhead];
        double xv[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int64_t i = group + 32 * u + lane;
            src[u] = i < hi ? source_of(a, i) : 0;
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) xv[u] = a.x[src[u] * a.sn];        // src = 0 for dead lanes: a valid address
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int64_t i = group + 32 * u + lane;
            if (group + 32 * u >= hi) break;                     // warp-uniform
            const bool live = i < hi;
            const int b = live ? bin_of(a, xv[u]) : -1 - lane;   // dead lanes: unique keys, matched with nobody
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            const int leader = __ffs(peers) - 1;
            const int rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t slot = 0;
            if (live && lane == leader) {                        // one lane per distinct bin: no two leaders share a counter
                slot = mine[b];
                mine[b] = slot + __popc(peers);
            }
            slot = __shfl_sync(0xffffffffu, slot, leader);
            if (live) a.perm[static_cast<int64_t>(slot) + rank] = src[u];
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(256) permute_gather_kernel(const double* __restrict__ src, int64_t sn,
                                                             const int64_t* __restrict__ perm, double* __restrict__ tmp,
                                                             int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        tmp[i] = src[perm[i] * sn];
}

__global__ void __launch_bounds__(256) permute_copy_kernel(const double* __restrict__ tmp, double* __restrict__ dst, int64_t sn,
                                                           int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i * sn] = tmp[i];
}

struct SortPlan {
    int shift, passes;
    int64_t nkeys;
    int64_t nwarps, chunk;
    size_t counts_bytes, bins_bytes, perm_bytes, tmp_bytes;
};

static int plan_sort(int64_t n, int64_t ncells, int cells_per_bin_log2, SortPlan& pl) {
    if (n < 0 || n >= (1ll << 32) || ncells < 1 || ncells > (1ll << 40) || cells_per_bin_log2 < 0 || cells_per_bin_log2 > 40)
        return TP_E_ARG;
    pl.shift = cells_per_bin_log2;
    pl.nkeys = ((ncells - 1) >> pl.shift) + 1;
    pl.passes = 1;
    while ((pl.nkeys - 1) >> (kSortDigitBits * pl.passes)) ++pl.passes;
    const DeviceProps& d = device_props();
    const int64_t want = static_cast<int64_t>(std::max(d.sm_count, 1)) * kSortWarps * kSortCtasPerSm;
    int64_t chunk = (std::max<int64_t>(n, 1) + want - 1) / want;
    chunk = std::max<int64_t>(1024, (chunk + 31) / 32 * 32);
    pl.chunk = chunk;
    pl.nwarps = std::max<int64_t>(1, (n + chunk - 1) / chunk);
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t bins = static_cast<size_t>(std::min<int64_t>(pl.nkeys, kSortMaxBins));
    pl.counts_bytes = up(sizeof(uint32_t) * bins * static_cast<size_t>(pl.nwarps));
    pl.bins_bytes = up(sizeof(uint32_t) * bins);
    pl.perm_bytes = up(sizeof(int64_t) * static_cast<size_t>(n));
    pl.tmp_bytes = up(sizeof(double) * static_cast<size_t>(n));
    return TP_OK;
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_sort_scratch_bytes(int64_t n, int64_t ncells, int cells_per_bin_log2) {
    SortPlan pl;
    if (plan_sort(n, ncells, cells_per_bin_log2, pl)) return -1;
    return static_cast<int64_t>(pl.counts_bytes + pl.bins_bytes + (pl.passes > 1 ? 2 : 1) * pl.perm_bytes + pl.tmp_bytes);
}

extern "C" int tp_sort_by_cell_f64(const tp_particles* p, int axis, double r_first, double dr, int64_t ncells,
                                   int cells_per_bin_log2, void* scratch, int64_t* perm_out, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!p || p->n < 0 || axis < 0 || axis > 2 || !(dr > 0.0)) return TP_E_ARG;
    if (p->pos_stride_n <= 0 || p->pos_stride_c <= 0 || p->mom_stride_n <= 0 || p->mom_stride_c <= 0) return TP_E_LAYOUT;
    SortPlan pl;
    int rc = plan_sort(p->n, ncells, cells_per_bin_log2, pl);
    if (rc) return rc;
    if (p->n == 0) return TP_OK;
    if (!p->pos || !p->mom || !scratch) return TP_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char* base = static_cast<unsigned char*>(scratch);
    SortArgs a;
    a.x = p->pos + axis * p->pos_stride_c; a.sn = p->pos_stride_n; a.n = p->n;
    a.r_first = r_first; a.inv_dr = 1.0 / dr;                   // IEEE on the host, like the gather's guess
    a.ncells = ncells; a.shift = pl.shift;
    a.nwarps = pl.nwarps; a.chunk = pl.chunk;
    a.counts = reinterpret_cast<uint32_t*>(base);
    a.bin_base = reinterpret_cast<uint32_t*>(base + pl.counts_bytes);
    int64_t* perm_a = reinterpret_cast<int64_t*>(base + pl.counts_bytes + pl.bins_bytes);
    int64_t* perm_b = pl.passes > 1 ? reinterpret_cast<int64_t*>(base + pl.counts_bytes + pl.bins_bytes + pl.perm_bytes) : nullptr;
    double* tmp = reinterpret_cast<double*>(base + pl.counts_bytes + pl.bins_bytes + (pl.passes > 1 ? 2 : 1) * pl.perm_bytes);
    const int blocks = static_cast<int>((pl.nwarps + kSortWarps - 1) / kSortWarps);
    const int64_t* current = nullptr;
    for (int pass = 0; pass < pl.passes; ++pass) {
        a.digit_shift = kSortDigitBits * pass;
        a.nbins = static_cast<int>(std::min<int64_t>(((pl.nkeys - 1) >> a.digit_shift) + 1, kSortMaxBins));
        a.in_perm = current;
        a.perm = (pass & 1) ? perm_b : perm_a;
        const size_t smem = sizeof(uint32_t) * kSortWarps * static_cast<size_t>(a.nbins);
        if (smem > 48 * 1024) {
            if ((rc = static_cast<int>(cudaFuncSetAttribute(sort_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))))) return rc;
            if ((rc = static_cast<int>(cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))))) return rc;
        }
        sort_hist_kernel<<<blocks, kSortThreads, smem, s>>>(a);
        sort_rows_kernel<<<a.nbins, kSortThreads, 0, s>>>(a);
        sort_bins_kernel<<<1, kSortThreads, 0, s>>>(a);
        sort_scatter_kernel<<<blocks, kSortThreads, smem, s>>>(a);
        count_launch(4);
        if ((rc = static_cast<int>(cudaGetLastError()))) return rc;
        current = a.perm;
    }
    const int pblocks = static_cast<int>(std::min<int64_t>((p->n + 255) / 256, static_cast<int64_t>(d.sm_count) * 16));
    for (int arr = 0; arr < 2; ++arr) {
        double* ptr = arr ? p->mom : p->pos;
        const int64_t sn = arr ? p->mom_stride_n : p->pos_stride_n, sc = arr ? p->mom_stride_c : p->pos_stride_c;
        // one component at a time: its 134 MB-per-16-Mi source plane is re-read from L2 by the three other particles of
        // each 32-byte sector (gathering the three planes in one kernel thrashes L2: measured 1.5x slower)
        for (int c = 0; c < 3; ++c) {
            permute_gather_kernel<<<pblocks, 256, 0, s>>>(ptr + c * sc, sn, current, tmp, p->n);
            permute_copy_kernel<<<pblocks, 256, 0, s>>>(tmp, ptr + c * sc, sn, p->n);
        }
    }
    count_launch(12);
    if ((rc = static_cast<int>(cudaGetLastError()))) return rc;
    if (perm_out)
        rc = static_cast<int>(cudaMemcpyAsync(perm_out, current, sizeof(int64_t) * static_cast<size_t>(p->n), cudaMemcpyDeviceToDevice, s));
    return rc;
}

extern "C" int tp_permute_f64(double* array, int64_t stride_n, int64_t n, const int64_t* perm, double* tmp, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (n < 0 || stride_n <= 0) return TP_E_ARG;
    if (n == 0) return TP_OK;
    if (!array || !perm || !tmp) return TP_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int pblocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(d.sm_count) * 16));
    permute_gather_kernel<<<pblocks, 256, 0, s>>>(array, stride_n, perm, tmp, n);
    permute_copy_kernel<<<pblocks, 256, 0, s>>>(tmp, array, stride_n, n);
    count_launch(2);
    return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// Local re-sort: restore the cell order of an ALREADY nearly ordered ensemble
// ---------------------------------------------------------------------------
// A cell-ordered ensemble decays slowly (config 3: a 960-particle tile spans ~10 cells after the sort and ~0.3
// cells more per step), but the windowed gather feels it quickly: once the 64 particles of a warp sit in 20
// different cells its window loads stop being broadcasts (measured: 92 % of the HBM peak right after a sort, 83 %
// 100 steps later, 66 % from step 300 on).  The global radix sort above costs 39 ms per 125 M particles -- 18 push
// steps -- because it gathers every component through a random permutation.  This kernel re-orders the particles
// INSIDE blocks of 1920 consecutive particles (two ring tiles) instead: one read and one write of the state, like a
// push step (~2.5 ms per 125 M particles).  Block b covers [offset + (b-1)*1920, offset + b*1920) (block 0 is the
// head [0, offset)); callers alternate offset = 0 / 960 from call to call, so particles also migrate across block
// boundaries over successive calls.
//
// Per block: TMA bulk loads of the six component slices into shared memory (double-buffered: the next block loads
// while this one is sorted), key = (cell - block's lowest cell, capped at 2^20-1) << 12 | index-in-block, a bitonic
// sort of the 2048 padded 32-bit keys -- the index in the low bits makes the order STABLE and unique, so the result
// equals numpy.argsort(key, kind="stable") per block (oracle/sort.py: resort_blocks) -- and a gather from shared
// memory straight into coalesced 128-bit global stores.
namespace tp {

constexpr int kRsBlock = 1920;
constexpr int kRsPad = 2048;
constexpr int kRsThreads = 512;
constexpr uint32_t kRsSlice = kRsBlock * sizeof(double);            // 15 360 B
constexpr uint32_t kRsBuf = 6 * kRsSlice;                           // 92 160 B
constexpr uint32_t kRsSmem = 128 + 2 * kRsBuf + kRsPad * sizeof(uint32_t);

struct ResortArgs {
    double* comp[6];            // x, y, z, px, py, pz component arrays (contiguous, 16-byte aligned)
    int axis;
    int64_t n, offset, nblocks;
    double r_first, inv_dr;
    int64_t ncells;
};

__device__ __forceinline__ uint32_t rs_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void rs_block_range(const ResortArgs& a, int64_t q, int64_t& first, int& count) {
    int64_t lo, hi;
    if (a.offset > 0) {
        lo = q == 0 ? 0 : a.offset + (q - 1) * kRsBlock;
        hi = q == 0 ? a.offset : lo + kRsBlock;
    } else {
        lo = q * kRsBlock; hi = lo + kRsBlock;
    }
    hi = min(hi, a.n);
    first = lo; count = static_cast<int>(hi - lo);
}

__global__ void __launch_bounds__(kRsThreads, 1) resort_blocks_kernel(const __grid_constant__ ResortArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);                 // two mbarriers
    double* bufs = reinterpret_cast<double*>(smem + 128);
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem + 128 + 2 * kRsBuf);
    __shared__ int s_min[kRsThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rs_smem_u32(&full[s])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // one elected thread streams a block's six slices into buffer s (the even part by the TMA; an odd last
    // element is copied by the same thread with a plain load, visible after the barrier below)
    auto issue = [&](int64_t q, int s) {
        int64_t first; int count;
        rs_block_range(a, q, first, count);
        const uint32_t bytes = static_cast<uint32_t>(count & ~1) * sizeof(double);
        double* dst = bufs + static_cast<size_t>(s) * 6 * kRsBlock;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rs_smem_u32(&full[s])), "r"(6 * bytes) : "memory");
        for (int c = 0; c < 6; ++c) {
            if (bytes)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 rs_smem_u32(dst + c * kRsBlock)),
                             "l"(a.comp[c] + first), "r"(bytes), "r"(rs_smem_u32(&full[s]))
                             : "memory");
            if (count & 1) dst[c * kRsBlock + count - 1] = a.comp[c][first + count - 1];
        }
    };
    int s = 0;
    uint32_t parity = 0u;                                              // bit s: phase parity of full[s]
    int64_t q = blockIdx.x;
    if (tid == 0 && q < a.nblocks) issue(q, 0);
    for (; q < a.nblocks; q += gridDim.x) {
        const int64_t qn = q + gridDim.x;
        if (tid == 0 && qn < a.nblocks) issue(qn, s ^ 1);              // buffer s^1 was released by the barrier that ended the last iteration
        // wait for this block's data
        asm volatile(
            "{\n.reg .pred P1;\nRS_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra.uni RS_DONE;\nbra.uni RS_WAIT;\nRS_DONE:\n}\n" ::"r"(
                rs_smem_u32(&full[s])),
            "r"((parity >> s) & 1u)
            : "memory");
        parity ^= 1u << s;
        __syncthreads();                                               // (the odd tail element, written by thread 0)
        int64_t first; int count;
        rs_block_range(a, q, first, count);
        const double* buf = bufs + static_cast<size_t>(s) * 6 * kRsBlock;
        const double* xs = buf + a.axis * kRsBlock;
        // cells and the block's lowest cell
        int cell[kRsPad / kRsThreads];
        int lo = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < kRsPad / kRsThreads; ++u) {
            const int i = tid + u * kRsThreads;
            cell[u] = 0x7fffffff;
            if (i < count) {
                const double t = (xs[i] - a.r_first) * a.inv_dr;
                int64_t c = t >= 0.0 ? static_cast<int64_t>(fmin(t, 9.0e18)) : 0;     // the key of tp_sort_by_cell_f64
                c = min(c, a.ncells - 1);
                cell[u] = static_cast<int>(min(c, static_cast<int64_t>(0x7ffffffe)));
                lo = min(lo, cell[u]);
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        if (lane == 0) s_min[warp] = lo;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kRsThreads / 32; ++w) lo = min(lo, s_min[w]);
#pragma unroll
        for (int u = 0; u < kRsPad / kRsThreads; ++u) {
            const int i = tid + u * kRsThreads;
            keys[i] = i < count ? (static_cast<uint32_t>(min(cell[u] - lo, (1 << 20) - 1)) << 12) | static_cast<uint32_t>(i)
                                : 0xffffffffu;
        }
        __syncthreads();
        // bitonic sort of the 2048 keys, ascending
        for (int k = 2; k <= kRsPad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int u = 0; u < kRsPad / 2 / kRsThreads; ++u) {
                    const int t = tid + u * kRsThreads;                 // compare-exchange number t of this stage
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)); // the lower index of the pair
                    const int p = i | j;
                    const uint32_t x = keys[i], y = keys[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                }
                __syncthreads();
            }
        }
        // gather from shared memory into coalesced global stores (two outputs per thread and component)
        for (int i2 = tid; 2 * i2 < count; i2 += kRsThreads) {
            const int i = 2 * i2;
            const int s0 = static_cast<int>(keys[i] & 0xfffu);
            if (i + 1 < count) {
                const int s1 = static_cast<int>(keys[i + 1] & 0xfffu);
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    st2(a.comp[c] + first + i, make_double2(buf[c * kRsBlock + s0], buf[c * kRsBlock + s1]));
            } else {
#pragma unroll
                for (int c = 0; c < 6; ++c) a.comp[c][first + i] = buf[c * kRsBlock + s0];
            }
        }
        __syncthreads();                                               // buffer s and the keys are free again
        s ^= 1;
    }
}

}  // namespace tp

extern "C" int tp_resort_blocks_f64(const tp_particles* p, int axis, double r_first, double dr, int64_t ncells,
                                    int64_t block_offset, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!p || p->n < 0 || axis < 0 || axis > 2 || !(dr > 0.0) || ncells < 1) return TP_E_ARG;
    if (block_offset < 0 || block_offset >= kRsBlock || (block_offset & 1)) return TP_E_ARG;
    if (p->n == 0) return TP_OK;
    if (!p->pos || !p->mom) return TP_E_ARG;
    // component arrays: contiguous (SoA) and 16-byte aligned, like the TMA particle ring
    if (p->pos_stride_n != 1 || p->mom_stride_n != 1 || (p->pos_stride_c & 1) || (p->mom_stride_c & 1) ||
        p->pos_stride_c < p->n || p->mom_stride_c < p->n || !aligned16(p->pos) || !aligned16(p->mom))
        return TP_E_LAYOUT;
    ResortArgs a;
    for (int c = 0; c < 3; ++c) { a.comp[c] = p->pos + c * p->pos_stride_c; a.comp[3 + c] = p->mom + c * p->mom_stride_c; }
    a.axis = axis; a.n = p->n; a.offset = std::min<int64_t>(block_offset, p->n);
    a.r_first = r_first; a.inv_dr = 1.0 / dr; a.ncells = ncells;
    const int64_t rest = p->n - a.offset;
    a.nblocks = (a.offset > 0 ? 1 : 0) + (rest + kRsBlock - 1) / kRsBlock;
    int rc = static_cast<int>(cudaFuncSetAttribute(resort_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(kRsSmem)));
    if (rc) return rc;
    const int blocks = static_cast<int>(std::min<int64_t>(a.nblocks, d.sm_count));
    resort_blocks_kernel<<<blocks, kRsThreads, kRsSmem, static_cast<cudaStream_t>(stream)>>>(a);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int64_t tp_resort_block_particles(void) { return kRsBlock; }
