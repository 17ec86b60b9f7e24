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
