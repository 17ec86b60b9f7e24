// Microbenchmark: cp.async.bulk.tensor.2d tile::gather4 (sm_100a) -- how many 64-byte table rows per cycle can one
// SM gather from random row indices of an L2-resident 2-D tensor into shared memory?  (Design input for the
// large-grid field gather: profiles/r1_large_grid.md.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather4_rate tma_gather4_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int kRowDoubles = 8;
constexpr int kSlot = 4 * kRowDoubles * 8;       // 256 B per gather4

template <int OPS>
__global__ void __launch_bounds__(512, 1) gather4_kernel(const __grid_constant__ CUtensorMap tmap, int64_t nrec, int iters,
                                                         int threads_active, long long* cycles, unsigned long long* errors) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    unsigned char* slots = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar[warp])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint64_t rng = 0x9E3779B97F4A7C15ull * (blockIdx.x * 512 + threadIdx.x + 1);
    uint32_t parity = 0;
    unsigned long long bad = 0;
    const long long t0 = clock64();
    if (threadIdx.x < threads_active) {
        for (int it = 0; it < iters; ++it) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp])), "r"(kSlot * OPS) : "memory");
            int rows[OPS][4];
#pragma unroll
            for (int k = 0; k < OPS; ++k) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                    rows[k][r] = static_cast<int>((rng >> 20) % static_cast<uint64_t>(nrec));
                }
                unsigned char* dst = slots + (static_cast<size_t>(k) * 256 + threadIdx.x) * kSlot;
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                        smem_u32(dst)),
                    "l"(&tmap), "r"(0), "r"(rows[k][0]), "r"(rows[k][1]), "r"(rows[k][2]), "r"(rows[k][3]), "r"(smem_u32(&bar[warp]))
                    : "memory");
            }
            uint32_t done = 0;
            for (int spin = 0; spin < (1 << 22) && !done; ++spin)
                asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                             : "=r"(done) : "r"(smem_u32(&bar[warp])), "r"(parity) : "memory");
            if (!done) { bad |= 1ull << 63; break; }              // never completed: report instead of hanging
            parity ^= 1u;
#pragma unroll
            for (int k = 0; k < OPS; ++k) {
                const double* s = reinterpret_cast<const double*>(slots + (static_cast<size_t>(k) * 256 + threadIdx.x) * kSlot);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (s[r * kRowDoubles] != static_cast<double>(rows[k][r]) || s[r * kRowDoubles + 7] != -static_cast<double>(rows[k][r])) ++bad;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(errors, bad);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int64_t nrec = argc > 1 ? atoll(argv[1]) : (1 << 18);
    const int box_rows = argc > 2 ? atoi(argv[2]) : 1;
    const int iters = 200;
    double* table; long long* cycles; unsigned long long* errors;
    cudaMalloc(&table, nrec * 64);
    {
        double* h = (double*)malloc(nrec * 64);
        for (int64_t j = 0; j < nrec; ++j) { for (int c = 0; c < 8; ++c) h[j * 8 + c] = 0.5 * c; h[j * 8] = (double)j; h[j * 8 + 7] = -(double)j; }
        cudaMemcpy(table, h, nrec * 64, cudaMemcpyHostToDevice);
        free(h);
    }
    cudaMalloc(&cycles, 148 * 8); cudaMalloc(&errors, 8); cudaMemset(errors, 0, 8);
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled entry point\n"); return 1;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {kRowDoubles, (cuuint64_t)nrec};
    const cuuint64_t gstride[1] = {64};
    const cuuint32_t box[2] = {kRowDoubles, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, table, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d (box %d x %d)\n", (int)rc, kRowDoubles, box_rows);
    if (rc != CUDA_SUCCESS) return 1;
    const size_t smem = 1024 + 512 * kSlot;
    cudaFuncSetAttribute(gather4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(gather4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int active : {32, 128, 256, 480}) {
        for (int ops = 1; ops <= (active <= 256 ? 2 : 1); ++ops) {
            for (int rep = 0; rep < 2; ++rep) {
                if (ops == 1) gather4_kernel<1><<<148, 512, smem>>>(tmap, nrec, iters, active, cycles, errors);
                else gather4_kernel<2><<<148, 512, smem>>>(tmap, nrec, iters, active, cycles, errors);
                cudaError_t e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[148]; unsigned long long herr = 0;
            cudaMemcpy(h, cycles, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemcpy(&herr, errors, 8, cudaMemcpyDeviceToHost);
            double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
            const double nrows = double(iters) * active * ops * 4;
            printf("table %.0f MB  threads %3d  gather4/thread/iter %d : %.0f cycles/iter, %.2f cycles per 64-B row per SM, errors %llx\n",
                   nrec * 64 / 1e6, active, ops, mean / iters, mean / nrows, herr);
            if (herr) return 2;
        }
    }
    return 0;
}
