// Microbenchmark: how fast can one SM issue small (128 B) cp.async.bulk gathers from random
// addresses of an L2-resident table into per-thread shared-memory slots?  (Design input for the
// large-grid gather: profiles/r1_large_grid.md.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather_rate tma_gather_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int OPS>
__global__ void __launch_bounds__(512, 1) gather_kernel(const double* table, int64_t nrec, int iters, int threads_active,
                                                        long long* cycles, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);            // one mbarrier per warp
    unsigned char* slots = smem + 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar[warp])));
    __syncthreads();
    uint64_t rng = 0x9E3779B97F4A7C15ull * (blockIdx.x * 512 + threadIdx.x + 1);
    double acc = 0.0;
    uint32_t parity = 0;
    const long long t0 = clock64();
    if (threadIdx.x < threads_active) {
        for (int it = 0; it < iters; ++it) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp])), "r"(128 * OPS) : "memory");
#pragma unroll
            for (int k = 0; k < OPS; ++k) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const int64_t j = static_cast<int64_t>((rng >> 20) % static_cast<uint64_t>(nrec - 1));
                unsigned char* dst = slots + (static_cast<size_t>(k) * 512 + threadIdx.x) * 144;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                             "l"(table + j * 8), "r"(128), "r"(smem_u32(&bar[warp]))
                             : "memory");
            }
            asm volatile(
                "{\n.reg .pred P1;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
                    smem_u32(&bar[warp])),
                "r"(parity)
                : "memory");
            parity ^= 1u;
#pragma unroll
            for (int k = 0; k < OPS; ++k)
                acc += *reinterpret_cast<const double*>(slots + (static_cast<size_t>(k) * 512 + threadIdx.x) * 144);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 12345.678) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int64_t nrec = argc > 1 ? atoll(argv[1]) : (1 << 18);      // 64 B records
    const int iters = 200;
    double* table; long long* cycles; double* sink;
    cudaMalloc(&table, nrec * 64); cudaMemset(table, 0, nrec * 64);
    cudaMalloc(&cycles, 148 * 8); cudaMalloc(&sink, 8);
    const size_t smem = 256 + 2 * 512 * 144;
    cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int active : {32, 128, 480}) {
        for (int ops = 1; ops <= 2; ++ops) {
            for (int rep = 0; rep < 2; ++rep) {
                if (ops == 1) gather_kernel<1><<<148, 512, smem>>>(table, nrec, iters, active, cycles, sink);
                else gather_kernel<2><<<148, 512, smem>>>(table, nrec, iters, active, cycles, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[148];
            cudaMemcpy(h, cycles, sizeof(h), cudaMemcpyDeviceToHost);
            double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
            const double nops = double(iters) * active * ops;
            printf("table %.0f MB  active threads %3d  ops/thread/iter %d : %.0f cycles/iter, %.2f cycles per 128-B gather per SM\n",
                   nrec * 64 / 1e6, active, ops, mean / iters, mean / nops);
        }
    }
    return 0;
}
