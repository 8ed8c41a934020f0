// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld.32x32b.x32 per SM (development probe).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nabo_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, int inflight, unsigned long long* cycles, unsigned* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc(&tbase, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128 % 512;
    unsigned acc = 0;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + (i & 3) * 32, v);
        if (inflight == 1 || (i % inflight) == inflight - 1) ptx::tmem_ld_wait();
        acc += v[i & 31];
    }
    ptx::tmem_ld_wait();
    unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tbase, 512);
}

int main() {
    unsigned long long* cyc; unsigned* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    for (int warps : {1, 4, 8, 12, 16}) {
        for (int inflight : {1, 4}) {
            const int iters = 4096;
            ldtm_kernel<<<148, warps * 32>>>(iters, inflight, cyc, sink);
            cudaDeviceSynchronize();
            ldtm_kernel<<<148, warps * 32>>>(iters, inflight, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double bytes = (double)warps * iters * 32 * 32 * 4;
            printf("warps %2d inflight %d: %llu cycles -> %.1f B/cycle/SM  (%.1f cycles per LDTM.x32 per warp) %s\n", warps, inflight,
                   h[0], bytes / h[0], (double)h[0] / iters, cudaGetErrorString(e));
        }
    }
    return 0;
}
