// Microbenchmark: cycles per tcgen05.mma (M=128, kind::f16) as a function of N and of the A source
// (shared memory vs tensor memory).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I.. mma_rate.cu
#include <cstdio>
#include "../constrained-model-based-policy-optimization_b200/csrc/tc_common.cuh"
using namespace tc;
void cmbpo_set_error(const char*, ...) {}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) {
        const uint32_t idesc = idesc_f16(0, N);
        const uint64_t dA = smem_desc_sw128(smem_u32(smem));
        const uint64_t dB = smem_desc_sw128(smem_u32(smem + 16384));
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (TS) mma_f16_ts(256u, (uint32_t)(ks * 8), dB + 2 * ks, idesc, 1);
                    else mma_f16(256u, dA + 2 * ks, dB + 2 * ks, idesc, 1);
                }
            }
            mma_commit(&bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(0, 512);
}

template <int N, bool TS>
void run(const char* name) {
    unsigned long long* d; cudaMalloc(&d, 148 * 16);
    auto k = rate_kernel<N, TS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    const int iters = 2000;
    for (int grid : {1, 148}) {
        k<<<grid, 128, 60000>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[296];
        cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
        double issue = 0, total = 0;
        for (int b = 0; b < grid; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; }
        printf("%-10s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", name, N, grid,
               issue / grid / (iters * 4.0), total / grid / (iters * 4.0), cudaGetErrorString(e));
    }
    cudaFree(d);
}

int main() {
    run<64, false>("SS"); run<128, false>("SS"); run<256, false>("SS");
    run<64, true>("TS"); run<128, true>("TS"); run<256, true>("TS");
    run<48, true>("TS"); run<32, true>("TS");
    return 0;
}
