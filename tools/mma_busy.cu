// Microbenchmark: does the tcgen05.mma issue/execution rate of one warp (TS form, M=128, N=64,
// K=16, f16) depend on (a) 16 other warps of the SM running drain-like arithmetic (SHFL, FFMA,
// MUFU.TANH, packed convert) on the same four sub-cores, (b) the operand data (zeros vs random)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_busy mma_busy.cu
#include <cstdio>
#include "../constrained-model-based-policy-optimization_b200/csrc/tc_common.cuh"
using namespace tc;
void cmbpo_set_error(const char*, ...) {}

__global__ void __launch_bounds__(640, 1) busy_kernel(int n_mma, int busy_iters, int random_data,
                                                       unsigned long long* out, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 20 B tiles of 8 KB (the main ring of the real kernel)
    for (int i = threadIdx.x; i < 20 * 8192 / 4; i += 640) {
        uint32_t v = 0;
        if (random_data) {
            uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            __half2 x = __floats2half2_rn(((h & 0xffff) / 32768.f) - 1.f, ((h >> 16) / 32768.f) - 1.f);
            v = *reinterpret_cast<uint32_t*>(&x);
        }
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 2) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp >= 4 && random_data) {           // A operand (H1 area, columns 0..255) random as well
        uint32_t q[16];
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int c = (warp - 4) / 4 * 64; c < (warp - 4) / 4 * 64 + 64; c += 16) {
            for (int i = 0; i < 16; ++i) {
                uint32_t h = (threadIdx.x * 977u + c * 131u + i) * 2654435761u; h ^= h >> 15;
                __half2 x = __floats2half2_rn(((h & 0xffff) / 32768.f) - 1.f, ((h >> 16) / 32768.f) - 1.f);
                q[i] = *reinterpret_cast<uint32_t*>(&x);
            }
            tmem_st16(c + lane_base, q);
        }
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1) {
        const uint32_t idesc = idesc_f16(0, 64);
        const uint64_t dB = smem_desc_sw128(smem_u32(smem));
        long long t0 = clock64();
        if (elect_one()) {
            uint32_t tile = 0;
            for (int i = 0; i < n_mma / 4; ++i) {
                const uint64_t d = dB + (uint64_t)((tile * 8192) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_f16_ts(256u + (i & 1) * 64, (uint32_t)(((i & 7) * 32) + ks * 8), d + 2 * ks, idesc, (i & 7) != 0 || ks != 0);
                if (++tile == 20) tile = 0;
            }
            mma_commit(&bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    } else if (warp >= 4) {
        float acc = 0.f;
        const float hb = 0.001f * lane;
        for (int it = 0; it < busy_iters; ++it) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(acc + (float)i, 0.5f, __shfl_sync(0xffffffffu, hb, i));
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = swish_half(v[i]);
#pragma unroll
            for (int c = 0; c < 16; ++c) acc += __uint_as_float(Cvt<0>::pack(v[2 * c], v[2 * c + 1])) * 1e-30f;
        }
        if (acc == 123.456f) sink[threadIdx.x] = acc;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 2) tmem_dealloc(0, 512);
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 16);
    float* sink; cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(busy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 170000);
    const int n_mma = 4000;
    for (int random_data = 0; random_data < 2; ++random_data)
        for (int busy : {0, 200}) {
            for (int rep = 0; rep < 2; ++rep) {
                busy_kernel<<<148, 640, 170000>>>(n_mma, busy, random_data, d, sink);
                cudaError_t e = cudaDeviceSynchronize();
                unsigned long long h[296];
                cudaMemcpy(h, d, 148 * 16, cudaMemcpyDeviceToHost);
                double issue = 0, total = 0;
                for (int b = 0; b < 148; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; }
                if (rep == 1)
                    printf("data=%s busy_warps=%s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n",
                           random_data ? "random" : "zero", busy ? "16" : "0", issue / 148 / n_mma, total / 148 / n_mma,
                           cudaGetErrorString(e));
            }
        }
    return 0;
}
