// Microbenchmark: SM-wide throughput of the MUFU ops the activation epilogue can use.
#include <cstdio>
#include <cuda_fp16.h>
template <int OP>
__global__ void k(float* out, int iters, unsigned long long* cyc) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) { unsigned u = __float_as_uint(a[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
            if (OP == 4) a[i] = fmaf(a[i], 1.0001f, 0.5f);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name) {
    float* o; unsigned long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
    for (int threads : {128, 512, 1024}) {
        const int iters = 2000;
        k<OP><<<148, threads>>>(o, iters, c); cudaDeviceSynchronize();
        unsigned long long h[148]; cudaMemcpy(h, c, 148 * 8, cudaMemcpyDeviceToHost);
        double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
        printf("%-14s threads/SM=%4d: %.2f lane-ops/cycle/SM\n", name, threads, (double)threads * iters * 8 / cy);
    }
}
int main() { run<0>("tanh.f32"); run<1>("ex2.f32"); run<2>("rcp.f32"); run<3>("tanh.f16x2"); run<4>("ffma"); return 0; }
