// Microbenchmark of the epilogue drain arithmetic: 32 elements per thread, W warps per SM.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t d; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a)); return d; }
template <int MODE>
__global__ void k(const float* in, uint32_t* out, int iters, unsigned long long* cyc) {
    float r[32]; uint32_t q[16];
    const float hb = in[threadIdx.x & 31];
    for (int i = 0; i < 32; ++i) r[i] = in[32 + i] * (threadIdx.x + 1);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float b = (MODE & 1) ? __shfl_sync(0xffffffffu, hb, i) : hb;
            v[i] = fmaf(r[i], 0.5f, b);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (MODE & 2) ? fmaf(v[i], tanh_approx(v[i]), v[i]) : fmaf(v[i], v[i], v[i]);
#pragma unroll
        for (int c = 0; c < 16; ++c) q[c] = pack(v[2 * c], v[2 * c + 1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] += __uint_as_float(q[i >> 1]) * 1e-9f;
    }
    long long t1 = clock64();
    uint32_t s = 0; for (int c = 0; c < 16; ++c) s ^= q[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name) {
    float* in; uint32_t* o; unsigned long long* c;
    cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
    for (int warps : {2, 4, 8, 16}) {
        const int iters = 200;
        k<MODE><<<148, warps * 32>>>(in, o, iters, c); cudaDeviceSynchronize();
        unsigned long long h[148]; cudaMemcpy(h, c, 148 * 8, cudaMemcpyDeviceToHost);
        double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
        printf("%-22s warps/SM=%2d: %.0f cycles per drain (32 el/thread)\n", name, warps, cy / iters);
    }
}
int main() { run<3>("shfl+tanh"); run<2>("tanh, bias in reg"); run<1>("shfl, no mufu"); run<0>("ffma only"); return 0; }
