// Microbenchmark of the activation epilogue INCLUDING its tensor-memory traffic: W epilogue warps
// (warp w -> TMEM lane quarter w % 4) loop over   tcgen05.ld 32 columns -> +bias, swish -> 16-bit ->
// tcgen05.st 16 columns.   Answers: what bounds a round of drains -- MUFU, the TMEM read port, SHFL?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/epi_rate tools/epi_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t d; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a)); return d; }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// MODE bits: 1 = tcgen05.ld, 2 = tcgen05.st, 4 = bias by SHFL, 8 = bias by LDS.128 broadcast, 16 = tanh (else FFMA stand-in)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* in, uint32_t* out, int iters, unsigned long long* cyc) {
    __shared__ uint32_t slot;
    __shared__ __align__(16) float s_bias[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) s_bias[threadIdx.x] = in[threadIdx.x];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int wq = warp >> 2;                       // 0..3: which 128 columns this warp cycles through
    uint32_t r[32]; uint32_t q[16];
    const float hb = in[lane];
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(in[32 + i] * (threadIdx.x + 1));
    for (int c = 0; c < 16; ++c) q[c] = 0;
    // initialise this warp's columns so that loads return defined data
    for (int c = 0; c < 128; c += 16) tmem_st16(tmem + lane_base + wq * 128 + c, q);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = wq * 128 + (it & 1) * 64;
        if (MODE & 1) {
            tmem_ld32(tmem + lane_base + col, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        float v[32];
        if (MODE & 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]), 0.5f, __shfl_sync(0xffffffffu, hb, i));
        } else if (MODE & 8) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(s_bias + (it & 1) * 32 + i);
                v[i] = fmaf(__uint_as_float(r[i]), 0.5f, b.x); v[i + 1] = fmaf(__uint_as_float(r[i + 1]), 0.5f, b.y);
                v[i + 2] = fmaf(__uint_as_float(r[i + 2]), 0.5f, b.z); v[i + 3] = fmaf(__uint_as_float(r[i + 3]), 0.5f, b.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (MODE & 16) ? fmaf(v[i], tanh_approx(v[i]), v[i]) : fmaf(v[i], v[i], v[i]);
#pragma unroll
        for (int c = 0; c < 16; ++c) q[c] = pack(v[2 * c], v[2 * c + 1]);
        if (MODE & 2) {
            tmem_st16(tmem + lane_base + col + 32, q);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (!(MODE & 1)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(q[i >> 1]) * 1e-9f);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0; for (int c = 0; c < 16; ++c) s ^= q[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int MODE> void run(const char* name) {
    float* in; uint32_t* o; unsigned long long* c;
    cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
    for (int warps : {4, 8, 16}) {
        const int iters = 200;
        k<MODE><<<148, warps * 32>>>(in, o, iters, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        unsigned long long h[148]; cudaMemcpy(h, c, 148 * 8, cudaMemcpyDeviceToHost);
        double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
        printf("%-34s warps/SM=%2d: %6.0f cycles per round (32 el/thread) = %5.2f el/clk/SM\n", name, warps, cy / iters,
               warps * 32.0 * 32.0 / (cy / iters));
    }
    cudaFree(in); cudaFree(o); cudaFree(c);
}
int main() {
    run<1>("ld only");
    run<1 | 2>("ld + st");
    run<16>("tanh only (no bias, no tmem)");
    run<4 | 16>("shfl bias + tanh");
    run<8 | 16>("lds bias + tanh");
    run<1 | 2 | 16>("ld + tanh + st");
    run<1 | 2 | 4 | 16>("ld + shfl bias + tanh + st");
    run<1 | 2 | 8 | 16>("ld + lds bias + tanh + st");
    run<1 | 2 | 8>("ld + lds bias + ffma + st");
    return 0;
}
