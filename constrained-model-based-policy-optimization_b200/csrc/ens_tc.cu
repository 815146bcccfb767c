// K1 on the tensor cores: the 3-layer ensemble MLP chain  X -> act(X W0+b0) -> act(. W1+b1) -> . W2+b2
// for one 128-row tile per CTA and all E members, with tcgen05.mma (kind::f16, fp32 accumulators
// in TMEM), weights streamed from L2 by the bulk-copy engine as pre-swizzled shared-memory tile
// images, and the hidden activations never leaving the SM:
//
//   layer 0   D[128 x 128-col chunk] = XA  x W0 chunk          (K = in_dim padded to 16)
//             epilogue: bias + act -> 16-bit -> H1 panels in shared memory (A operand of layer 1)
//   layer 1   D chunk = sum over 64-wide K panels  H1[kp] x W1[kp, chunk]
//             epilogue: bias + act -> 16-bit -> one of two H2 half-chunk panels
//   layer 2   OUT[128 x NP] += H2 half x W2[half rows, :]        (accumulated across chunks in TMEM)
//             epilogue: + b2 -> raw outputs (fp32) to global
//
// Warp roles (384 threads): warp 0 = weight producer (one lane), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4-11 = two epilogue warpgroups (each owns 64 of a chunk's 128 columns;
// thread <-> row, TMEM lane = row).
//
// Replaces models/pens/fc.py:74-95 x3 + the input scaler of models/pens/utils.py:156 (fused into
// the XA load).  The output scaler / exp are applied by the consumer (ens_head_kernel or the
// rollout row math) exactly as on the fp32 path.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int PANEL = 16384;      // [128 rows x 64 el x 2 B]
constexpr int NTHREADS = 384;

struct Stage {        // one weight tile image in the packed stream
    int layer, n0, k0, rows;
    unsigned long long off;
};

struct TcParams {
    const uint8_t* wpack; unsigned long long member_bytes;
    const float* bias; int bias_stride;
    int E, K0, KS0, Nout, NP;
    const float* x; long long N; int ldx;
    const float *mu_in, *sig_in;
    float* out; long long out_member_stride;   // out[e*stride + row*Nout + c]
    int ntiles;
};

__host__ __device__ constexpr int n_stages_of(int hd) { return hd == 512 ? 3 : 4; }

template <int HD>
struct Smem {
    static constexpr int NC = HD / 128, KP = HD / 64, NS = n_stages_of(HD);
    static constexpr int XA = 0;
    static constexpr int H1 = XA + PANEL;
    static constexpr int H2 = H1 + KP * PANEL;
    static constexpr int WR = H2 + 2 * PANEL;
    static constexpr int BAR = WR + NS * PANEL;
    static constexpr int TOTAL = BAR + 512;
    // barrier indices
    static constexpr int W_FULL = 0, W_EMPTY = NS, D_FULL = 2 * NS, D_EMPTY = 2 * NS + 2,
                         H1_FULL = 2 * NS + 4, H2_FULL = 2 * NS + 5, H2_EMPTY = 2 * NS + 7,
                         OUT_FULL = 2 * NS + 9, OUT_EMPTY = 2 * NS + 10, X_FULL = 2 * NS + 11,
                         NBAR = 2 * NS + 12;
};

template <int ACT> __device__ __forceinline__ float activate(float x) {
    if (ACT == CMBPO_ACT_SWISH) return swish_fast(x);
    if (ACT == CMBPO_ACT_TANH) return tanh_approx(x);
    return x;
}

// 64 accumulator columns of this thread's row -> bias + act -> 16-bit -> 8 swizzled 16-B chunks
template <int FMT, int ACT>
__device__ __forceinline__ void drain_half(uint32_t taddr, const float* __restrict__ bias, uint8_t* panel,
                                           int row) {
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        uint32_t r[32];
        tmem_ld32(taddr + part * 32, r);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(bias + part * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {       // 4 chunks of 8 columns
            float4 ba = __ldg(b4 + 2 * c), bb = __ldg(b4 + 2 * c + 1);
            float v0 = activate<ACT>(__uint_as_float(r[8 * c + 0]) + ba.x);
            float v1 = activate<ACT>(__uint_as_float(r[8 * c + 1]) + ba.y);
            float v2 = activate<ACT>(__uint_as_float(r[8 * c + 2]) + ba.z);
            float v3 = activate<ACT>(__uint_as_float(r[8 * c + 3]) + ba.w);
            float v4 = activate<ACT>(__uint_as_float(r[8 * c + 4]) + bb.x);
            float v5 = activate<ACT>(__uint_as_float(r[8 * c + 5]) + bb.y);
            float v6 = activate<ACT>(__uint_as_float(r[8 * c + 6]) + bb.z);
            float v7 = activate<ACT>(__uint_as_float(r[8 * c + 7]) + bb.w);
            uint4 q;
            q.x = Cvt<FMT>::pack(v0, v1); q.y = Cvt<FMT>::pack(v2, v3);
            q.z = Cvt<FMT>::pack(v4, v5); q.w = Cvt<FMT>::pack(v6, v7);
            *reinterpret_cast<uint4*>(panel + panel_off(row, part * 4 + c)) = q;
        }
    }
}

template <int HD, int FMT, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1) ens_mlp3_tc_kernel(const TcParams p) {
    using S = Smem<HD>;
    constexpr int NC = S::NC, KP = S::KP, NS = S::NS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sXA = smem + S::XA;
    uint8_t* sH1 = smem + S::H1;
    uint8_t* sH2 = smem + S::H2;
    uint8_t* sW = smem + S::WR;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + S::NBAR * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar + S::W_FULL + i, 1); mbar_init(bar + S::W_EMPTY + i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar + S::D_FULL + i, 1); mbar_init(bar + S::D_EMPTY + i, 8);
            mbar_init(bar + S::H2_FULL + i, 4); mbar_init(bar + S::H2_EMPTY + i, 1);
        }
        mbar_init(bar + S::H1_FULL, 8 * NC);
        mbar_init(bar + S::OUT_FULL, 1); mbar_init(bar + S::OUT_EMPTY, 4);
        mbar_init(bar + S::X_FULL, 4);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t w2_bytes = (uint32_t)p.NP * 128u;

    if (warp == 0) {
        // ===== weight producer: streams the per-member stage program in consumption order =====
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                for (int e = 0; e < p.E; ++e) {
                    const uint8_t* src = p.wpack + (unsigned long long)e * p.member_bytes;
                    auto push = [&](uint32_t bytes) {
                        mbar_wait(bar + S::W_EMPTY + s, ph ^ 1);
                        mbar_expect_tx(bar + S::W_FULL + s, bytes);
                        bulk_g2s(sW + s * PANEL, src, bytes, bar + S::W_FULL + s);
                        src += bytes;
                        if (++s == NS) { s = 0; ph ^= 1; }
                    };
                    for (int j = 0; j < NC; ++j) push(PANEL);
                    for (int j = 0; j < NC; ++j) {
                        for (int kp = 0; kp < KP; ++kp) push(PANEL);
                        if (j >= 1) { push(w2_bytes); push(w2_bytes); }
                    }
                    push(w2_bytes); push(w2_bytes);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            const uint32_t idesc_h = idesc_f16(FMT, 128);
            const uint32_t idesc_o = idesc_f16(FMT, p.NP);
            const uint64_t dXA = smem_desc_sw128(smem_u32(sXA));
            uint32_t s = 0, ph = 0, g = 0, m = 0, c1 = 0, it = 0;
            auto next_stage = [&]() { if (++s == NS) { s = 0; ph ^= 1; } };
            auto l2_partials = [&](int jj) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (jj == 0 && h == 0) { mbar_wait(bar + S::OUT_EMPTY, (m & 1) ^ 1); }
                    mbar_wait(bar + S::H2_FULL + h, c1 & 1);
                    mbar_wait(bar + S::W_FULL + s, ph);
                    tc_fence_after();
                    const uint64_t dA = smem_desc_sw128(smem_u32(sH2 + h * PANEL));
                    const uint64_t dB = smem_desc_sw128(smem_u32(sW + s * PANEL));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma_f16(tmem + 256, dA + 2 * ks, dB + 2 * ks, idesc_o, !(jj == 0 && h == 0 && ks == 0));
                    mma_commit(bar + S::W_EMPTY + s);
                    next_stage();
                    mma_commit(bar + S::H2_EMPTY + h);
                }
                ++c1;
            };
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                mbar_wait(bar + S::X_FULL, it & 1);
                tc_fence_after();
                for (int e = 0; e < p.E; ++e) {
                    for (int j = 0; j < NC; ++j) {              // layer 0
                        const uint32_t buf = g & 1, n = g >> 1;
                        mbar_wait(bar + S::D_EMPTY + buf, (n & 1) ^ 1);
                        mbar_wait(bar + S::W_FULL + s, ph);
                        tc_fence_after();
                        const uint64_t dB = smem_desc_sw128(smem_u32(sW + s * PANEL));
                        for (int ks = 0; ks < p.KS0; ++ks)
                            mma_f16(tmem + buf * 128, dXA + 2 * ks, dB + 2 * ks, idesc_h, ks > 0);
                        mma_commit(bar + S::W_EMPTY + s);
                        next_stage();
                        mma_commit(bar + S::D_FULL + buf);
                        ++g;
                    }
                    mbar_wait(bar + S::H1_FULL, m & 1);
                    tc_fence_after();
                    for (int j = 0; j < NC; ++j) {              // layer 1 (+ layer 2 of the previous chunk)
                        const uint32_t buf = g & 1, n = g >> 1;
                        mbar_wait(bar + S::D_EMPTY + buf, (n & 1) ^ 1);
                        tc_fence_after();
                        for (int kp = 0; kp < KP; ++kp) {
                            mbar_wait(bar + S::W_FULL + s, ph);
                            tc_fence_after();
                            const uint64_t dA = smem_desc_sw128(smem_u32(sH1 + kp * PANEL));
                            const uint64_t dB = smem_desc_sw128(smem_u32(sW + s * PANEL));
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_f16(tmem + buf * 128, dA + 2 * ks, dB + 2 * ks, idesc_h, (kp | ks) > 0);
                            mma_commit(bar + S::W_EMPTY + s);
                            next_stage();
                        }
                        mma_commit(bar + S::D_FULL + buf);
                        ++g;
                        if (j >= 1) l2_partials(j - 1);
                    }
                    l2_partials(NC - 1);
                    mma_commit(bar + S::OUT_FULL);
                    ++m;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 2 warpgroups x 128 threads, thread <-> row =====
        const int wg = (warp - 4) >> 2;
        const int wq = warp & 3;                   // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        uint32_t g = 0, m = 0, c1 = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const long long grow = (long long)tile * 128 + row;
            if (wg == 0) {
                // XA: this row of the input, scaled (pens/utils.py:156), 16-bit, zero padded to 64
                const float* xr = p.x + grow * p.ldx;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = c * 8 + i;
                        float t = 0.f;
                        if (k < p.K0 && grow < p.N) {
                            t = xr[k];
                            if (p.mu_in) t = __fdiv_rn(__fsub_rn(t, p.mu_in[k]), p.sig_in[k]);
                        }
                        v[i] = t;
                    }
                    uint4 q;
                    q.x = Cvt<FMT>::pack(v[0], v[1]); q.y = Cvt<FMT>::pack(v[2], v[3]);
                    q.z = Cvt<FMT>::pack(v[4], v[5]); q.w = Cvt<FMT>::pack(v[6], v[7]);
                    *reinterpret_cast<uint4*>(sXA + panel_off(row, c)) = q;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + S::X_FULL);
            }
            for (int e = 0; e < p.E; ++e) {
                const float* bias = p.bias + (long long)e * p.bias_stride;
                for (int j = 0; j < NC; ++j) {                  // layer-0 chunk -> H1 panel 2j+wg
                    const uint32_t buf = g & 1, n = g >> 1;
                    mbar_wait(bar + S::D_FULL + buf, n & 1);
                    tc_fence_after();
                    drain_half<FMT, ACT>(tmem + buf * 128 + wg * 64 + lane_base, bias + j * 128 + wg * 64,
                                         sH1 + (2 * j + wg) * PANEL, row);
                    tc_fence_before();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(bar + S::D_EMPTY + buf); mbar_arrive(bar + S::H1_FULL); }
                    ++g;
                }
                for (int j = 0; j < NC; ++j) {                  // layer-1 chunk -> H2 half wg
                    const uint32_t buf = g & 1, n = g >> 1;
                    mbar_wait(bar + S::D_FULL + buf, n & 1);
                    mbar_wait(bar + S::H2_EMPTY + wg, (c1 & 1) ^ 1);
                    tc_fence_after();
                    drain_half<FMT, ACT>(tmem + buf * 128 + wg * 64 + lane_base,
                                         bias + HD + j * 128 + wg * 64, sH2 + wg * PANEL, row);
                    tc_fence_before();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(bar + S::D_EMPTY + buf); mbar_arrive(bar + S::H2_FULL + wg); }
                    ++g; ++c1;
                }
                if (wg == 0) {                                   // OUT -> global raw outputs
                    mbar_wait(bar + S::OUT_FULL, m & 1);
                    tc_fence_after();
                    float* orow = p.out + (long long)e * p.out_member_stride + grow * p.Nout;
                    const float* b2 = bias + 2 * HD;
                    for (int c0 = 0; c0 < p.NP; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(tmem + 256 + c0 + lane_base, r);
                        tmem_ld_wait();
                        if (grow < p.N) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.Nout) orow[c0 + i] = __uint_as_float(r[i]) + b2[c0 + i];
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar + S::OUT_EMPTY);
                }
                ++m;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---- weight packing -----------------------------------------------------------------------------
// W_l fp32 [E, K, M] (k-major rows, fc.py layout) -> stream of swizzled B tiles: tile row n = output
// neuron n0+n, 64 consecutive k from k0; zero padded.
template <int FMT>
__global__ void pack_weights_kernel(const float* W0, const float* W1, const float* W2, int K0, int HD,
                                    int Nout, const Stage* stages, int n_stages,
                                    unsigned long long member_bytes, uint8_t* out) {
    const int e = blockIdx.y;
    const Stage st = stages[blockIdx.x];
    const float* W; int K, M;
    if (st.layer == 0) { W = W0 + (size_t)e * K0 * HD; K = K0; M = HD; }
    else if (st.layer == 1) { W = W1 + (size_t)e * HD * HD; K = HD; M = HD; }
    else { W = W2 + (size_t)e * HD * Nout; K = HD; M = Nout; }
    uint8_t* dst = out + (size_t)e * member_bytes + st.off;
    for (int idx = threadIdx.x; idx < st.rows * 64; idx += blockDim.x) {
        const int n = idx >> 6, k = idx & 63;
        const int gk = st.k0 + k, gn = st.n0 + n;
        float v = (gk < K && gn < M) ? W[(size_t)gk * M + gn] : 0.f;
        uint16_t h;
        if (FMT == 0) { __half x = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); h = *reinterpret_cast<uint16_t*>(&x); }
        else { __nv_bfloat16 x = __float2bfloat16_rn(v); h = *reinterpret_cast<uint16_t*>(&x); }
        *reinterpret_cast<uint16_t*>(dst + panel_off(n, k >> 3) + (k & 7) * 2) = h;
    }
}

__global__ void pack_bias_kernel(const float* b0, const float* b1, const float* b2, int HD, int Nout, int NP,
                                 float* out) {
    const int e = blockIdx.x;
    const int stride = 2 * HD + NP;
    for (int i = threadIdx.x; i < stride; i += blockDim.x) {
        float v;
        if (i < HD) v = b0[(size_t)e * HD + i];
        else if (i < 2 * HD) v = b1[(size_t)e * HD + i - HD];
        else v = (i - 2 * HD < Nout) ? b2[(size_t)e * Nout + i - 2 * HD] : 0.f;
        out[(size_t)e * stride + i] = v;
    }
}

std::vector<Stage> stage_program(int HD, int NP, unsigned long long* total) {
    const int NC = HD / 128, KP = HD / 64;
    std::vector<Stage> v;
    unsigned long long off = 0;
    auto add = [&](int layer, int n0, int k0, int rows) {
        v.push_back(Stage{layer, n0, k0, rows, off});
        off += (unsigned long long)rows * 128;
    };
    for (int j = 0; j < NC; ++j) add(0, j * 128, 0, 128);
    for (int j = 0; j < NC; ++j) {
        for (int kp = 0; kp < KP; ++kp) add(1, j * 128, kp * 64, 128);
        if (j >= 1) { add(2, 0, (j - 1) * 128, NP); add(2, 0, (j - 1) * 128 + 64, NP); }
    }
    add(2, 0, (NC - 1) * 128, NP); add(2, 0, (NC - 1) * 128 + 64, NP);
    *total = off;
    return v;
}

template <int HD, int FMT, int ACT>
int launch_tc(cmbpo_ctx* ctx, const TcParams& p) {
    using S = Smem<HD>;
    const int smem = S::TOTAL + 1024;
    auto kern = ens_mlp3_tc_kernel<HD, FMT, ACT>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = p.ntiles < ctx->sm_count ? p.ntiles : ctx->sm_count;
    kern<<<grid, NTHREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <int HD, int FMT>
int launch_tc_act(cmbpo_ctx* ctx, const TcParams& p, int act) {
    if (act == CMBPO_ACT_SWISH) return launch_tc<HD, FMT, CMBPO_ACT_SWISH>(ctx, p);
    return launch_tc<HD, FMT, CMBPO_ACT_TANH>(ctx, p);
}

template <int FMT>
int launch_tc_hd(cmbpo_ctx* ctx, const TcParams& p, int hd, int act) {
    if (hd == 128) return launch_tc_act<128, FMT>(ctx, p, act);
    if (hd == 256) return launch_tc_act<256, FMT>(ctx, p, act);
    return launch_tc_act<512, FMT>(ctx, p, act);
}

}  // namespace

bool ens_tc_supported(const Net& n) {
    if (!n.loaded || n.n_layers != 3) return false;
    const int hd = n.dims[1];
    if (n.dims[2] != hd || (hd != 128 && hd != 256 && hd != 512)) return false;
    if (n.dims[0] > 64 || n.dims[3] > 128) return false;
    if (n.acts[0] != n.acts[1] || n.acts[2] != CMBPO_ACT_NONE) return false;
    return n.acts[0] == CMBPO_ACT_SWISH || n.acts[0] == CMBPO_ACT_TANH;
}

// pack both 16-bit formats once per weight upload
int ens_tc_prepare(cmbpo_ctx* ctx, Net& net) {
    const int HD = net.dims[1], K0 = net.dims[0], Nout = net.dims[3];
    const int NP = ((Nout + 15) / 16) * 16;
    unsigned long long member_bytes = 0;
    std::vector<Stage> prog = stage_program(HD, NP, &member_bytes);
    Stage* d_prog;
    CUDA_TRY(cudaMalloc(&d_prog, prog.size() * sizeof(Stage)));
    CUDA_TRY(cudaMemcpyAsync(d_prog, prog.data(), prog.size() * sizeof(Stage), cudaMemcpyHostToDevice, ctx->stream));
    for (int prec = CMBPO_PREC_BF16; prec <= CMBPO_PREC_FP16; ++prec) {
        CUDA_TRY(cudaMalloc(&net.tc_pack[prec], (size_t)net.E * member_bytes));
        net.tc_pack_bytes[prec] = (size_t)member_bytes;
        dim3 grid((unsigned)prog.size(), net.E);
        if (prec == CMBPO_PREC_FP16)
            pack_weights_kernel<0><<<grid, 256, 0, ctx->stream>>>(net.W[0], net.W[1], net.W[2], K0, HD, Nout, d_prog,
                                                               (int)prog.size(), member_bytes, (uint8_t*)net.tc_pack[prec]);
        else
            pack_weights_kernel<1><<<grid, 256, 0, ctx->stream>>>(net.W[0], net.W[1], net.W[2], K0, HD, Nout, d_prog,
                                                               (int)prog.size(), member_bytes, (uint8_t*)net.tc_pack[prec]);
    }
    CUDA_TRY(cudaMalloc(&net.tc_bias, (size_t)net.E * (2 * HD + NP) * sizeof(float)));
    pack_bias_kernel<<<net.E, 256, 0, ctx->stream>>>(net.b[0], net.b[1], net.b[2], HD, Nout, NP, net.tc_bias);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaFree(d_prog));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ens_forward_tc(cmbpo_ctx* ctx, Net& net, const float* x, int64_t N, float* out_raw, int precision) {
    CMBPO_CHECK(precision == CMBPO_PREC_BF16 || precision == CMBPO_PREC_FP16, "bad precision %d", precision);
    CMBPO_CHECK(net.tc_pack[precision], "tcgen05 weights not packed");
    if (N <= 0) return 0;
    const int HD = net.dims[1];
    TcParams p;
    p.wpack = (const uint8_t*)net.tc_pack[precision];
    p.member_bytes = net.tc_pack_bytes[precision];
    p.Nout = net.dims[3];
    p.NP = ((p.Nout + 15) / 16) * 16;
    p.bias = net.tc_bias; p.bias_stride = 2 * HD + p.NP;
    p.E = net.E; p.K0 = net.dims[0]; p.KS0 = (p.K0 + 15) / 16;
    p.x = x; p.N = N; p.ldx = net.dims[0];
    p.mu_in = net.has_in ? net.mu_in : nullptr; p.sig_in = net.sig_in;
    p.out = out_raw; p.out_member_stride = (long long)N * p.Nout;
    p.ntiles = (int)((N + 127) / 128);
    if (precision == CMBPO_PREC_FP16) return launch_tc_hd<0>(ctx, p, HD, net.acts[0]);
    return launch_tc_hd<1>(ctx, p, HD, net.acts[0]);
}
