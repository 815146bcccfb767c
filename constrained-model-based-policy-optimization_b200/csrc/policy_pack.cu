// Merged policy ensemble for the tcgen05 path: the Gaussian actor's mean MLP (ac_network.py:99-123) and
// the V / VC value ensembles (cpo_policy.py:452-468) are evaluated on the SAME observations every
// rollout step (cpo_policy.py:801-823).  All of them are 2-hidden-layer nets of equal width, so they are
// packed as ONE (1 + Ev + Evc)-member ensemble that shares a single input panel: one kernel launch
// instead of three, and one pass over the observations.
//
// Each member keeps its own input scaler by folding it into its first layer.  The shared panel holds
//     x' = (x - c) / s               (c, s) = the V ensemble's scaler (or identity)
// and member n, whose own normalisation is z = (x - mu_n) / sigma_n  (pens/utils.py:156), uses
//     W0'[k,:] = W0[k,:] * s_k / sigma_n,k        b0' = b0 + sum_k ((c_k - mu_n,k) / sigma_n,k) W0[k,:]
// which is algebraically identical (x = x' s + c); only the 16-bit rounding point moves.  The common
// transform keeps |x'| = O(1), so converting x' to fp16/bf16 loses nothing a per-net conversion keeps.
// The fp32 CUDA-core path does not use the merged net.
#include "common.cuh"

namespace {

// one block per (member, output neuron block); folds the scalers of member `e`
__global__ void fold_first_layer_kernel(const float* W0, const float* b0, int K, int HD, const float* mu_n,
                                        const float* sig_n, const float* c, const float* s, float* W0o,
                                        float* b0o) {
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < HD; h += gridDim.x * blockDim.x) {
        double acc = b0[h];
        for (int k = 0; k < K; ++k) {
            const float w = W0[(size_t)k * HD + h];
            const float sk = s ? s[k] : 1.0f, ck = c ? c[k] : 0.0f;
            const float sn = sig_n ? sig_n[k] : 1.0f, mn = mu_n ? mu_n[k] : 0.0f;
            W0o[(size_t)k * HD + h] = w * (sk / sn);
            acc += (double)((ck - mn) / sn) * (double)w;
        }
        b0o[h] = (float)acc;
    }
}

// W2 [HD, nout_src] -> [HD, Nout] (zero padded), b2 likewise
__global__ void pad_last_layer_kernel(const float* W2, const float* b2, int HD, int nsrc, int Nout, float* W2o,
                                      float* b2o) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HD * Nout; i += gridDim.x * blockDim.x) {
        const int h = i / Nout, c = i - h * Nout;
        W2o[i] = c < nsrc ? W2[(size_t)h * nsrc + c] : 0.f;
    }
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < Nout; c += blockDim.x) b2o[c] = c < nsrc ? b2[c] : 0.f;
}

bool mergeable(const Net& n, int hd, int O) {
    return n.loaded && n.n_layers == 3 && n.dims[0] == O && n.dims[1] == hd && n.dims[2] == hd &&
           n.acts[0] == n.acts[1] && n.acts[2] == CMBPO_ACT_NONE &&
           (n.acts[0] == CMBPO_ACT_SWISH || n.acts[0] == CMBPO_ACT_TANH) && !n.probabilistic;
}

}  // namespace

// (re)build ctx->polnet when the actor, V and VC are all loaded and compatible; otherwise leave it
// unloaded (the three nets then run as separate launches)
int policy_pack_build(cmbpo_ctx* ctx) {
    net_free(ctx->polnet);
    const Net& act = ctx->nets[CMBPO_NET_ACTOR];
    const Net& v = ctx->nets[CMBPO_NET_V];
    const Net& vc = ctx->nets[CMBPO_NET_VC];
    if (!act.loaded || !v.loaded || !vc.loaded) return 0;
    const int hd = act.dims[1], O = act.dims[0], A = act.dims[3];
    if (hd != 128 && hd != 256) return 0;
    if (!mergeable(act, hd, O) || !mergeable(v, hd, O) || !mergeable(vc, hd, O)) return 0;
    if (act.E != 1 || v.dims[3] != 1 || vc.dims[3] != 1 || O > 64) return 0;
    const int E = 1 + v.E + vc.E;
    if (E > CMBPO_MAX_E) return 0;
    Net& pn = ctx->polnet;
    pn.E = E; pn.n_layers = 3; pn.probabilistic = false;
    pn.dims[0] = O; pn.dims[1] = hd; pn.dims[2] = hd; pn.dims[3] = A;
    pn.D = A;
    pn.acts[0] = pn.acts[1] = CMBPO_ACT_SWISH; pn.acts[2] = CMBPO_ACT_NONE;   // per member below
    const size_t szW0 = (size_t)O * hd, szW1 = (size_t)hd * hd, szW2 = (size_t)hd * A;
    CUDA_TRY(cudaMalloc(&pn.W[0], E * szW0 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&pn.b[0], (size_t)E * hd * sizeof(float)));
    CUDA_TRY(cudaMalloc(&pn.W[1], E * szW1 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&pn.b[1], (size_t)E * hd * sizeof(float)));
    CUDA_TRY(cudaMalloc(&pn.W[2], E * szW2 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&pn.b[2], (size_t)E * A * sizeof(float)));
    // common input transform = V's scaler (every CMBPO config gives V and VC scalers, cpo_policy.py:462-463)
    const float* c = v.has_in ? v.mu_in : nullptr;
    const float* s = v.has_in ? v.sig_in : nullptr;
    if (v.has_in) {
        pn.has_in = true;
        CUDA_TRY(cudaMalloc(&pn.mu_in, O * sizeof(float)));
        CUDA_TRY(cudaMalloc(&pn.sig_in, O * sizeof(float)));
        CUDA_TRY(cudaMemcpyAsync(pn.mu_in, v.mu_in, O * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(pn.sig_in, v.sig_in, O * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    int m = 0;
    auto add = [&](const Net& n, int e_src) -> int {
        const int nsrc = n.dims[3];
        fold_first_layer_kernel<<<cdiv(hd, 128), 128, 0, ctx->stream>>>(
            n.W[0] + e_src * szW0, n.b[0] + (size_t)e_src * hd, O, hd, n.has_in ? n.mu_in : nullptr,
            n.has_in ? n.sig_in : nullptr, c, s, pn.W[0] + m * szW0, pn.b[0] + (size_t)m * hd);
        CUDA_TRY(cudaMemcpyAsync(pn.W[1] + m * szW1, n.W[1] + e_src * szW1, szW1 * sizeof(float),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(pn.b[1] + (size_t)m * hd, n.b[1] + (size_t)e_src * hd, hd * sizeof(float),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        pad_last_layer_kernel<<<cdiv(hd * A, 256), 256, 0, ctx->stream>>>(
            n.W[2] + (size_t)e_src * hd * nsrc, n.b[2] + (size_t)e_src * nsrc, hd, nsrc, A, pn.W[2] + m * szW2,
            pn.b[2] + (size_t)m * A);
        pn.member_act[m] = n.acts[0];
        ++m;
        return 0;
    };
    if (add(act, 0)) return 1;
    for (int e = 0; e < v.E; ++e) if (add(v, e)) return 1;
    for (int e = 0; e < vc.E; ++e) if (add(vc, e)) return 1;
    ctx->pol_nv = v.E; ctx->pol_nvc = vc.E;
    ctx->launches += 2 * E;
    pn.loaded = true;
    CUDA_TRY(cudaGetLastError());
    if (!ens_tc_supported(pn)) { net_free(pn); return 0; }
    return ens_tc_prepare(ctx, pn);
}
