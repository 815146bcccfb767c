// Per-row pieces of one rollout step shared by the step-wise kernels (rollout.cu) and the fused
// tcgen05 step kernel (ens_tc.cu), so that both apply literally the same code:
//   policy head   cpo_policy.py:801-835 (value heads through PE.predict, pe.py:343), ac_network.py:105-111
//   sampler rules model_sampler.py:275-367 + ModelBuffer.store_multiple (modelbuffer.py:114-135)
#pragma once
#include "common.cuh"
#include "row_math.cuh"

struct ValueHead {            // one non-probabilistic value ensemble read through PE.predict
    const float* raw;         // [E, N, ld], value in column 0
    int E, ld;
    const float *mu_out, *sig_out;   // [1] or null
};

__device__ __forceinline__ float value_of(const ValueHead& h, int64_t N, int64_t p) {
    // tf.reduce_mean over members of inverse_transform(out) (pe.py:343, pens/utils.py:167)
    float s = 0.f;
    for (int e = 0; e < h.E; ++e) {
        float m = h.raw[((int64_t)e * N + p) * h.ld];
        if (h.mu_out) m = __fadd_rn(__fmul_rn(h.sig_out[0], m), h.mu_out[0]);
        s = (e == 0) ? m : __fadd_rn(s, m);
    }
    return __fdiv_rn(s, (float)h.E);
}

// Gaussian head of one row: pi, mu rows and logp; no per-row arrays (everything streams through
// registers).  eps_row: injected standard normals of this PATH (null -> Philox keyed by gid, step).
template <bool FAST>
__device__ __forceinline__ float policy_head_row(const float* mu_row, const float* log_std, const float* eps_row,
                                                 uint64_t seed, int64_t gid, int step, int A, float* pi_row,
                                                 float* mu_out_row, float* pi_out_row2 = nullptr) {
    NpSumStream acc(A);
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
    for (int a = 0; a < A; ++a) {
        if (!eps_row && (a & 3) == 0) {            // one Philox block and two Box-Muller pairs per four values
            uint32_t o[4];
            philox4x32_10((uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), (uint32_t)step,
                          ((uint32_t)RNG_STREAM_ACT << 16) | (uint32_t)(a >> 2), (uint32_t)seed,
                          (uint32_t)(seed >> 32), o);
            box_muller(o[0], o[1], z0, z1);
            if (a + 2 < A) box_muller(o[2], o[3], z2, z3);
        }
        const int k = a & 3;
        const float eps = eps_row ? eps_row[a] : (k == 0 ? z0 : (k == 1 ? z1 : (k == 2 ? z2 : z3)));
        const float mu = mu_row[a];
        float pi;
        acc.add(a, actor_dim<FAST>(mu, log_std[a], eps, pi));
        if (pi_row) pi_row[a] = pi;
        if (mu_out_row) mu_out_row[a] = mu;
        if (pi_out_row2) pi_out_row2[a] = pi;
    }
    return acc.result();
}

// ---- what a step does with one fed row once FakeEnv.step's outputs are known ---------------------
struct StepRules {
    int64_t B; int t, last_storable;          // last_storable = T-2: storing it ends the path (horizon)
    int uncertainty; double dkl_lim;
    int no_store;                             // CMBPO_ROLLOUT_NO_STORE: per-path results and statistics only
    uint8_t* alive;                           // [B] by path
    uint8_t* pending;                         // [B] by path: bit0 last_val, bit1 last_cval wanted from the next policy pass
    cmbpo_rollout_bufs b;
};

struct RowCarry { float v, vc, logp; double dkl, ret, cost; };

// returns 1 = the step was stored, 2 = the path was cut as too uncertain BEFORE the store.
// st[0..3] += rows fed, sum dkl, rows stored, sum ep_var  (the caller reduces them per step)
__device__ __forceinline__ int step_row_commit(const StepRules& a, int64_t p, const EnvRowOut& o, const RowCarry& pf,
                                               double& st0, double& st1, double& st2, double& st3) {
    const float v = pf.v, vc = pf.vc;
    // uncertainty cut-off BEFORE the step is stored (model_sampler.py:275-290)
    const double next_dkl = pf.dkl + (double)o.dkl_path;
    const bool cut = a.uncertainty && next_dkl >= a.dkl_lim;
    st0 += 1.0; st1 += (double)o.dkl_path;
    if (cut) {
        a.alive[p] = 0;
        a.b.end_reason[p] = CMBPO_END_UNCERTAIN;
        a.b.last_val[p] = v; a.b.last_cval[p] = vc;      // V(s_t), VC(s_t): model_sampler.py:401-407
        return 2;
    }
    const int t = a.t;
    if (!a.no_store) {
        const int64_t row = (int64_t)t * a.B + p;      // ModelBuffer.store_multiple, time-major
        a.b.rew[row] = o.rew; a.b.val[row] = v; a.b.cost[row] = o.cost; a.b.cval[row] = vc;
        a.b.logp[row] = pf.logp; a.b.dyn_error[row] = o.ep_var_mean; a.b.dkl[row] = o.dkl_path;
        a.b.term[row] = o.term ? 1 : 0;
    }
    a.b.length[p] = t + 1;
    a.b.cum_dkl[p] = next_dkl;                        // model_sampler.py:332
    a.b.path_return[p] = pf.ret + (double)o.rew;      // :317-318
    a.b.path_cost[p] = pf.cost + (double)o.cost;
    st2 += 1.0; st3 += (double)o.ep_var_sum;
    if (t >= a.last_storable) {                       // path_length >= max_path_length-1 (:352)
        a.alive[p] = 0; a.b.end_reason[p] = CMBPO_END_HORIZON; a.pending[p] = 3;
    } else if (o.term) {                              // env terminal (:357-364)
        a.alive[p] = 0; a.b.end_reason[p] = CMBPO_END_TERMINAL;
        a.b.last_val[p] = 0.f; a.pending[p] = 2;
    }
    return 1;
}

// ---- everything the fused step kernel needs besides the GEMM operands ----------------------------
// Raw outputs of the dynamics ensemble never reach DRAM in the [E,N,2D] form: the GEMM epilogue writes
// them TILE-TRANSPOSED into an L2-resident scratch -- tile slot s, member e, column c, row r at
// ((s*E + e)*W + c)*128 + r -- so both the epilogue's stores (thread = row, loop over columns) and the
// row math's loads (lane = row) are fully coalesced.
struct FusedStep {
    StepRules rules;
    int O, A, n_elite;
    EnvRowCfg c;
    int64_t path_base; uint64_t seed;
    float* cur_obs;                 // [B,O] by row: in (s_t) / out (s_{t+1})
    // policy GEMM output of this step (merged actor + V + VC ensemble): [1 + nv + nvc, B, A]
    const float* pol_raw;
    ValueHead v, vc;
    const float* log_std;           // [A]
    float *pi, *mu, *logp, *vrow, *vcrow;     // per-row stash [B,A],[B,A],[B],[B],[B]
    const float* act_eps;           // [B,A] by path (this step's slice) or null
    const int32_t* elite_pos;       // [B]   by path or null
    const float* state_eps;         // [B,O] by path or null
    const int32_t* row_path;        // [B] compact row -> path, or null (row == path)
    float* raw_tiles;               // scratch, see above
    int* tile_cnt;                  // [ntiles] members finished per row tile (returns to 0)
    int rp_shift;                   // rows per row-math pass = 1 << rp_shift (staging budget)
};

// accessor of the tile-transposed scratch for one row (same interface as RawDyn / RawStaged)
struct RawTile {
    const float* row0; int estride;
    __device__ RawTile(const float* tile_base, int W, int r) : row0(tile_base + r), estride(W * 128) {}
    __device__ float operator()(int e, int c) const { return __ldcg(row0 + e * estride + c * 128); }
    __device__ const float* ptr(int c) const { return row0 + c * 128; }
    __device__ static float load(const float* q) { return __ldcg(q); }
};
