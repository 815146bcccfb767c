"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, never shipped, never on the product path.

A plain-numpy restatement of the reference's model-rollout + GAE hot path
(anyboby/Constrained-Model-Based-Policy-Optimization).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it, and
only as the checker / reported CPU baseline.

Parity status: the reference ships **no tests, fixtures or golden vectors**
(SURVEY.md section 4), so parity is *unpinned by the reference's own tests*.  The
oracle is instead pinned against the reference's UNMODIFIED host code executed in
the build container (`oracle/ref_stubs.py`, `tests/test_oracle_vs_reference.py`)
and against golden vectors generated from it (`oracle/gen_golden.py` ->
`tests/golden/*.npz`).  The two TensorFlow-1.14 graphs on the path (PE forward,
Gaussian actor) cannot run here (TF 1.14 needs Python <= 3.7); they are restated
from the graph-construction code and that restatement is self-pinned only.

Each function cites the reference file:line it follows (paths relative to the
reference root).  All arithmetic is float32 unless the reference promotes.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------
# task tables (models/statics.py:56-70)
# ----------------------------------------------------------------------------
TERM_NO_DONE, TERM_ANTSAFE = 0, 1
COST_ZERO, COST_HCS, COST_ANTSAFE = 0, 1, 2

TASK_TABLE = {
    # task name            (term fn id,   cost fn id)
    "default":            (TERM_NO_DONE, COST_ZERO),
    "HalfCheetah-v2":     (TERM_NO_DONE, COST_ZERO),
    "HalfCheetahSafe-v2": (TERM_NO_DONE, COST_HCS),
    "AntSafe-v2":         (TERM_ANTSAFE, COST_ANTSAFE),
}


def task_fns(task):
    """fake_env.py:134-146 -- unknown tasks fall back to `default` terms and a
    `zeros_like(terms)` cost."""
    return TASK_TABLE.get(task, TASK_TABLE["default"])


# ----------------------------------------------------------------------------
# statics (models/statics.py:3-53)
# ----------------------------------------------------------------------------
def no_done(next_obs):
    """statics.py:3-8: nothing ever terminates."""
    return np.zeros(next_obs.shape[:-1] + (1,), dtype=bool)


def hcs_cost(next_obs):
    """statics.py:10-15: cost 1 while |10 * last obs coordinate| < 2."""
    x = next_obs[..., -1] * 10
    return (np.abs(x) < 2.0).astype(F32)[..., None]


def _antsafe_notdone(next_obs):
    """statics.py:24-27 / 43-46 with the precedence exactly as written: the health
    flags multiply z_rot *before* the comparison with -0.7."""
    z = next_obs[..., 0]
    q = next_obs[..., 1:5]
    z_rot = 1 - 2 * (q[..., 1] ** 2 + q[..., 2] ** 2)
    flags = np.isfinite(next_obs).all(axis=-1) * (z >= 0.2) * (z <= 1.0)
    with np.errstate(invalid="ignore"):
        return (flags * z_rot) >= -0.7


def antsafe_term(next_obs):
    """statics.py:17-31."""
    return (~_antsafe_notdone(next_obs))[..., None]


def antsafe_cost(next_obs):
    """statics.py:33-53 (float64 result, values in {0,1})."""
    obj = np.any(np.abs(next_obs[..., -1:]) > 3.2, axis=-1)[..., None] * 1.0
    done = (~_antsafe_notdone(next_obs))[..., None] * 1.0
    return np.clip(done + obj, 0, 1)


def apply_term(term_id, next_obs):
    return antsafe_term(next_obs) if term_id == TERM_ANTSAFE else no_done(next_obs)


def apply_cost(cost_id, next_obs, terms):
    if cost_id == COST_HCS:
        return hcs_cost(next_obs)
    if cost_id == COST_ANTSAFE:
        return antsafe_cost(next_obs)
    return np.zeros_like(terms)          # fake_env.py:146 -- a *bool* array


# ----------------------------------------------------------------------------
# ensemble disagreement (models/pens/utils.py:15-57)
# ----------------------------------------------------------------------------
KL_EPS = 1e-10


def gaussian_kl(mu0, ls0, mu1, ls1):
    """pens/utils.py:15-25: element-wise KL(N0 || N1), clipped to [0, 1e10]."""
    v0, v1 = np.exp(2 * ls0), np.exp(2 * ls1)
    kl = 0.5 * (((mu1 - mu0) ** 2 + v0) / (v1 + KL_EPS) - 1) + ls1 - ls0
    return np.clip(kl, 0, 1 / KL_EPS)


def average_dkl(mu, std):
    """pens/utils.py:30-57: sum of KL over ALL ordered member pairs (i == j
    included), divided by E(E-1)+eps; accumulated i-outer / j-inner."""
    with np.errstate(divide="ignore"):
        ls = np.clip(np.log(std), -100, 1e8)
    E = len(mu)
    acc = None
    for i in range(E):
        for j in range(E):
            k = gaussian_kl(mu[i], ls[i], mu[j], ls[j])
            acc = k if acc is None else acc + k
    return acc / (E * (E - 1) + KL_EPS)


# ----------------------------------------------------------------------------
# probabilistic ensemble forward (models/pens/pe.py:789-838, fc.py:74-95,
# pens/utils.py:156,167,187)
# ----------------------------------------------------------------------------
def _act(name, x):
    """fc.py:13-20."""
    if name is None:
        return x
    if name == "swish":
        return x * (F32(1) / (F32(1) + np.exp(-x)))
    if name == "tanh":
        return np.tanh(x)
    if name == "ReLU":
        return np.maximum(x, 0)
    if name == "sigmoid":
        return F32(1) / (F32(1) + np.exp(-x))
    raise ValueError(name)


@dataclass
class Ensemble:
    """Weights of one `PE` (pe.py:31-109): E stacked FC layers + optional scalers."""
    W: List[np.ndarray]                 # [E, in, out] per layer (fc.py:135-139)
    b: List[np.ndarray]                 # [E, 1, out]
    acts: List[Optional[str]]           # last entry is None (pe.py:183-185)
    probabilistic: bool                 # NLL / MSPE losses (pe.py:412-414)
    mu_in: Optional[np.ndarray] = None  # [1, in]   (pens/utils.py:104-111)
    var_in: Optional[np.ndarray] = None
    mu_out: Optional[np.ndarray] = None  # [1, D]
    var_out: Optional[np.ndarray] = None
    elite_inds: List[int] = field(default_factory=list)

    @property
    def num_nets(self):
        return self.W[0].shape[0]

    @property
    def in_dim(self):
        return self.W[0].shape[1]

    @property
    def out_dim(self):
        d = self.W[-1].shape[2]
        return d // 2 if self.probabilistic else d


def _sigma(var):
    """pens/utils.py:156: std clamped from below at 1e-2."""
    return np.maximum(np.sqrt(var), F32(1e-2))


# bench.py's CPU-baseline legs set this: the member contraction as ONE multi-threaded BLAS GEMM (einsum optimize=True ->
# tensordot) instead of numpy's single-threaded einsum loops -- ~40x faster on 10 k rows, different rounding.  Parity
# tests and golden fixtures use the default (False), whose op order the fixtures pin bit-exactly.
FAST_GEMM = False


def pe_forward(ens: Ensemble, x):
    """`_compile_outputs(scale_output=True)` (pe.py:789-838).  `x` is [N, in] (every
    member sees the same rows, fc.py:87-88) or [E, N, in] (fc.py:89-90).
    Returns (mean, var) [E, N, D] if probabilistic else mean [E, N, D]."""
    h = np.asarray(x, dtype=F32)
    if ens.mu_in is not None:
        h = (h - ens.mu_in) / _sigma(ens.var_in)
    for W, b, a in zip(ens.W, ens.b, ens.acts):
        if h.ndim == 2:
            h = np.einsum("ij,ajk->aik", h, W, optimize=FAST_GEMM) + b
        else:
            h = np.matmul(h, W) + b
        h = _act(a, h).astype(F32)
    D = ens.out_dim
    mean = h[..., :D]
    if ens.mu_out is not None:
        mean = _sigma(ens.var_out) * mean + ens.mu_out
    if not ens.probabilistic:
        return mean.astype(F32)
    logvar = h[..., D:]
    if ens.mu_out is not None:
        logvar = 2 * np.log(_sigma(ens.var_out)) + logvar
    with np.errstate(over="ignore"):
        var = np.exp(logvar)
    return mean.astype(F32), var.astype(F32)


def pe_predict(ens: Ensemble, x):
    """`PE.predict` (pe.py:648-669 with the tensors of pe.py:326-330, 343): mean over
    ALL members; variance = mean variance + variance of the means."""
    if ens.probabilistic:
        m, v = pe_forward(ens, x)
        mean = m.mean(axis=0)
        return mean, v.mean(axis=0) + np.mean(np.square(m - mean), axis=0)
    return pe_forward(ens, x).mean(axis=0)


# ----------------------------------------------------------------------------
# Gaussian actor + value heads (network/ac_network.py:26-33,46-48,99-123;
# policies/cpo_policy.py:801-835)
# ----------------------------------------------------------------------------
LOGP_EPS = 1e-8          # utilities/utils.py:19


@dataclass
class Actor:
    W: List[np.ndarray]          # [in, out] dense kernels, tanh hidden, linear output
    b: List[np.ndarray]          # [out]
    log_std: np.ndarray          # [A] state-independent (ac_network.py:104)


def actor_mu(actor: Actor, obs):
    """ac_network.py:26-33 via 103: tanh MLP, linear last layer."""
    h = np.asarray(obs, dtype=F32)
    n = len(actor.W)
    for i, (W, b) in enumerate(zip(actor.W, actor.b)):
        h = h @ W + b
        if i < n - 1:
            h = np.tanh(h)
    return h.astype(F32)


def gaussian_logp(x, mu, log_std):
    """ac_network.py:46-48."""
    pre = -0.5 * (((x - mu) / (np.exp(log_std) + F32(LOGP_EPS))) ** 2
                  + 2 * log_std + F32(np.log(2 * np.pi)))
    return pre.sum(axis=1).astype(F32)


def policy_kl(actor: "Actor", obs, old_mu, old_log_std):
    """`CPOPolicy.compute_DKL` for one 2-D batch (cpo_policy.py:837-845): the `d_kl` tensor of
    ac_network.py:114 = gaussian_kl(mu, log_std, old_mu, old_log_std) (ac_network.py:50-55), float32."""
    mu0 = actor_mu(actor, obs)
    ls0 = np.asarray(actor.log_std, F32)
    mu1, ls1 = np.asarray(old_mu, F32), np.asarray(old_log_std, F32)
    var0, var1 = np.exp(F32(2) * ls0), np.exp(F32(2) * ls1)
    pre = F32(0.5) * (((mu1 - mu0) ** 2 + var0) / (var1 + F32(LOGP_EPS)) - F32(1)) + ls1 - ls0
    return F32(np.mean(np.sum(pre.astype(F32), axis=1, dtype=F32), dtype=F32))


def epochs_list(epoch_archive):
    """cpobuffer.py:149-154."""
    bins = np.bincount(epoch_archive[epoch_archive >= 0])
    return np.squeeze(np.nonzero(bins), axis=0)


def boltz_dist(epoch_archive, kls, alpha=1):
    """`CPOBuffer.boltz_dist` (cpobuffer.py:385-396), same numpy operations."""
    ep_probs = np.exp(alpha * np.negative(kls))
    ep_probs /= np.sum(ep_probs)
    sample_p = np.bincount(epoch_archive[epoch_archive >= 0]).astype(np.float32)
    sample_p[sample_p > 0] = ep_probs / sample_p[sample_p > 0]
    return np.where(epoch_archive >= 0, sample_p[epoch_archive], 0)


class OraclePolicy:
    """Inference side of `CPOPolicy` (cpo_policy.py:801-835).  `eps_fn(n)` supplies the
    standard-normal draws `tf.random_normal` would (ac_network.py:109)."""

    class _Agent:
        reward_penalized = False      # cpo_policy.py:369

    def __init__(self, actor: Actor, v: Ensemble, vc: Ensemble, eps_fn):
        self.actor, self.v, self.vc = actor, v, vc
        self.eps_fn = eps_fn
        self.agent = self._Agent()
        self.ctx = None               # (step, global path ids) set by the sampler

    def reset(self):
        pass

    def get_action_outs(self, obs):
        obs = np.asarray(obs, dtype=F32)
        mu = actor_mu(self.actor, obs)
        ls = self.actor.log_std.astype(F32)
        eps = np.asarray(self.eps_fn(self, obs.shape[0]), dtype=F32)
        pi = mu + eps * np.exp(ls)                       # ac_network.py:109
        log_std_b = np.outer(np.ones(obs.shape[0], F32), ls)   # ac_network.py:119
        return {"pi": pi.astype(F32), "logp_pi": gaussian_logp(pi, mu, ls),
                "pi_info": {"mu": mu, "log_std": log_std_b},
                "v": self.get_v(obs), "vc": self.get_vc(obs)}

    def get_v(self, obs):
        return np.squeeze(pe_predict(self.v, np.asarray(obs, F32)), axis=-1)

    def get_vc(self, obs):
        return np.squeeze(pe_predict(self.vc, np.asarray(obs, F32)), axis=-1)


# ----------------------------------------------------------------------------
# FakeEnv.step (models/fake_env.py:66-172), 2-D inputs
# ----------------------------------------------------------------------------
class OracleModel:
    """The duck-typed model contract FakeEnv consumes (models/base_model.py:3-43)."""

    def __init__(self, ens: Ensemble):
        self.ens = ens
        self.is_ensemble = ens.num_nets > 1
        self.is_probabilistic = ens.probabilistic
        self.in_dim = ens.in_dim
        self.out_dim = None if ens.probabilistic else ens.out_dim   # pe.py:425-430 quirk
        self.elite_inds = list(ens.elite_inds)

    def predict_ensemble(self, x):
        return pe_forward(self.ens, x)

    def predict(self, x):
        return pe_predict(self.ens, x)


class OracleFakeEnv:
    """fake_env.py:15-198 for an ensemble, probabilistic, delta-predicting model.
    `idx_fn(env, n)` returns the n elite *positions* `np.random.choice(elite_inds, n)`
    would draw (fake_env.py:174-176); `state_eps_fn(env, n)` (optional) supplies the
    multiplier of std in non-deterministic mode -- the reference uses 1 (fake_env.py:105-106)."""

    def __init__(self, obs_dim, act_dim, task, model: OracleModel, idx_fn,
                 predicts_delta=True, predicts_rew=True, predicts_cost=False,
                 state_eps_fn=None):
        self.obs_dim, self.act_dim, self.task = obs_dim, act_dim, task
        self.model, self.idx_fn, self.state_eps_fn = model, idx_fn, state_eps_fn
        self.predicts_delta, self.predicts_rew, self.predicts_cost = \
            predicts_delta, predicts_rew, predicts_cost
        self.term_id, self.cost_id = task_fns(task)
        self.ctx = None

    def step(self, obs, act, deterministic=True):
        obs = np.asarray(obs, F32)
        act = np.asarray(act, F32)
        O = self.obs_dim
        x = np.concatenate((obs, act), axis=-1)                       # :81
        mean, var = self.model.predict_ensemble(x)                    # :88-97
        std = np.sqrt(var)                                            # :104
        nxt = mean[..., :O]
        if not deterministic:                                         # :105-108
            mult = 1.0 if self.state_eps_fn is None else np.asarray(
                self.state_eps_fn(self, obs.shape[0]), F32)
            nxt = nxt + std[..., :O] * mult
        ep_var = np.var(nxt, axis=0)                                  # :112
        dkl_path = np.mean(average_dkl(nxt, std[..., :O]), axis=-1)   # :113
        dkl_mean = np.mean(dkl_path)                                  # :114
        n = obs.shape[0]
        pos = np.asarray(self.idx_fn(self, n))
        members = np.asarray(self.model.elite_inds)[pos]              # :121-123
        rows = np.arange(n)
        nxt = nxt[members, rows]                                      # :125
        if self.predicts_delta:
            nxt = nxt + obs                                           # :130-131
        terms = apply_term(self.term_id, nxt)                         # :134-137
        if self.predicts_cost:                                        # :139-142
            cost = mean[..., -1:][members, rows]
            mean = mean[..., :-1]
        else:
            cost = apply_cost(self.cost_id, nxt, terms)               # :143-146
        rew = mean[..., -1:][members, rows]                           # :148-151
        info = {"ensemble_dkl_mean": dkl_mean, "ensemble_dkl_path": dkl_path,
                "ensemble_ep_var": ep_var, "rew": rew, "cost": cost}
        return nxt.astype(F32), rew, terms, info


# ----------------------------------------------------------------------------
# GAE (utilities/utils.py:159-211 branch 186-188; modelbuffer.py:163-179;
# cpobuffer.py:179-207)
# ----------------------------------------------------------------------------
def discount_cumsum(x, discount, lam):
    """y_t = x_t + (discount*lam) * y_{t+1} along the last axis.

    The reference flips, runs `scipy.signal.lfilter([1],[1,-d])` and flips back.
    lfilter's direct-form-II-transposed recurrence for these coefficients is, in
    float64 and strictly sequentially, y[n] = x[n] + z ; z = d * y[n]; restated
    here so the GPU kernel has an exact op order to match."""
    x = np.asarray(x)
    y = np.empty(x.shape, dtype=np.float64)
    if x.size == 0:
        return np.array(x)
    d = float(discount * lam)
    z = np.zeros(x.shape[:-1], dtype=np.float64)
    for t in range(x.shape[-1] - 1, -1, -1):
        cur = x[..., t].astype(np.float64) + z
        y[..., t] = cur
        z = d * cur
    return y


def gae_path(rew, val, cost, cval, last_val, last_cval, gamma, lam, cgamma, clam):
    """One finished path slice (modelbuffer.py:163-179 / cpobuffer.py:180-198).
    Inputs float32 [..., T]; last_* [...].  Returns float32 adv, ret, cadv, cret --
    the float64 scan output is rounded when stored into the float32 buffers and
    `ret` is formed from the *rounded* advantage."""
    vals = np.append(val, np.asarray(last_val)[..., None], axis=-1)
    deltas = rew + gamma * vals[..., 1:] - vals[..., :-1]
    adv = discount_cumsum(deltas, gamma, lam).astype(F32)
    ret = (adv + val).astype(F32)
    cvals = np.append(cval, np.asarray(last_cval)[..., None], axis=-1)
    cdeltas = cost + cgamma * cvals[..., 1:] - cvals[..., :-1]
    cadv = discount_cumsum(cdeltas, cgamma, clam).astype(F32)
    cret = (cadv + cval).astype(F32)
    return adv, ret, cadv, cret


def stats_scalar(x):
    """utilities/mpi_tools.py:71-92 at world size 1: float32 two-pass mean / std."""
    x = np.array(x, dtype=F32)
    s, n = np.asarray([np.sum(x), len(x)], dtype=F32)
    mean = s / n
    ssq = np.asarray([np.sum((x - mean) ** 2)], dtype=F32)[0]
    return mean, np.sqrt(ssq / n)


ADV_EPS = 1e-8   # utilities/utils.py:19


# ----------------------------------------------------------------------------
# ModelBuffer (buffers/modelbuffer.py:18-226)
# ----------------------------------------------------------------------------
class OracleModelBuffer:
    FIELDS = ("obs", "act", "nextobs", "rew", "val", "cost", "cval", "logp", "dyn_error",
              "adv", "ret", "cadv", "cret")

    def __init__(self, batch_size, obs_dim, act_dim, max_path_length):
        self.B, self.O, self.A, self.T = batch_size, obs_dim, act_dim, max_path_length
        self.pi_keys = ()
        self.gamma = self.lam = self.cost_gamma = self.cost_lam = None
        self.reset()

    def initialize(self, pi_info_shapes, gamma=0.99, lam=0.95, cost_gamma=0.99, cost_lam=0.95):
        self.pi_shapes = dict(pi_info_shapes)
        self.pi_keys = tuple(sorted(pi_info_shapes))
        self.gamma, self.lam, self.cost_gamma, self.cost_lam = gamma, lam, cost_gamma, cost_lam
        self.reset()

    def reset(self, batch_size=None):                                 # :53-98
        if batch_size is not None:
            self.B = batch_size
        B, T = self.B, self.T
        z = lambda *s: np.zeros((B, T) + s, dtype=F32)
        self.obs, self.nextobs, self.act = z(self.O), z(self.O), z(self.A)
        for k in ("rew", "val", "cost", "cval", "logp", "dyn_error", "adv", "ret", "cadv", "cret"):
            setattr(self, k, z())
        self.term = np.zeros((B, T), dtype=bool)
        self.pi = {k: z(*self.pi_shapes[k]) for k in self.pi_keys}
        self.ptr = 0
        self.populated = np.zeros((B, T), dtype=bool)
        self.terminated = np.zeros(B, dtype=bool)

    @property
    def size(self):
        return self.populated.sum()

    @property
    def has_room(self):
        return self.ptr < self.T

    @property
    def alive_paths(self):
        return ~self.terminated

    def store_multiple(self, obs, act, next_obs, rew, val, cost, cval, dyn_error, logp, pi_info, term):
        """:114-135 -- row i of every argument goes to the i-th alive path, column ptr."""
        assert self.ptr < self.T
        a, p = self.alive_paths, self.ptr
        self.obs[a, p], self.act[a, p], self.nextobs[a, p] = obs, act, next_obs
        self.rew[a, p], self.val[a, p], self.cost[a, p], self.cval[a, p] = rew, val, cost, cval
        self.logp[a, p], self.term[a, p], self.dyn_error[a, p] = logp, term, dyn_error
        for k in self.pi_keys:
            self.pi[k][a, p] = pi_info[k]
        self.populated[a, p] = True
        self.ptr += 1

    def finish_path_multiple(self, term_mask, last_val, last_cval):
        """:138-182."""
        term_mask = np.asarray(term_mask, dtype=bool)
        if not term_mask.any():
            return
        alive_idx = np.flatnonzero(self.alive_paths)
        assert len(alive_idx) == len(term_mask)
        fin = np.zeros(self.B, dtype=bool)
        fin[alive_idx[term_mask]] = True
        if self.ptr > 0:
            s = slice(0, self.ptr)
            adv, ret, cadv, cret = gae_path(
                self.rew[fin, s], self.val[fin, s], self.cost[fin, s], self.cval[fin, s],
                last_val, last_cval, self.gamma, self.lam, self.cost_gamma, self.cost_lam)
            self.adv[fin, s], self.ret[fin, s] = adv, ret
            self.cadv[fin, s], self.cret[fin, s] = cadv, cret
        self.terminated |= fin

    def get(self):
        """:184-226.  List order is the contract of cpo_policy.py:472-477."""
        assert self.terminated.all()
        m = self.populated
        if self.size > 0:
            mean, std = stats_scalar(self.adv[m].flatten())
            self.adv[m] = (self.adv[m] - mean) / (std + ADV_EPS)
            cmean, _ = stats_scalar(self.cadv[m].flatten())
            self.cadv[m] -= cmean
            ret_mean, cret_mean = self.ret[m].mean(), self.cret[m].mean()
        else:
            ret_mean = cret_mean = 0
        bufs = [self.obs, self.act, self.adv, self.cadv, self.ret, self.cret, self.logp,
                self.val, self.cval, self.cost] + [self.pi[k] for k in self.pi_keys]
        out = [b[m] for b in bufs]
        diag = dict(poolm_batch_size=m.sum(), poolm_ret_mean=ret_mean, poolm_cret_mean=cret_mean)
        self.reset()
        return out, diag


# ----------------------------------------------------------------------------
# CPOBuffer.finish_path / get (buffers/cpobuffer.py:160-207, 249-290), flat layout
# ----------------------------------------------------------------------------
def cpobuffer_gae_flat(rew, val, cost, cval, seg_offsets, last_vals, last_cvals,
                       gamma, lam, cgamma, clam):
    """A sequence of `finish_path` calls over consecutive slices [off[i], off[i+1])."""
    n = len(rew)
    adv, ret, cadv, cret = (np.zeros(n, F32) for _ in range(4))
    for i in range(len(seg_offsets) - 1):
        s = slice(int(seg_offsets[i]), int(seg_offsets[i + 1]))
        if s.stop == s.start:
            continue
        adv[s], ret[s], cadv[s], cret[s] = gae_path(
            rew[s], val[s], cost[s], cval[s], F32(last_vals[i]), F32(last_cvals[i]),
            gamma, lam, cgamma, clam)
    return adv, ret, cadv, cret


def cpobuffer_normalise(adv, cadv):
    """cpobuffer.py:262-268."""
    mean, std = stats_scalar(adv)
    cmean, _ = stats_scalar(cadv)
    return ((adv - mean) / (std + ADV_EPS)).astype(F32), (cadv - cmean).astype(F32)


# ----------------------------------------------------------------------------
# ModelSampler (samplers/model_sampler.py:203-444)
# ----------------------------------------------------------------------------
SAMPLER_EPS = 1e-8


class OracleModelSampler:
    def __init__(self, max_path_length, batch_size=1000, rollout_mode=False):
        self.max_path_length = max_path_length
        self.batch_size = batch_size
        self.rollout_mode = rollout_mode
        self.dkl_lim = None
        self.obs = None

    def initialize(self, env, policy, pool):
        self.env, self.policy, self.pool = env, policy, pool

    def set_rollout_dkl(self, dkl):
        self.dkl_lim = dkl

    def set_max_path_length(self, n):
        self.max_path_length = n

    def reset(self, observations):                                    # :203-237
        self.batch_size = observations.shape[0]
        self.obs = np.asarray(observations)
        self.policy.reset()
        self.pool.reset(self.batch_size)
        B = self.batch_size
        self.path_length = np.zeros(B)
        self.path_return = np.zeros(B)
        self.path_cost = np.zeros(B)
        self.path_dyn_var = np.zeros(B)
        self.dkl_path = np.zeros(B)
        self.total_samples = 0
        self.n_episodes = 0
        self.total_Vs = self.total_CVs = self.total_cost = self.total_rew = 0
        self.total_dyn_ep_var = self.total_dkl = self.max_dkl = 0
        self.max_path_return = 0

    def _ids(self):
        return np.flatnonzero(self.pool.alive_paths)

    def compute_dynamics_dkl(self, obs_batch, depth=1):               # :151-167
        """Mean ensemble disagreement over `depth` model steps from `obs_batch`, times depth; the rows of
        terminated paths are dropped between steps.  Needs reset() first (the accumulators of :203-237)."""
        obs_batch = np.asarray(obs_batch)
        ids = np.arange(obs_batch.shape[0])
        for step in range(depth):
            self.policy.ctx = self.env.ctx = (step, ids)
            a = self.policy.get_action_outs(obs_batch)["pi"]
            next_obs, _, terminal, info = self.env.step(obs_batch, a)
            n_paths = next_obs.shape[0]
            self.total_dkl += info.get("ensemble_dkl_mean", 0) * n_paths
            self.total_samples += n_paths
            keep = np.squeeze(~terminal, -1) if terminal.ndim > 1 else ~terminal
            obs_batch, ids = next_obs[keep], ids[keep]
        return self.total_dkl / (self.total_samples + SAMPLER_EPS) * depth

    def sample(self, max_samples=None):                               # :239-375
        pool = self.pool
        assert pool.has_room and self.obs is not None and pool.alive_paths.any()
        step = self.n_episodes
        self.n_episodes += 1
        alive = pool.alive_paths
        cur = self.obs
        self.policy.ctx = self.env.ctx = (step, self._ids())
        outs = self.policy.get_action_outs(cur)
        a, logp, pi_info, v, vc = outs["pi"], outs["logp_pi"], outs["pi_info"], outs["v"], outs["vc"]
        nxt, rew, term, info = self.env.step(cur, a)
        rew = np.squeeze(rew, axis=-1)
        c = np.squeeze(info["cost"])
        term = np.squeeze(term, axis=-1)
        dkl_mean, dkl_path, ep_var = (info["ensemble_dkl_mean"], info["ensemble_dkl_path"],
                                      info["ensemble_ep_var"])
        if self.rollout_mode == "uncertainty":                        # :275-279
            cut = (self.dkl_path[alive] + dkl_path) >= self.dkl_lim
        else:
            cut = np.zeros(alive.sum(), dtype=bool)
        if max_samples:                                               # :282-287
            n = self.total_samples + alive.sum() - cut.sum()
            n = max(n - max_samples, 0)
            early = np.zeros((~cut).sum(), dtype=bool)
            early[:n] = True
            cut[~cut] = early
        keep = self._finish(cut, True, True)                          # :290
        alive = pool.alive_paths
        if not alive.any():
            info["alive_ratio"] = 0
            return nxt, rew, term, info
        cur, a, nxt, rew, v, c, vc, term, dkl_path, logp, ep_var = (
            arr[keep] for arr in (cur, a, nxt, rew, v, c, vc, term, dkl_path, logp, ep_var))
        pi_info = {k: val[keep] for k, val in pi_info.items()}
        n_alive = alive.sum()
        self.total_samples += n_alive                                 # :314-333
        self.total_cost += c.sum()
        self.total_rew += rew.sum()
        self.path_return[alive] += rew
        self.path_cost[alive] += c
        self.path_length[alive] += 1
        self.path_dyn_var[alive] += np.mean(ep_var, axis=-1)
        self.total_dyn_ep_var += ep_var.sum()
        self.total_Vs += v.sum()
        self.total_CVs += vc.sum()
        self.total_dkl += dkl_mean * n_alive
        self.max_dkl = max(self.max_dkl, np.max(dkl_path))
        self.dkl_path[alive] += dkl_path
        self.max_path_return = max(self.max_path_return, np.max(self.path_return))
        pool.store_multiple(cur, a, nxt, rew, v, c, vc, np.mean(ep_var, axis=-1),
                            logp, pi_info, term)                      # :336-346
        self.obs = nxt                                                # :350
        end = (self.path_length >= self.max_path_length - 1)[alive]   # :352
        keep = self._finish(end, True, True)
        if not keep.any():
            info["alive_ratio"] = 0
            return nxt, rew, term, info
        self.obs = self.obs[keep]
        keep = self._finish(term, False, True)                        # :357-364
        if not keep.any():
            info["alive_ratio"] = 0
            return nxt, rew, term, info
        self.obs = self.obs[keep]
        info["alive_ratio"] = pool.alive_paths.sum() / self.batch_size
        return nxt, rew, term, info

    def _finish(self, mask, boot_v, boot_vc):                         # :377-416
        mask = np.asarray(mask, dtype=bool)
        if not mask.any():
            return ~mask
        lv = self.policy.get_v(self.obs[mask]) if boot_v else np.zeros(mask.sum())
        lc = self.policy.get_vc(self.obs[mask]) if boot_vc else np.zeros(mask.sum())
        self.pool.finish_path_multiple(mask, lv, lc)
        return ~mask

    def finish_all_paths(self):                                       # :418-444
        alive = self.pool.alive_paths
        if alive.any():
            mask = np.ones(alive.sum(), dtype=bool)
            self.pool.finish_path_multiple(mask, self.policy.get_v(self.obs),
                                           self.policy.get_vc(self.obs))
        return self.diagnostics()

    def diagnostics(self):                                            # :89-133
        n = self.total_samples + SAMPLER_EPS
        return {
            "msampler/samples_added": self.total_samples,
            "msampler/rollout_H_max": self.n_episodes,
            "msampler/rollout_H_mean": self.total_samples / (self.batch_size + SAMPLER_EPS),
            "msampler/dyn_var_perstep": self.total_dyn_ep_var / n,
            "msampler/cost_rate": np.sum(self.path_cost) / n,
            "msampler/rew_rate": np.sum(self.path_return) / n,
            "msampler/v_mean": self.total_Vs / n,
            "msampler/cv_mean": self.total_CVs / n,
            "msampler/ens_DKL": self.total_dkl / n,
            "msampler/max_path_return": self.max_path_return,
            "msampler/max_dkl": self.max_dkl,
        }


# ----------------------------------------------------------------------------
# synthetic weights / inputs (fc.py:135-139 initialiser)
# ----------------------------------------------------------------------------
def _trunc_normal(rng, shape, std):
    """tf.truncated_normal_initializer: resample beyond two standard deviations."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return (x * std).astype(F32)


def make_ensemble(rng, in_dim, out_dim, hidden, num_nets, num_elites, probabilistic,
                  act="swish", scalers=None, gain=1.0, bias_std=0.0):
    """Weights shaped like `build_PE` would create them (pe_factory.py:50-55;
    fc.py:135-139: trunc-normal std 1/(2 sqrt(in)), zero bias).  `gain`>1 rescales the
    weights so activations are O(1) (a trained net is not at initialisation scale).
    `scalers` = (mu_in, var_in, mu_out, var_out) or None for the identity scalers a
    fresh `TensorStandardScaler` holds (pens/utils.py:104-111)."""
    last = 2 * out_dim if probabilistic else out_dim
    dims = [in_dim] + list(hidden) + [last]
    W, b, acts = [], [], []
    for i in range(len(dims) - 1):
        W.append(_trunc_normal(rng, (num_nets, dims[i], dims[i + 1]),
                               gain / (2 * np.sqrt(dims[i]))))
        b.append((bias_std * rng.standard_normal((num_nets, 1, dims[i + 1]))).astype(F32))
        acts.append(act if i < len(dims) - 2 else None)
    if scalers is None:
        scalers = (np.zeros((1, in_dim)), np.ones((1, in_dim)),
                   np.zeros((1, out_dim)), np.ones((1, out_dim)))
    mu_in, var_in, mu_out, var_out = (np.asarray(a, F32).reshape(1, -1) for a in scalers)
    elites = [int(i) for i in rng.permutation(num_nets)[:num_elites]]
    return Ensemble(W, b, acts, probabilistic, mu_in, var_in, mu_out, var_out, elites)


def make_scalers(rng, task, obs_dim, act_dim):
    """Data-like scalers: per-dimension obs scale s in [0.05, 3] with two nearly constant
    dimensions (variance 1e-6 -> exercises the 1e-2 sigma clamp of pens/utils.py:156),
    model outputs (state deltas) 20x smaller than the state scale so H-step rollouts stay
    in range, and task-specific placement of the coordinates the statics read
    (statics.py:13, 20-22, 40)."""
    O, A = obs_dim, act_dim
    s = np.exp(rng.uniform(np.log(0.05), np.log(3.0), O))
    mu = rng.standard_normal(O) * s
    flat = rng.choice(np.arange(5, O - 1), size=2, replace=False)
    s[flat] = 1e-3
    mu[flat] = 0.5 * rng.standard_normal(2)
    if task == "AntSafe-v2":
        mu[0], s[0] = 0.6, 0.12                   # torso height, healthy in [0.2, 1.0]
        mu[1:5], s[1:5] = (0.8, 0.2, 0.2, 0.1), (0.1, 0.45, 0.45, 0.1)
        mu[-1], s[-1] = 0.0, 2.5                  # y distance, cost beyond 3.2
    else:
        mu[-1], s[-1] = 0.0, 0.3                  # HCS: cost while |x| < 0.2
    mu_in = np.concatenate([mu, np.zeros(A)])
    var_in = np.concatenate([s ** 2, np.full(A, 1.0 / 3.0)])
    mu_out = np.concatenate([0.01 * s * rng.standard_normal(O), [0.5]])
    var_out = np.concatenate([(0.05 * s) ** 2, [1.0]])
    if task == "AntSafe-v2":                      # let height / tilt drift into termination
        var_out[0], var_out[2], var_out[3] = 0.05 ** 2, 0.2 ** 2, 0.2 ** 2
    obs_scalers = (mu, s ** 2)
    return (mu_in, var_in, mu_out, var_out), obs_scalers


def make_actor(rng, obs_dim, act_dim, hidden=(128, 128), gain=1.0):
    """Glorot-uniform dense kernels (tf.layers.dense default), log_std = -0.5
    (ac_network.py:104)."""
    dims = [obs_dim] + list(hidden) + [act_dim]
    W, b = [], []
    for i in range(len(dims) - 1):
        lim = gain * np.sqrt(6.0 / (dims[i] + dims[i + 1]))
        W.append(rng.uniform(-lim, lim, (dims[i], dims[i + 1])).astype(F32))
        b.append(np.zeros(dims[i + 1], F32))
    return Actor(W, b, np.full(act_dim, -0.5, F32))


def make_problem(seed, obs_dim, act_dim, hidden=(512, 512), num_nets=7, num_elites=5,
                 vf_nets=3, vf_hidden=(128, 128), a_hidden=(128, 128),
                 task="HalfCheetahSafe-v2", data_scalers=True, gain=2.0):
    """All weights of one CMBPO rollout problem: dynamics PE (cmbpo.py:120-136),
    actor, V and VC ensembles (cpo_policy.py:452-468)."""
    rng = np.random.default_rng(seed)
    dyn_sc = v_sc = vc_sc = None
    if data_scalers:
        dyn_sc, (mu_o, var_o) = make_scalers(rng, task, obs_dim, act_dim)
        v_sc = (mu_o, var_o, [3.0], [25.0])
        vc_sc = (mu_o, var_o, [1.0], [4.0])
    dyn = make_ensemble(rng, obs_dim + act_dim, obs_dim + 1, hidden, num_nets, num_elites,
                        True, scalers=dyn_sc, gain=gain)
    actor = make_actor(rng, obs_dim, act_dim, a_hidden)
    if data_scalers:      # the actor sees raw observations: fold (obs-mu)/sigma into layer 0
        sig = np.maximum(np.sqrt(var_o), 0.1)     # the actor has no scaler: keep its gains moderate
        w0 = actor.W[0].astype(np.float64)
        actor.W[0] = (w0 / sig[:, None]).astype(F32)
        actor.b[0] = (-(mu_o / sig) @ w0).astype(F32)
    v = make_ensemble(rng, obs_dim, 1, vf_hidden, vf_nets, 2, False, scalers=v_sc, gain=gain)
    vc = make_ensemble(rng, obs_dim, 1, vf_hidden, vf_nets, 2, False, scalers=vc_sc, gain=gain)
    return dyn, actor, v, vc


def make_states(seed, n, obs_dim, act_dim, dyn: Optional[Ensemble] = None):
    """Synthetic rows: obs ~ N(mu_in, sigma_in), act ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((n, obs_dim)).astype(F32)
    if dyn is not None and dyn.mu_in is not None:
        obs = obs * _sigma(dyn.var_in)[:, :obs_dim] + dyn.mu_in[:, :obs_dim]
    act = rng.uniform(-1, 1, (n, act_dim)).astype(F32)
    return obs.astype(F32), act


class TableNoise:
    """Noise keyed by (step, global path id) so any execution order sees the same draws."""

    def __init__(self, seed, steps, n_paths, act_dim, n_elite, obs_dim=0):
        rng = np.random.default_rng(seed)
        self.act_eps = rng.standard_normal((steps, n_paths, act_dim)).astype(F32)
        self.elite_pos = rng.integers(0, n_elite, (steps, n_paths)).astype(np.int32)
        self.state_eps = (rng.standard_normal((steps, n_paths, obs_dim)).astype(F32)
                          if obs_dim else None)

    def eps_fn(self, policy, n):
        t, ids = policy.ctx
        assert len(ids) == n
        return self.act_eps[t, ids]

    def idx_fn(self, env, n):
        t, ids = env.ctx
        assert len(ids) == n
        return self.elite_pos[t, ids]

    def state_eps_fn(self, env, n):
        t, ids = env.ctx
        return self.state_eps[t, ids]


def run_rollout(dyn, actor, v, vc, task, start_obs, noise: TableNoise, max_path_length,
                rollout_mode=False, dkl_lim=None, max_samples=None, stop_alive_ratio=None,
                gamma=0.99, lam=0.95, cgamma=0.97, clam=0.5, n_steps=None):
    """reset -> sample x n -> finish_all_paths -> get (cmbpo.py:251-269)."""
    B, O = start_obs.shape
    A = actor.W[-1].shape[1]
    policy = OraclePolicy(actor, v, vc, noise.eps_fn)
    env = OracleFakeEnv(O, A, task, OracleModel(dyn), noise.idx_fn)
    pool = OracleModelBuffer(B, O, A, max_path_length)
    pool.initialize({"mu": (A,), "log_std": (A,)}, gamma, lam, cgamma, clam)
    smp = OracleModelSampler(max_path_length, B, rollout_mode)
    smp.initialize(env, policy, pool)
    smp.set_rollout_dkl(dkl_lim)
    smp.reset(start_obs)
    steps = 0
    while pool.alive_paths.any() and pool.has_room:
        if n_steps is not None and steps >= n_steps:
            break
        _, _, _, info = smp.sample(max_samples)
        steps += 1
        if stop_alive_ratio is not None and info["alive_ratio"] <= stop_alive_ratio:
            break
    diag = smp.finish_all_paths()
    snapshot = {k: getattr(pool, k).copy() for k in
                ("obs", "act", "nextobs", "rew", "val", "cost", "cval", "logp", "dyn_error",
                 "adv", "ret", "cadv", "cret", "term", "populated")}
    snapshot["mu"], snapshot["log_std"] = pool.pi["mu"].copy(), pool.pi["log_std"].copy()
    out, bdiag = pool.get()
    return out, bdiag, diag, snapshot
