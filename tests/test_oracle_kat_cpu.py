"""Narrows the "parity unpinned" part of the oracle -- the two TF-1.14 graphs that cannot run here
(PE forward, models/pens/pe.py:789-838 + fc.py:74-95 + pens/utils.py:156-187; Gaussian actor,
network/ac_network.py:26-48,99-123) -- with two checks that do not share code with oracle/cmbpo_oracle.py:

 1. a hand-computed known-answer table for a 2-member toy network (derived step by step from the
    reference's formulas; the arithmetic is reproduced in the comments and in plain `math` calls);
 2. an independent float64 restatement written straight from the reference lines (scalar loops, no numpy
    broadcasting tricks, no helper of the oracle) compared with the oracle on random ensembles.

A shared misreading of the graph code would have to be made twice, in two different forms, to pass."""
import math

import numpy as np

from oracle import cmbpo_oracle as orc


# ------------------------------------------------------------------------------------------------
# 1. toy KAT: E = 2, in = 2, one hidden layer of 2 swish units, probabilistic output D = 1 (mean | logvar)
# ------------------------------------------------------------------------------------------------
def _toy():
    W0 = np.array([[[1.0, 0.0], [0.0, 1.0]],            # member 0: identity
                   [[0.5, -1.0], [2.0, 0.0]]], np.float32)          # member 1: [in, out]
    b0 = np.array([[[0.0, 0.0]], [[0.0, 1.0]]], np.float32)
    W1 = np.array([[[1.0, 0.0], [1.0, 1.0]],            # [hidden, 2D]: column 0 = mean, column 1 = log-variance
                   [[2.0, 0.5], [0.0, -1.0]]], np.float32)
    b1 = np.array([[[0.5, 0.0]], [[0.0, 0.0]]], np.float32)
    return orc.Ensemble(W=[W0, W1], b=[b0, b1], acts=["swish", None], probabilistic=True,
                        mu_in=np.array([[1.0, 0.0]], np.float32), var_in=np.array([[4.0, 1e-6]], np.float32),
                        mu_out=np.array([[10.0]], np.float32), var_out=np.array([[9.0]], np.float32),
                        elite_inds=[0, 1])


def _swish(x):
    return x / (1.0 + math.exp(-x))


def test_pe_forward_toy_known_answers():
    ens = _toy()
    x = np.array([[3.0, 0.02]], np.float32)
    # input scaler (pens/utils.py:156): sigma = max(sqrt(var), 1e-2) = (2, 1e-2)  [sqrt(1e-6) = 1e-3 is clamped]
    #   z = ((3 - 1) / 2, (0.02 - 0) / 1e-2) = (1, 2)
    # member 0: h = swish(z) = (1 * sigmoid(1), 2 * sigmoid(2)) = (0.7310585786, 1.7615941560)
    #   out = h @ W1 + b1 = (h0 + h1 + 0.5, h1) = (2.9926527346, 1.7615941560)
    # member 1: pre = z @ W0 + b0 = (0.5*1 + 2*2, -1*1 + 0*2 + 1) = (4.5, 0.0);  h = (4.5 * sigmoid(4.5), 0) = (4.4505587582, 0)
    #   out = (2 * h0, 0.5 * h0) = (8.9011175163, 2.2252793791)
    # output scaler (utils.py:167,187): sigma_out = 3: mean = 3 * out_mean + 10 ; logvar = 2 ln 3 + out_logvar ; var = exp(logvar)
    h00, h01 = _swish(1.0), _swish(2.0)
    h10 = _swish(4.5)
    want_mean = [3.0 * (h00 + h01 + 0.5) + 10.0, 3.0 * (2.0 * h10) + 10.0]
    want_var = [9.0 * math.exp(h01), 9.0 * math.exp(0.5 * h10)]
    assert abs(want_mean[0] - 18.9779582038) < 1e-9 and abs(want_mean[1] - 36.7033525490) < 1e-9
    assert abs(want_var[0] - 52.39540) < 1e-4 and abs(want_var[1] - 83.30462) < 1e-4
    mean, var = orc.pe_forward(ens, x)
    assert mean.shape == (2, 1, 1) and var.shape == (2, 1, 1)
    assert np.allclose(mean[:, 0, 0], want_mean, rtol=3e-7)
    assert np.allclose(var[:, 0, 0], want_var, rtol=3e-6)
    # PE.predict (pe.py:326-343): mean over members; variance = mean of variances + variance of the means
    pm, pv = orc.pe_predict(ens, x)
    mbar = 0.5 * (want_mean[0] + want_mean[1])
    vbar = 0.5 * (want_var[0] + want_var[1]) + 0.5 * ((want_mean[0] - mbar) ** 2 + (want_mean[1] - mbar) ** 2)
    assert np.allclose(pm[0, 0], mbar, rtol=3e-7) and np.allclose(pv[0, 0], vbar, rtol=3e-6)


def test_actor_toy_known_answers():
    # tanh MLP with a linear last layer (ac_network.py:26-33); logp (ac_network.py:46-48)
    W = [np.array([[1.0, -1.0], [0.5, 2.0]], np.float32), np.array([[2.0], [1.0]], np.float32)]
    b = [np.array([0.0, 0.5], np.float32), np.array([-1.0], np.float32)]
    actor = orc.Actor(W=W, b=b, log_std=np.array([-0.5], np.float32))
    obs = np.array([[1.0, 2.0]], np.float32)
    # hidden pre = (1*1 + 2*0.5, -1*1 + 2*2 + 0.5) = (2, 3.5); h = tanh -> (0.9640275801, 0.9981778976)
    # mu = 2 * h0 + h1 - 1 = 1.9262330578
    mu = 2.0 * math.tanh(2.0) + math.tanh(3.5) - 1.0
    assert abs(mu - 1.9262330578) < 1e-9
    got = orc.actor_mu(actor, obs)
    assert np.allclose(got, [[mu]], rtol=3e-7)
    # pi = mu + eps * exp(log_std) with eps = 2: std = e^-0.5 = 0.6065306597; pi = mu + 1.2130613194
    # logp = -0.5 * (((pi - mu) / (std + 1e-8))^2 + 2 * (-0.5) + ln(2 pi)) = -0.5 * (4 - 1 + 1.8378770664) = -2.4189385332
    pi = np.array([[mu + 2.0 * math.exp(-0.5)]], np.float32)
    lp = orc.gaussian_logp(pi, got, actor.log_std)
    assert np.allclose(lp, [-2.4189385332], rtol=2e-6)


# ------------------------------------------------------------------------------------------------
# 2. independent float64 restatement, scalar loops, written from the reference lines
# ------------------------------------------------------------------------------------------------
def _fc64(x_row, W, b, act):
    """fc.py:87-95 for ONE member and ONE row: raw = sum_i x_i W[i, k] + b[k]; swish = x * sigmoid(x)."""
    out = []
    for k in range(W.shape[1]):
        s = 0.0
        for i in range(W.shape[0]):
            s += float(x_row[i]) * float(W[i, k])
        s += float(b[k])
        if act == "swish":
            s = s * (1.0 / (1.0 + math.exp(-s)))
        elif act == "tanh":
            s = math.tanh(s)
        elif act is not None:
            raise AssertionError(act)
        out.append(s)
    return out


def _pe64(ens, x_row, e):
    """pe.py:789-838 with scale_output=True for member e and one input row, in float64."""
    h = [float(v) for v in x_row]
    if ens.mu_in is not None:                     # pens/utils.py:156
        h = [(h[i] - float(ens.mu_in[0, i])) / max(math.sqrt(float(ens.var_in[0, i])), 1e-2) for i in range(len(h))]
    for W, b, a in zip(ens.W, ens.b, ens.acts):
        h = _fc64(h, W[e], b[e].reshape(-1), a)
    D = len(h) // 2 if ens.probabilistic else len(h)
    mean = h[:D]
    if ens.mu_out is not None:                    # pens/utils.py:167
        mean = [max(math.sqrt(float(ens.var_out[0, d])), 1e-2) * mean[d] + float(ens.mu_out[0, d]) for d in range(D)]
    if not ens.probabilistic:
        return mean, None
    logvar = h[D:]
    if ens.mu_out is not None:                    # pens/utils.py:187
        logvar = [2.0 * math.log(max(math.sqrt(float(ens.var_out[0, d])), 1e-2)) + logvar[d] for d in range(D)]
    return mean, [math.exp(v) for v in logvar]


def test_pe_forward_matches_independent_float64_restatement():
    for seed, (O, A, hidden, task) in enumerate([(17, 6, (24, 24), "HalfCheetahSafe-v2"), (29, 8, (16, 16), "AntSafe-v2")]):
        dyn, actor, v, vc = orc.make_problem(700 + seed, O, A, hidden=hidden, task=task)
        obs, act = orc.make_states(710 + seed, 6, O, A, dyn)
        x = np.concatenate([obs, act], -1)
        mean, var = orc.pe_forward(dyn, x)
        sig = np.maximum(np.sqrt(dyn.var_out[0]), 1e-2)
        for e in range(dyn.num_nets):
            for r in range(x.shape[0]):
                m64, v64 = _pe64(dyn, x[r], e)
                # float32 evaluation of a ~40-term dot product chain: a few 1e-6 relative to the output scale
                assert np.allclose(mean[e, r], m64, rtol=2e-5, atol=2e-5 * float(sig.max())), (seed, e, r)
                assert np.allclose(var[e, r], v64, rtol=2e-4), (seed, e, r)
        # value ensembles: non-probabilistic, PE.predict = mean over members (pe.py:343)
        pv = orc.pe_predict(v, obs)
        for r in range(obs.shape[0]):
            ms = [_pe64(v, obs[r], e)[0][0] for e in range(v.num_nets)]
            assert np.allclose(pv[r, 0], sum(ms) / len(ms), rtol=2e-5, atol=2e-5)
        # actor
        mu = orc.actor_mu(actor, obs)
        for r in range(obs.shape[0]):
            h = [float(t) for t in obs[r]]
            for i, (W, b) in enumerate(zip(actor.W, actor.b)):
                h = _fc64(h, W, b, "tanh" if i < len(actor.W) - 1 else None)
            assert np.allclose(mu[r], h, rtol=2e-5, atol=2e-5), (seed, r)


def test_average_dkl_matches_independent_float64_restatement():
    """pens/utils.py:15-57 written as the textbook double sum over ordered member pairs, float64."""
    rng = np.random.default_rng(3)
    E, N, D = 5, 4, 3
    mean = rng.normal(size=(E, N, D)).astype(np.float32)
    std = np.exp(rng.normal(scale=0.5, size=(E, N, D))).astype(np.float32)
    got = orc.average_dkl(mean, std)
    for n in range(N):
        for d in range(D):
            tot = 0.0
            for i in range(E):
                for j in range(E):
                    m0, s0, m1, s1 = (float(mean[i, n, d]), float(std[i, n, d]), float(mean[j, n, d]), float(std[j, n, d]))
                    kl = 0.5 * (((m1 - m0) ** 2 + s0 ** 2) / (s1 ** 2 + 1e-10) - 1.0) + math.log(s1) - math.log(s0)
                    tot += min(max(kl, 0.0), 1e10)
            assert np.isclose(got[n, d], tot / (E * (E - 1) + 1e-10), rtol=2e-5, atol=1e-6), (n, d)
