"""Synthetic CMBPO rollout problems for bench.py and the debug tools: ensemble / actor / value
weights and start states of the shapes SURVEY.md section 8d names (there is no network for real
checkpoints).  Plain numpy, no dependency on the test oracle; tests/test_workload_cpu.py checks that
it generates exactly what the oracle's own generator does, so both sides see the same problem.

Initialisers follow the reference: truncated-normal weights with std 1/(2 sqrt(in)) and zero biases
(models/pens/fc.py:135-139), Glorot-uniform actor layers (tf.layers.dense), log_std = -0.5
(network/ac_network.py:104); scalers are data-like (sigma in [0.05, 3], two nearly constant
dimensions that exercise the 1e-2 clamp of pens/utils.py:156).
"""
from types import SimpleNamespace

import numpy as np

F32 = np.float32


def _trunc_normal(rng, shape, std):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return (x * std).astype(F32)


def make_ensemble(rng, in_dim, out_dim, hidden, num_nets, num_elites, probabilistic, act="swish",
                  scalers=None, gain=1.0, bias_std=0.0):
    last = 2 * out_dim if probabilistic else out_dim
    dims = [in_dim] + list(hidden) + [last]
    W, b, acts = [], [], []
    for i in range(len(dims) - 1):
        W.append(_trunc_normal(rng, (num_nets, dims[i], dims[i + 1]), gain / (2 * np.sqrt(dims[i]))))
        b.append((bias_std * rng.standard_normal((num_nets, 1, dims[i + 1]))).astype(F32))
        acts.append(act if i < len(dims) - 2 else None)
    if scalers is None:
        scalers = (np.zeros((1, in_dim)), np.ones((1, in_dim)), np.zeros((1, out_dim)), np.ones((1, out_dim)))
    mu_in, var_in, mu_out, var_out = (np.asarray(a, F32).reshape(1, -1) for a in scalers)
    elites = [int(i) for i in rng.permutation(num_nets)[:num_elites]]
    return SimpleNamespace(W=W, b=b, acts=acts, probabilistic=probabilistic, mu_in=mu_in, var_in=var_in,
                           mu_out=mu_out, var_out=var_out, elite_inds=elites)


def make_scalers(rng, task, obs_dim, act_dim):
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
    return (mu_in, var_in, mu_out, var_out), (mu, s ** 2)


def make_actor(rng, obs_dim, act_dim, hidden=(128, 128), gain=1.0):
    dims = [obs_dim] + list(hidden) + [act_dim]
    W, b = [], []
    for i in range(len(dims) - 1):
        lim = gain * np.sqrt(6.0 / (dims[i] + dims[i + 1]))
        W.append(rng.uniform(-lim, lim, (dims[i], dims[i + 1])).astype(F32))
        b.append(np.zeros(dims[i + 1], F32))
    return SimpleNamespace(W=W, b=b, log_std=np.full(act_dim, -0.5, F32))


def make_problem(seed, obs_dim, act_dim, hidden=(512, 512), num_nets=7, num_elites=5, vf_nets=3,
                 vf_hidden=(128, 128), a_hidden=(128, 128), task="HalfCheetahSafe-v2", gain=2.0):
    """(dynamics ensemble, actor, V ensemble, VC ensemble) as namespaces of numpy arrays."""
    rng = np.random.default_rng(seed)
    dyn_sc, (mu_o, var_o) = make_scalers(rng, task, obs_dim, act_dim)
    v_sc = (mu_o, var_o, [3.0], [25.0])
    vc_sc = (mu_o, var_o, [1.0], [4.0])
    dyn = make_ensemble(rng, obs_dim + act_dim, obs_dim + 1, hidden, num_nets, num_elites, True,
                        scalers=dyn_sc, gain=gain)
    actor = make_actor(rng, obs_dim, act_dim, a_hidden)
    sig = np.maximum(np.sqrt(var_o), 0.1)         # the actor sees raw observations: fold (obs-mu)/sigma in
    w0 = actor.W[0].astype(np.float64)
    actor.W[0] = (w0 / sig[:, None]).astype(F32)
    actor.b[0] = (-(mu_o / sig) @ w0).astype(F32)
    v = make_ensemble(rng, obs_dim, 1, vf_hidden, vf_nets, 2, False, scalers=v_sc, gain=gain)
    vc = make_ensemble(rng, obs_dim, 1, vf_hidden, vf_nets, 2, False, scalers=vc_sc, gain=gain)
    return dyn, actor, v, vc


def make_states(seed, n, obs_dim, act_dim, dyn=None):
    """obs ~ N(mu_in, sigma_in) with sigma = max(sqrt(var), 1e-2); act ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((n, obs_dim)).astype(F32)
    if dyn is not None and dyn.mu_in is not None:
        sigma = np.maximum(np.sqrt(dyn.var_in), F32(1e-2)).astype(F32)
        obs = obs * sigma[:, :obs_dim] + dyn.mu_in[:, :obs_dim]
    act = rng.uniform(-1, 1, (n, act_dim)).astype(F32)
    return obs.astype(F32), act
