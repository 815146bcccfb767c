"""GPU: the CUDA path against the committed golden fixtures produced by the reference's unmodified
host code (tests/golden, oracle/gen_golden.py).  fp32 CUDA-core variant: discrete outputs exact
where no threshold is within float noise, floats within rtol 1e-4 (single step) / 2e-3 (rollouts);
GAE scans of golden inputs bit-exact."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from golden_io import load, problem
from helpers import load_problem, ShapeEnv

pytestmark = pytest.mark.gpu
GAE = dict(gamma=0.99, lam=0.95, cost_gamma=0.97, cost_lam=0.5)


@pytest.mark.parametrize("key", ["hcs", "ant", "hum"])
def test_fakeenv_step_vs_golden(engine, key):
    import cmbpo_b200 as cb
    z = load("fakeenv_step_%s_tfrestated.npz" % key)
    dyn, actor, v, vc = problem(z)
    model, _ = load_problem(engine, dyn, actor, v, vc)
    obs, act = z["obs"], z["act"]
    env = cb.FakeEnv(ShapeEnv(obs.shape[1], act.shape[1]), str(z["task"]), model, True, True, False)
    nxt, r, term, info = env.step(obs, act, elite_pos=z["elite_pos"])
    assert np.allclose(nxt, z["next_obs"], rtol=1e-4, atol=1e-5)
    assert np.allclose(r, z["rew"], rtol=1e-4, atol=1e-5)
    assert np.allclose(info["ensemble_dkl_path"], z["dkl_path"], rtol=2e-3, atol=1e-6)
    assert np.allclose(info["ensemble_ep_var"], z["ep_var"], rtol=2e-3, atol=1e-9)
    assert term.dtype == z["term"].dtype and info["cost"].dtype == z["cost"].dtype
    assert (term != z["term"]).mean() < 5e-3 and (info["cost"] != z["cost"]).mean() < 5e-3


def test_cpobuffer_vs_golden(engine):
    """CPOBuffer.store / finish_path / get against the reference run (cpobuffer.py:160-290)."""
    import cmbpo_b200 as cb
    z = load("cpobuffer.npz")
    n, seg = len(z["rew"]), z["seg"]
    for scan in (0, 1):
        buf = cb.CPOBuffer(1000, 5000, ShapeEnv(5, 2).observation_space, ShapeEnv(5, 2).action_space, engine=engine)
        buf.initialize({"mu": (2,), "log_std": (2,)}, **GAE)
        buf.scan_mode = scan
        for s in range(len(seg) - 1):
            for i in range(seg[s], seg[s + 1]):
                buf.store(z["obs"][i], z["act"][i], z["obs"][i], z["rew"][i], z["val"][i], z["cost"][i],
                          z["cval"][i], z["logp"][i], {"mu": z["mu"][i], "log_std": z["log_std"][i]}, False, 0)
            buf.finish_path(z["last_val"][s:s + 1], z["last_cval"][s:s + 1])
        for k in ("adv", "ret", "cadv", "cret"):
            got = getattr(buf, k + "_buf")[:n]
            if scan == 0:
                assert np.array_equal(got, z["pre_" + k]), k          # strict scan: bit-exact vs scipy lfilter
            else:
                assert np.all(np.abs(got - z["pre_" + k]) <= 2 * np.spacing(np.abs(z["pre_" + k]))), k
        out, diag = buf.get()
        assert len(out) == 12
        for i in (0, 1, 4, 5, 6, 7, 8, 9, 10, 11):
            assert np.array_equal(out[i], z["out%d" % i]), i
        if scan == 0:
            assert np.allclose(out[2], z["out2"], rtol=1e-5, atol=1e-6)   # normalised adv: float32 sum order
            assert np.allclose(out[3], z["out3"], rtol=1e-5, atol=1e-6)
            assert np.isclose(diag["poolr_ret_mean"], z["ret_mean"], rtol=1e-5)


def test_discount_cumsum_golden_on_device(engine):
    import torch
    z = load("discount_cumsum.npz")
    x = z["x"]                                             # [40, 34] deltas: feed as rewards with zero values
    B, T = x.shape
    zero = np.zeros_like(x)
    d = lambda a: engine.to_device(np.ascontiguousarray(a))
    adv, ret, cadv, cret = engine.gae_paths(d(x), d(zero), d(x), d(zero), None, d(np.zeros(B, np.float32)),
                                            d(np.zeros(B, np.float32)), 0.99, 0.95, 0.97, 0.5, B, T, T, 1)
    assert np.array_equal(adv.cpu().numpy(), z["out_gae"].astype(np.float32))
    assert np.array_equal(cadv.cpu().numpy(), z["out_cgae"].astype(np.float32))


@pytest.mark.parametrize("key", ["hcs", "ant", "hum"])
def test_rollout_vs_golden(engine, key):
    import cmbpo_b200 as cb
    z = load("rollout_%s_tfrestated.npz" % key)
    dyn, actor, v, vc = problem(z)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    obs = z["start_obs"]
    B, O = obs.shape
    A, T = actor.W[-1].shape[1], int(z["T"])
    mode = False if str(z["mode"]) == "False" else str(z["mode"])
    lim = None if z["dkl_lim"] < 0 else float(z["dkl_lim"])
    ms = None if z["max_samples"] < 0 else int(z["max_samples"])
    env = cb.FakeEnv(ShapeEnv(O, A), str(z["task"]), model, True, True, False)
    pool = cb.ModelBuffer(B, O, A, T, engine=engine)
    pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    smp = cb.ModelSampler(T, B, mode, logger=object())
    smp.initialize(env, policy, pool)
    smp.set_rollout_dkl(lim)
    smp.injected = dict(act_eps=z["act_eps"], elite_pos=z["elite_pos"])
    smp.reset(obs)
    while True:
        _, _, _, info = smp.sample(ms)
        if ms and smp._total_samples >= 0.99 * ms:
            break
        if info["alive_ratio"] <= 0.1:
            break
    smp.finish_all_paths()
    pop = pool.populated_mask
    want_pop = z["snap_populated"]
    same = (pop == want_pop).all(1)
    assert same.mean() >= 0.95                      # threshold flips only within float noise of a limit
    len_g, len_w = pop.sum(1), want_pop.sum(1)
    out, diag = pool.get()
    if same.all():
        assert diag["poolm_batch_size"] == int(z["poolm_batch_size"])
    # never vacuous: the rows of the agreeing paths are compared whatever happened to the others
    # (get() is path-major: path p owns len[p] consecutive rows)
    sel_g, sel_w = np.repeat(same, len_g), np.repeat(same, len_w)
    assert sel_g.sum() == sel_w.sum() > 0
    for i in range(12):
        w = z["out%d" % i]
        assert out[i].dtype == w.dtype and out[i].shape[1:] == w.shape[1:], i
        assert out[i].shape[0] == len_g.sum() and w.shape[0] == len_w.sum(), i
        g_, w_ = out[i][sel_g], w[sel_w]
        if i == 9:
            assert (g_ != w_).mean() < 0.02
        elif i in (2, 3):
            # normalised advantages: the batch mean / std include the flipped paths (mpi_statistics_scalar over
            # the whole batch), so they only agree tightly when every path agrees
            if same.all():
                assert np.allclose(g_, w_, rtol=5e-3, atol=5e-3), i
            else:
                assert np.corrcoef(g_.ravel(), w_.ravel())[0, 1] > 0.999, i
        else:
            assert np.allclose(g_, w_, rtol=5e-3, atol=5e-3), i


def test_philox_shard_invariance(engine):
    """Philox noise is keyed by (seed, GLOBAL path id, step): rolling out B paths in one go or as two
    shards with path_id_base gives bit-identical buffers (what makes G=1 and G=8 agree)."""
    import cmbpo_b200 as cb
    from cmbpo_b200 import _lib as L
    task, O, A = "AntSafe-v2", 29, 8
    B, T = 512, 10
    dyn, actor, v, vc = orc.make_problem(5, O, A, hidden=(64, 64), task=task)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    obs, _ = orc.make_states(6, B, O, A, dyn)
    cfg = L.EnvCfg(L.TERM_ANTSAFE, L.COST_ANTSAFE, 0, 0, 1)        # non-deterministic: exercises state noise too

    def run(lo, hi):
        bufs = cb.RolloutBuffers(engine, hi - lo, T, O, A)
        bufs.set_inputs(obs[lo:hi])
        bufs.run(cfg, seed=99, path_id_base=lo)
        return bufs
    whole, a, b = run(0, B), run(0, 200), run(200, B)
    for name in ("obs", "act", "nextobs", "rew", "cost", "val", "logp", "term", "dkl"):
        w = whole.host(name)
        parts = np.concatenate([a.host(name), b.host(name)], axis=0)
        assert np.array_equal(w, parts, equal_nan=True), name
    assert np.array_equal(whole.length.cpu().numpy(),
                          np.concatenate([a.length.cpu().numpy(), b.length.cpu().numpy()]))
    # the draws are not degenerate: elite members vary across paths, actions differ from the mean
    assert np.abs(whole.host("act") - whole.host("mu"))[whole.populated_mask()].mean() > 0.1
