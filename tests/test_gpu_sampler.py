"""Drop-in parity of ModelSampler / ModelBuffer (fused and step-wise modes) against the oracle's
reset -> sample x n -> finish_all_paths -> get cycle (algorithms/cmbpo.py:251-269), including the
max_samples cap and the alive-ratio stop.  Row order / count / masks exact; floats rtol 2e-3."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from helpers import TASKS, GAE, load_problem, ShapeEnv, calibrated_dkl_lim

pytestmark = pytest.mark.gpu


def _drive(smp, pool, max_samples, stop_ratio):
    while True:
        _, _, _, info = smp.sample(max_samples)
        if max_samples and smp._total_samples >= 0.99 * max_samples:
            break
        if info["alive_ratio"] <= stop_ratio:
            break
    diag = smp.finish_all_paths()
    out, bdiag = pool.get()
    return out, bdiag, diag


def _oracle_cycle(dyn, actor, v, vc, task, obs, noise, T, mode, lim, max_samples, stop_ratio):
    O, A = obs.shape[1], actor.W[-1].shape[1]
    policy = orc.OraclePolicy(actor, v, vc, noise.eps_fn)
    env = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), noise.idx_fn)
    pool = orc.OracleModelBuffer(obs.shape[0], O, A, T)
    pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    smp = orc.OracleModelSampler(T, obs.shape[0], mode)
    smp.initialize(env, policy, pool)
    smp.set_rollout_dkl(lim)
    smp.reset(obs)
    while True:
        _, _, _, info = smp.sample(max_samples)
        if max_samples and smp.total_samples >= 0.99 * max_samples:
            break
        if info["alive_ratio"] <= stop_ratio:
            break
    diag = smp.finish_all_paths()
    out, bdiag = pool.get()
    return out, bdiag, diag


@pytest.mark.parametrize("key,mode,max_samples", [
    ("hcs", "uncertainty", None), ("hcs", False, 1500), ("ant", "uncertainty", 900), ("ant", False, None)])
def test_fused_sampler_cycle(engine, key, mode, max_samples):
    import cmbpo_b200 as cb
    task, O, A = TASKS[key]
    B, T = 256, 10
    dyn, actor, v, vc = orc.make_problem(61, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(62, B, O, A, dyn)
    noise = orc.TableNoise(63, T, B, A, len(dyn.elite_inds))
    lim = calibrated_dkl_lim(dyn, task, obs, act) if mode else None
    want, wb, wd = _oracle_cycle(dyn, actor, v, vc, task, obs, noise, T, mode, lim, max_samples, 0.1)

    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    pool = cb.ModelBuffer(B, O, A, T, engine=engine)
    pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    smp = cb.ModelSampler(T, B, mode, logger=object())
    smp.initialize(env, policy, pool)
    smp.set_rollout_dkl(lim)
    smp.injected = dict(act_eps=noise.act_eps, elite_pos=noise.elite_pos)
    assert smp.fused
    smp.reset(obs)
    got, gb, gd = _drive(smp, pool, max_samples, 0.1)

    assert len(got) == len(want) == 12
    n_w, n_g = len(want[0]), len(got[0])
    # discrete outcomes can flip only within fp32 noise of a threshold; counts must be near-equal
    assert abs(n_w - n_g) <= max(2, 0.02 * n_w), (n_w, n_g)
    if n_w == n_g:
        for i, (g, w) in enumerate(zip(got, want)):
            assert g.shape == w.shape and g.dtype == w.dtype, i
            if i == 9:      # cost: discrete
                assert (g != w).mean() < 0.01
            else:
                assert np.allclose(g, w, rtol=5e-3, atol=5e-3), i
        assert gb["poolm_batch_size"] == wb["poolm_batch_size"]
        assert np.isclose(gb["poolm_ret_mean"], wb["poolm_ret_mean"], rtol=2e-3, atol=1e-3)
        for k in ("msampler/samples_added", "msampler/rollout_H_max"):
            assert gd[k] == wd[k], k
        for k in ("msampler/v_mean", "msampler/cv_mean", "msampler/ens_DKL", "msampler/rew_rate",
                  "msampler/dyn_var_perstep", "msampler/max_dkl", "msampler/max_path_return",
                  "msampler/cost_rate"):
            assert np.isclose(gd[k], wd[k], rtol=5e-3, atol=1e-4), (k, gd[k], wd[k])


def test_stepwise_sampler_matches_fused(engine):
    """The step-wise mode (external policy objects) and the fused mode give the same get() list."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["ant"]
    B, T = 128, 8
    dyn, actor, v, vc = orc.make_problem(71, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(72, B, O, A, dyn)
    noise = orc.TableNoise(73, T, B, A, len(dyn.elite_inds))
    lim = calibrated_dkl_lim(dyn, task, obs, act)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    results = []
    for fused in (True, False):
        env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
        pool = cb.ModelBuffer(B, O, A, T, engine=engine)
        pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
        smp = cb.ModelSampler(T, B, "uncertainty", logger=object())
        if fused:
            smp.initialize(env, policy, pool)
            smp.injected = dict(act_eps=noise.act_eps, elite_pos=noise.elite_pos)
        else:
            class HostPolicy:      # opaque policy object: forces the step-wise path
                agent = policy.agent

                def reset(self):
                    pass

                def get_action_outs(self, o, _s=smp, _p=pool):
                    ids = np.flatnonzero(_p.alive_paths)
                    return policy.get_action_outs(o, eps=noise.act_eps[_s._n_episodes - 1, ids])

                get_v = staticmethod(policy.get_v)
                get_vc = staticmethod(policy.get_vc)

            class HostEnv:
                def step(self, o, a, _s=smp, _p=pool):
                    ids = np.flatnonzero(_p.alive_paths)
                    return env.step(o, a, elite_pos=noise.elite_pos[_s._n_episodes - 1, ids])

                def close(self):
                    pass
            smp.initialize(HostEnv(), HostPolicy(), pool)
            assert not smp.fused
        smp.set_rollout_dkl(lim)
        smp.reset(obs)
        results.append(_drive(smp, pool, 700, 0.1))
    (g1, b1, d1), (g2, b2, d2) = results
    assert len(g1[0]) == len(g2[0])
    for i, (a, b) in enumerate(zip(g1, g2)):
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=1e-4, atol=1e-5), i
    assert d1["msampler/samples_added"] == d2["msampler/samples_added"]


def test_get_async_pipelines_batches_and_equals_get(engine):
    """ModelBuffer.get_async(): the sample list of batch i is copied while batch i+1 rolls out; its
    result() is bit-identical to get() of the same batch, with two handles outstanding."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["hcs"]
    B, T = 512, 12
    dyn, actor, v, vc = orc.make_problem(91, O, A, hidden=(64, 64), task=task)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    pool = cb.ModelBuffer(B, O, A, T, engine=engine)
    pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    batches = [orc.make_states(92 + i, B, O, A, dyn)[0] for i in range(3)]

    def run(obs, seed, asynchronous):
        smp = cb.ModelSampler(T, B, False, logger=object(), seed=seed)
        smp.initialize(env, policy, pool)
        smp.reset(obs)
        while smp.sample(None)[3]["alive_ratio"] > 0.1:
            pass
        smp.finish_all_paths()
        return pool.get_async() if asynchronous else pool.get()

    sync = [run(o, 5 + i, False) for i, o in enumerate(batches)]
    sync = [([a.copy() for a in out], diag) for out, diag in sync]
    pending, got = None, []
    for i, o in enumerate(batches):
        nxt = run(o, 5 + i, True)
        if pending is not None:
            out, diag = pending.result()
            got.append(([a.copy() for a in out], diag))      # consume before the buffers are recycled
        pending = nxt
    out, diag = pending.result()
    got.append(([a.copy() for a in out], diag))
    for (wo, wd), (go, gd) in zip(sync, got):
        assert wd == gd and len(wo) == len(go) == 12
        for a, b in zip(wo, go):
            assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize("key,depth", [("hcs", 1), ("hcs", 3), ("ant", 4)])
def test_compute_dynamics_dkl_device_mode(engine, key, depth):
    """ModelSampler.compute_dynamics_dkl (samplers/model_sampler.py:151-167, called at algorithms/cmbpo.py:198-200)
    as one depth-step device rollout that stores nothing, against the oracle's host loop with the same noise."""
    import cmbpo_b200 as cb
    task, O, A = TASKS[key]
    B = 700
    dyn, actor, v, vc = orc.make_problem(81, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(82, B, O, A, dyn)
    noise = orc.TableNoise(83, depth, B, A, len(dyn.elite_inds))
    policy_o = orc.OraclePolicy(actor, v, vc, noise.eps_fn)
    env_o = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), noise.idx_fn)
    pool_o = orc.OracleModelBuffer(B, O, A, 8)
    pool_o.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    so = orc.OracleModelSampler(8, B, False)
    so.initialize(env_o, policy_o, pool_o)
    so.reset(obs)
    want = so.compute_dynamics_dkl(obs, depth)

    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    pool = cb.ModelBuffer(B, O, A, 8, engine=engine)
    pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    smp = cb.ModelSampler(8, B, False, logger=object())
    smp.initialize(env, policy, pool)
    smp.injected = dict(act_eps=noise.act_eps, elite_pos=noise.elite_pos)
    assert smp.fused
    got = smp.compute_dynamics_dkl(obs, depth)
    assert want > 0
    assert np.isclose(got, want, rtol=2e-3), (got, want)
    # rows fed per step: exact unless a termination flipped within float noise of its threshold
    assert abs(smp._total_samples - so.total_samples) <= max(1, 0.002 * so.total_samples)
    assert np.isclose(smp.dyn_dkl, so.total_dkl / (so.total_samples + 1e-8), rtol=2e-3)
    # Philox noise (no injection) runs too and gives a value of the same size
    smp2 = cb.ModelSampler(8, B, False, logger=object(), seed=3)
    smp2.initialize(env, policy, pool)
    got2 = smp2.compute_dynamics_dkl(engine.to_device(obs), depth)        # device tensor in
    assert 0.5 * want < got2 < 2.0 * want


@pytest.mark.parametrize("H", [1, 2, 3])
def test_short_horizons_store_max_1_Hminus1_steps(engine, H):
    """set_max_path_length(H) with H <= 3 (schedule mode with min_length = 1): the reference stores
    max(1, H - 1) steps per path (model_sampler.py:352).  ADVICE round 1: H = 1 used to run the full rollout."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["hcs"]
    B, T = 200, 10
    dyn, actor, v, vc = orc.make_problem(91, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(92, B, O, A, dyn)
    noise = orc.TableNoise(93, T, B, A, len(dyn.elite_inds))
    policy_o = orc.OraclePolicy(actor, v, vc, noise.eps_fn)
    env_o = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), noise.idx_fn)
    pool_o = orc.OracleModelBuffer(B, O, A, T)
    pool_o.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    so = orc.OracleModelSampler(T, B, False)
    so.initialize(env_o, policy_o, pool_o)
    so.set_max_path_length(H)
    so.reset(obs)
    while pool_o.alive_paths.any():
        so.sample()
    wd = so.finish_all_paths()
    want, wb = pool_o.get()

    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    pool = cb.ModelBuffer(B, O, A, T, engine=engine)
    pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
    smp = cb.ModelSampler(T, B, False, logger=object())
    smp.initialize(env, policy, pool)
    smp.set_max_path_length(H)
    smp.injected = dict(act_eps=noise.act_eps, elite_pos=noise.elite_pos)
    smp.reset(obs)
    n_calls = 0
    while True:
        _, _, _, info = smp.sample()
        n_calls += 1
        if info["alive_ratio"] <= 0:
            break
    gd = smp.finish_all_paths()
    got, gb = pool.get()
    assert n_calls == max(1, H - 1)
    assert len(got[0]) == len(want[0]) == B * max(1, H - 1)
    assert gd["msampler/samples_added"] == wd["msampler/samples_added"] == B * max(1, H - 1)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape and g.dtype == w.dtype, i
        if i != 9:
            assert np.allclose(g, w, rtol=5e-3, atol=5e-3), i


def test_return_arrays_and_get_device(engine):
    """ModelSampler.return_arrays (sample() returns next_obs / reward / terminal of the stored rows like the
    reference) and ModelBuffer.get_device() (the 12 arrays of get() as device tensors, same values)."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["ant"]
    B, T = 160, 7
    dyn, actor, v, vc = orc.make_problem(95, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(96, B, O, A, dyn)
    noise = orc.TableNoise(97, T, B, A, len(dyn.elite_inds))
    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)

    def cycle(return_arrays, device):
        pool = cb.ModelBuffer(B, O, A, T, engine=engine)
        pool.initialize({"mu": (A,), "log_std": (A,)}, **GAE)
        smp = cb.ModelSampler(T, B, False, logger=object())
        smp.initialize(env, policy, pool)
        smp.injected = dict(act_eps=noise.act_eps, elite_pos=noise.elite_pos)
        smp.return_arrays = return_arrays
        smp.reset(obs)
        steps = []
        while True:
            nxt, rew, term, info = smp.sample()
            steps.append((nxt, rew, term))
            if info["alive_ratio"] <= 0:
                break
        smp.finish_all_paths()
        if device:
            out, diag = pool.get_device()
            assert all(hasattr(x, "is_cuda") and x.is_cuda for x in out)
            # consumer hand-off through the DLPack protocol: a consumer that only speaks DLPack
            # sees the same device memory (zero copy), including the stride-0 log_std view
            import torch
            for x in out:
                y = torch.from_dlpack(x.__dlpack__())
                assert y.data_ptr() == x.data_ptr() and y.shape == x.shape and torch.equal(y, x)
            out = [x.cpu().numpy() for x in out]
            pool.reset()
        else:
            out, diag = pool.get()
        return steps, out, diag

    steps, out_dev, d1 = cycle(True, True)
    steps0, out_host, d2 = cycle(False, False)
    assert all(s[0] is None for s in steps0)
    assert len(out_dev) == len(out_host) == 12
    for i, (a, b) in enumerate(zip(out_dev, out_host)):
        assert a.shape == b.shape and a.dtype == b.dtype, i
        assert np.array_equal(a, b), i
    assert d1["poolm_batch_size"] == d2["poolm_batch_size"]
    # the per-step arrays are the stored rows of that step, in path order
    bufs_len = np.zeros(B, np.int64)
    total = 0
    for t, (nxt, rew, term) in enumerate(steps):
        assert nxt.shape[1] == O and nxt.shape[0] == rew.shape[0] == term.shape[0]
        assert term.dtype == np.bool_
        total += nxt.shape[0]
    assert total == out_host[0].shape[0]
    # path-major get() order: the rows of step 0 are the first row of every path
    first_rows = np.concatenate([[0], np.cumsum(np.bincount(np.repeat(np.arange(B), 1), minlength=B))[:-1]])
    assert steps[0][0].shape[0] == B and len(first_rows) == B
