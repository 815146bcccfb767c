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
