"""GPU parity of the ensemble forward, FakeEnv.step, policy and the H-step rollout (fp32 CUDA-core
variant = logic parity) against the oracle on identical inputs and identical injected noise.

Tolerances (stated): single-step mean/var rtol 1e-4 + atol 1e-5*sigma_out (fp32 GEMM reordering);
masks / indices / lengths bit-exact; H-step float fields rtol 2e-3 (error grows along the path)."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from helpers import TASKS, GAE, load_problem, ShapeEnv, calibrated_dkl_lim

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,hidden", [("hcs", (64, 64)), ("ant", (200, 200, 200, 200)), ("hum", (512, 512))])
def test_predict_ensemble_fp32(engine, key, hidden):
    task, O, A = TASKS[key]
    dyn, actor, v, vc = orc.make_problem(11, O, A, hidden=hidden, task=task)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    obs, act = orc.make_states(12, 1000, O, A, dyn)
    x = np.concatenate([obs, act], -1)
    mean, var = model.predict_ensemble(x)
    wm, wv = orc.pe_forward(dyn, x)
    sig = np.maximum(np.sqrt(dyn.var_out), 1e-2)
    assert mean.shape == wm.shape == (7, 1000, O + 1)
    assert np.all(np.abs(mean - wm) <= 1e-4 * np.abs(wm) + 1e-5 * sig)
    assert np.allclose(var, wv, rtol=2e-4, atol=0)
    # 3-D inputs: member i on slice i (fc.py:89-90)
    x3 = np.stack([x[i * 100:(i + 1) * 100] for i in range(7)])
    m3, v3 = model.predict_ensemble(x3)
    w3m, w3v = orc.pe_forward(dyn, x3)
    assert np.all(np.abs(m3 - w3m) <= 1e-4 * np.abs(w3m) + 1e-5 * sig)
    # PE.predict (mean over all members) for the value heads
    pv = policy.get_v(obs)
    assert pv.shape == (1000,)
    assert np.allclose(pv, np.squeeze(orc.pe_predict(v, obs), -1), rtol=1e-4, atol=1e-5)
    pm, pvv = model.predict(x)
    om, ov = orc.pe_predict(dyn, x)
    assert np.all(np.abs(pm - om) <= 1e-4 * np.abs(om) + 1e-5 * sig)
    assert np.allclose(pvv, ov, rtol=3e-4, atol=1e-9)


@pytest.mark.parametrize("key", ["hcs", "ant", "hum"])
def test_fakeenv_step(engine, key):
    import cmbpo_b200 as cb
    task, O, A = TASKS[key]
    dyn, actor, v, vc = orc.make_problem(21, O, A, hidden=(64, 64), task=task)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    obs, act = orc.make_states(22, 2000, O, A, dyn)
    pos = np.random.default_rng(1).integers(0, 5, 2000)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    nxt, r, term, info = env.step(obs, act, elite_pos=pos)
    oenv = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), lambda e, n: pos)
    wn, wr, wt, winfo = oenv.step(obs, act)
    assert nxt.shape == wn.shape and r.shape == wr.shape == (2000, 1) and term.shape == wt.shape
    assert term.dtype == np.bool_ and info["cost"].dtype == winfo["cost"].dtype
    assert np.allclose(nxt, wn, rtol=1e-4, atol=1e-5)
    assert np.allclose(r, wr, rtol=1e-4, atol=1e-5)
    assert np.allclose(info["ensemble_dkl_path"], winfo["ensemble_dkl_path"], rtol=2e-3, atol=1e-6)
    assert np.allclose(info["ensemble_ep_var"], winfo["ensemble_ep_var"], rtol=2e-3, atol=1e-9)
    assert np.isclose(info["ensemble_dkl_mean"], winfo["ensemble_dkl_mean"], rtol=1e-3)
    # statics are bit-exact functions of next_obs: evaluate the oracle statics on the kernel's own
    # next_obs, and require equality with the oracle run wherever no threshold is within 1e-4
    tid, cid = orc.task_fns(task)
    assert np.array_equal(term, orc.apply_term(tid, nxt))
    assert np.array_equal(info["cost"], orc.apply_cost(cid, nxt, term))
    assert (term != wt).mean() < 2e-3 and (info["cost"] != winfo["cost"]).mean() < 2e-3
    # single (1-D) observation path, fake_env.py:71-77, 157-161
    n1, r1, t1, i1 = env.step(obs[0], act[0], elite_pos=pos[:1])
    assert n1.shape == (O,) and r1.shape == (1,) and t1.shape == (1,)
    assert np.allclose(n1, wn[0], rtol=1e-4, atol=1e-5)


def test_statics_truth_table(engine):
    """AntSafe precedence quirk (statics.py:24-27), NaN handling, HCS threshold: bit-exact."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["ant"]
    dyn, actor, v, vc = orc.make_problem(31, O, A, hidden=(64, 64), task=task)
    # zero the model so next_obs == obs exactly: weights 0, bias 0, mu_out 0
    for W in dyn.W:
        W[:] = 0
    dyn.mu_out[:] = 0
    model, _ = load_problem(engine, dyn, actor, v, vc)
    rows = []
    for z in (0.1, 0.2, 0.6, 1.0, 1.1, np.nan, np.inf):
        for q in ((0.0, 0.0), (0.93, 0.0), (0.65, 0.66), (np.nan, 0.0)):
            for y in (0.0, 3.2, 3.3, -4.0, np.nan):
                o = np.zeros(O, np.float32)
                o[0], o[2], o[3], o[-1] = z, q[0], q[1], y
                rows.append(o)
    o = np.zeros(O, np.float32); o[0] = 0.6; o[7] = np.nan; rows.append(o)   # NaN elsewhere
    obs = np.array(rows, np.float32)
    act = np.zeros((len(obs), A), np.float32)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    nxt, r, term, info = env.step(obs, act, elite_pos=np.zeros(len(obs), int))
    # a NaN/inf anywhere in the input row poisons the whole prediction (0 * nan), in the reference too
    oenv = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), lambda e, n: np.zeros(n, int))
    wn, _, wt, winfo = oenv.step(obs, act)
    clean = np.isfinite(obs).all(1)
    assert np.array_equal(nxt[clean], obs[clean]) and np.array_equal(np.isnan(nxt), np.isnan(wn))
    assert np.array_equal(term, wt) and np.array_equal(info["cost"], winfo["cost"])
    assert np.array_equal(term, orc.antsafe_term(nxt))
    assert np.array_equal(info["cost"], orc.antsafe_cost(nxt))
    assert term[clean].any() and not term[clean].all()
    # HCS cost threshold |10 x| < 2
    task, O, A = TASKS["hcs"]
    dyn, actor, v, vc = orc.make_problem(32, O, A, hidden=(64, 64), task=task)
    for W in dyn.W:
        W[:] = 0
    dyn.mu_out[:] = 0
    model, _ = load_problem(engine, dyn, actor, v, vc)
    xs = np.array([0.0, 0.19999999, 0.2, np.float32(0.2), -0.2, 0.20000002, np.nan, -0.1], np.float32)
    obs = np.zeros((len(xs), O), np.float32); obs[:, -1] = xs
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    nxt, r, term, info = env.step(obs, np.zeros((len(xs), A), np.float32), elite_pos=np.zeros(len(xs), int))
    assert np.array_equal(info["cost"], orc.hcs_cost(obs)) and not term.any()


def _run_device_rollout(engine, dyn, actor, v, vc, task, obs, noise, T, mode, lim, precision=None):
    import cmbpo_b200 as cb
    O, A = obs.shape[1], actor.W[-1].shape[1]
    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    bufs = cb.RolloutBuffers(engine, obs.shape[0], T, O, A)
    bufs.set_inputs(obs, noise.act_eps, noise.elite_pos)
    bufs.run(env.env_cfg(True), uncertainty_mode=(mode == "uncertainty"), dkl_lim=lim or 0.0,
             precision=precision)
    bufs.gae(GAE["gamma"], GAE["lam"], GAE["cost_gamma"], GAE["cost_lam"])
    return bufs


@pytest.mark.parametrize("key,mode", [("hcs", False), ("hcs", "uncertainty"), ("ant", False),
                                      ("ant", "uncertainty"), ("hum", False)])
def test_rollout_fp32_vs_oracle(engine, key, mode):
    task, O, A = TASKS[key]
    B, T = 300, 12
    dyn, actor, v, vc = orc.make_problem(41, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(42, B, O, A, dyn)
    noise = orc.TableNoise(43, T, B, A, len(dyn.elite_inds))
    lim = calibrated_dkl_lim(dyn, task, obs, act) if mode else None
    out, bdiag, diag, snap = orc.run_rollout(dyn, actor, v, vc, task, obs, noise, T, mode, lim,
                                             gamma=GAE["gamma"], lam=GAE["lam"],
                                             cgamma=GAE["cost_gamma"], clam=GAE["cost_lam"])
    bufs = _run_device_rollout(engine, dyn, actor, v, vc, task, obs, noise, T, mode, lim)
    pop = bufs.populated_mask()
    want_len = snap["populated"].sum(1)
    got_len = bufs.length.cpu().numpy()
    same = got_len == want_len
    # discrete outcomes may flip only where a threshold is within float32 noise of the value
    assert same.mean() > 0.97, (same.mean(), np.flatnonzero(~same)[:10])
    ok = same
    for name in ("obs", "act", "nextobs", "rew", "val", "cost", "cval", "logp", "mu", "dyn_error"):
        g, w = bufs.host(name), snap[name]
        m = pop & snap["populated"] & ok[:, None]
        if name == "cost":
            assert (g[m] != w[m]).mean() < 5e-3
            continue
        assert np.allclose(g[m], w[m], rtol=2e-3, atol=2e-4), name
    m = pop & snap["populated"] & ok[:, None]
    assert (bufs.host("term")[m] != snap["term"][m]).mean() < 5e-3
    for name in ("adv", "ret", "cadv", "cret"):
        g, w = bufs.host(name), snap[name]
        assert np.allclose(g[m], w[m], rtol=5e-3, atol=2e-3), name
    # unpopulated cells are zero (modelbuffer.py:53-98)
    assert not bufs.host("obs")[~pop].any() and not bufs.host("adv")[~pop].any()


def test_gae_on_device_rollout_is_bit_exact(engine):
    """GAE recomputed by the oracle from the DEVICE's own rew/val/cost/cval equals the device GAE
    bit-for-bit (isolates the scan from the GEMM precision)."""
    task, O, A = TASKS["ant"]
    B, T = 257, 35
    dyn, actor, v, vc = orc.make_problem(51, O, A, hidden=(64, 64), task=task)
    obs, act = orc.make_states(52, B, O, A, dyn)
    noise = orc.TableNoise(53, T, B, A, len(dyn.elite_inds))
    lim = calibrated_dkl_lim(dyn, task, obs, act, factor=12.0)
    bufs = _run_device_rollout(engine, dyn, actor, v, vc, task, obs, noise, T, "uncertainty", lim)
    ln = bufs.length.cpu().numpy()
    lv, lc = bufs.last_val.cpu().numpy(), bufs.last_cval.cpu().numpy()
    rew, val, cost, cval = (bufs.host(k) for k in ("rew", "val", "cost", "cval"))
    reasons = bufs.end_reason.cpu().numpy()
    assert set(np.unique(reasons)) <= {1, 2, 3}
    assert np.all(lv[reasons == 3] == 0)                      # env terminal: no value bootstrap
    for L in np.unique(ln):
        if L == 0:
            continue
        m = ln == L
        want = orc.gae_path(rew[m, :L], val[m, :L], cost[m, :L], cval[m, :L], lv[m], lc[m],
                            GAE["gamma"], GAE["lam"], GAE["cost_gamma"], GAE["cost_lam"])
        for name, w in zip(("adv", "ret", "cadv", "cret"), want):
            assert np.array_equal(bufs.host(name)[m, :L], w), name
