"""tcgen05 path (fp16 / bf16 operands, fp32 accumulate in TMEM) against the oracle and against the
fp32 CUDA-core variant.  Stated tolerance (BASELINE.json north_star): single-step
mean within rtol 1e-3 + atol 1e-3*sigma_out, variance within rtol 1e-3 (fp16) / 1e-2 (bf16)."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from helpers import TASKS, GAE, load_problem, ShapeEnv, calibrated_dkl_lim

pytestmark = pytest.mark.gpu


def _err(got, want, sig):
    d = np.abs(got - want)
    return float(np.max(d / (1e-3 * np.abs(want) + 1e-3 * sig)))


@pytest.mark.parametrize("key", ["hcs", "ant", "hum"])
@pytest.mark.parametrize("prec,vtol", [("fp16", 1e-3), ("bf16", 1e-2)])
def test_predict_ensemble_tc(engine, key, prec, vtol):
    task, O, A = TASKS[key]
    dyn, actor, v, vc = orc.make_problem(81, O, A, hidden=(512, 512), task=task)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    N = 1000 if key != "hcs" else 40000          # > 148 tiles: exercises the persistent tile loop
    obs, act = orc.make_states(82, N, O, A, dyn)
    x = np.concatenate([obs, act], -1)
    mean, var = (t.cpu().numpy() for t in model.predict_ensemble_device(x, precision=prec))
    wm, wv = orc.pe_forward(dyn, x)
    sig = np.maximum(np.sqrt(dyn.var_out), 1e-2)
    e = _err(mean, wm, sig)
    rel_v = float(np.max(np.abs(var - wv) / wv))
    print("tc %s %s: mean err (units of tol) %.3f, var rel %.2e" % (key, prec, e, rel_v))
    scale = {"fp16": 1.0, "bf16": 8.0}[prec]   # bf16: 3 fewer mantissa bits
    assert e <= scale, e
    assert rel_v <= vtol, rel_v
    # the value heads (128-wide nets, tanh/swish) through the same engine
    out = engine.policy_act(obs[:3000], eps=np.zeros((min(N, 3000), A), np.float32), precision=prec)
    ref = engine.policy_act(obs[:3000], eps=np.zeros((min(N, 3000), A), np.float32), precision="fp32")
    for k, tol in (("v", 2e-2), ("vc", 2e-2), ("mu", 2e-2)):
        g, w = out[k].cpu().numpy(), ref[k].cpu().numpy()
        assert np.allclose(g, w, rtol=tol * scale, atol=tol * scale), k


@pytest.mark.parametrize("N", [127, 129, 256, 257, 385, 128 * 149 + 5])
def test_cluster_pair_tile_counts(engine, N):
    """The cluster-pair kernel (two CTAs stream one copy of the weights, ens_tc.cu CL = 2) against the oracle for
    row counts around its edge cases: one tile (pairs not used), exactly one pair, an odd tile count (the last
    pair's second CTA runs the protocol on a tile without rows), more pairs than the GPU holds."""
    task, O, A = TASKS["hcs"]
    dyn, actor, v, vc = orc.make_problem(83, O, A, hidden=(512, 512), task=task)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    obs, act = orc.make_states(84, N, O, A, dyn)
    x = np.concatenate([obs, act], -1)
    mean, var = (t.cpu().numpy() for t in model.predict_ensemble_device(x, precision="fp16"))
    wm, wv = orc.pe_forward(dyn, x)
    sig = np.maximum(np.sqrt(dyn.var_out), 1e-2)
    assert mean.shape == wm.shape and np.isfinite(mean).all()
    assert _err(mean, wm, sig) <= 1.0
    assert float(np.max(np.abs(var - wv) / wv)) <= 1e-3
    # row r depends on row r only: the same rows inside a larger batch give the same bits
    if N >= 257:
        m2, v2 = (t.cpu().numpy() for t in model.predict_ensemble_device(x[:200], precision="fp16"))
        assert np.array_equal(m2, mean[:, :200]) and np.array_equal(v2, var[:, :200])


def test_rollout_fp16_vs_fp32(engine):
    """H-step: the tcgen05 rollout tracks the fp32 rollout; stated per-step tolerance 2e-3*(t+1)
    relative to the state scale, lengths equal except near-threshold flips (< 2%)."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["hcs"]
    B, T = 2048, 16
    dyn, actor, v, vc = orc.make_problem(91, O, A, hidden=(512, 512), task=task)
    obs, act = orc.make_states(92, B, O, A, dyn)
    noise = orc.TableNoise(93, T, B, A, len(dyn.elite_inds))
    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    res = {}
    for prec in ("fp32", "fp16"):
        bufs = cb.RolloutBuffers(engine, B, T, O, A)
        bufs.set_inputs(obs, noise.act_eps, noise.elite_pos)
        bufs.run(env.env_cfg(True), precision=prec)
        bufs.gae(GAE["gamma"], GAE["lam"], GAE["cost_gamma"], GAE["cost_lam"])
        res[prec] = bufs
    a, b = res["fp32"], res["fp16"]
    assert np.array_equal(a.length.cpu().numpy(), b.length.cpu().numpy())
    scale = np.sqrt(dyn.var_in[0, :O])
    na, nb = a.host("nextobs"), b.host("nextobs")
    for t in range(T - 1):
        err = np.max(np.abs(na[:, t] - nb[:, t]) / np.maximum(scale, 1e-2))
        assert err <= 2e-3 * (t + 1), (t, err)
    assert (a.host("cost") != b.host("cost")).mean() < 0.02


@pytest.mark.parametrize("num_nets,hidden,B", [(5, (256, 256), 333), (3, (128, 128), 1), (7, (512, 512), 129)])
def test_rollout_tc_other_ensemble_shapes(engine, num_nets, hidden, B):
    """Ensemble sizes other than 7 (generic fast row path, runtime E), widths 128 / 256 (grouped and
    ungrouped tcgen05 kernels), row counts that are not multiples of the 128-row tile, a single row."""
    import cmbpo_b200 as cb
    task, O, A = TASKS["ant"]
    T = 7
    dyn, actor, v, vc = orc.make_problem(95, O, A, hidden=hidden, num_nets=num_nets, num_elites=max(1, num_nets - 2), task=task)
    obs, act = orc.make_states(96, B, O, A, dyn)
    noise = orc.TableNoise(97, T, B, A, len(dyn.elite_inds))
    model, policy = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    res = {}
    for prec in ("fp32", "fp16"):
        bufs = cb.RolloutBuffers(engine, B, T, O, A)
        bufs.set_inputs(obs, noise.act_eps, noise.elite_pos)
        bufs.run(env.env_cfg(True), precision=prec)
        engine.synchronize()
        res[prec] = bufs
    a, b = res["fp32"], res["fp16"]
    la, lb = a.length.cpu().numpy(), b.length.cpu().numpy()
    assert (la != lb).mean() <= 0.02 + 1.0 / B
    same = la == lb
    scale = np.maximum(np.sqrt(dyn.var_in[0, :O]), 1e-2)
    na, nb = a.host("nextobs"), b.host("nextobs")
    da, db = a.host("dkl"), b.host("dkl")
    for t in range(T - 1):
        m = same & (la > t)
        if not m.any():
            continue
        err = np.max(np.abs(na[m, t] - nb[m, t]) / scale)
        assert err <= 3e-3 * (t + 1), (t, err)
        assert np.allclose(da[m, t], db[m, t], rtol=5e-2, atol=1e-4), t      # closed-form vs all-pairs KL, fp16 inputs


@pytest.mark.parametrize("key,O,hidden,B", [("hcs", 17, (200, 200), 500), ("hcs", 20, (512, 512), 300),
                                            ("ant", 29, (300, 300), 257), ("hcs", 20, (64, 64), 100),
                                            ("hum", 47, (200, 200), 130),
                                            ("hcs", 17, (200, 200, 200, 200), 700), ("ant", 29, (128, 128, 128), 129),
                                            ("hcs", 20, (256, 256, 256), 5000)])
def test_rollout_tc_padded_widths_and_real_hcs_obs(engine, key, O, hidden, B):
    """Hidden widths that are not a kernel instantiation (200, 300, 64: zero padded to 256 / 512 / 128 at pack
    time), three and four hidden layers (the (200,200,200,200) code default of algorithms/cmbpo.py:54: the deep
    tcgen05 variant with two hidden buffers in tensor memory) and the O = 20 observation of the real HalfCheetahSafe environment (the kernels must
    be generic in O / A): single-step prediction against the oracle within the tolerance of test_predict_ensemble_tc,
    and a short rollout against the fp32 CUDA-core path."""
    import cmbpo_b200 as cb
    task, _, A = TASKS[key]
    T = 6
    dyn, actor, v, vc = orc.make_problem(61, O, A, hidden=hidden, task=task)
    rb = np.random.default_rng(60)
    for bl in dyn.b:                                  # non-zero biases: the packed bias layout of every layer matters
        bl += (0.1 * rb.standard_normal(bl.shape)).astype(np.float32)
    obs, act = orc.make_states(62, B, O, A, dyn)
    model, policy = load_problem(engine, dyn, actor, v, vc)
    x = np.concatenate([obs, act], axis=1)
    want_m, want_v = orc.pe_forward(dyn, x)
    got_m, got_v = (t.cpu().numpy() for t in model.predict_ensemble_device(x, precision="fp16"))
    sig = np.maximum(np.sqrt(dyn.var_out), 1e-2)
    assert _err(got_m, want_m, sig) <= 1.0
    assert float(np.max(np.abs(got_v - want_v) / want_v)) <= 1.5e-3
    noise = orc.TableNoise(63, T, B, A, len(dyn.elite_inds))
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    res = {}
    for prec in ("fp32", "fp16"):
        bufs = cb.RolloutBuffers(engine, B, T, O, A)
        bufs.set_inputs(obs, noise.act_eps, noise.elite_pos)
        bufs.run(env.env_cfg(True), precision=prec)
        engine.synchronize()
        res[prec] = bufs
    a, b = res["fp32"], res["fp16"]
    la, lb = a.length.cpu().numpy(), b.length.cpu().numpy()
    assert (la != lb).mean() <= 0.02 + 1.0 / B
    same = la == lb
    scale = np.maximum(np.sqrt(dyn.var_in[0, :O]), 1e-2)
    na, nb = a.host("nextobs"), b.host("nextobs")
    va, vb = a.host("val"), b.host("val")
    for t in range(T - 1):
        m = same & (la > t)
        if m.any():
            assert np.max(np.abs(na[m, t] - nb[m, t]) / scale) <= 3e-3 * (t + 1), t
            assert np.allclose(va[m, t], vb[m, t], rtol=2e-2, atol=2e-2), t


@pytest.mark.parametrize("E,N", [(2, 20000), (5, 5000), (5, 20000)])
def test_width256_two_part_output_many_units(engine, E, N):
    """Regression: 256-wide nets with a two-part output (> 64 columns, single-buffered H2) and several
    work units per CTA dead-locked intermittently (one mbarrier shared by both epilogue pairs let a pair
    that was a phase ahead pass the parity test).  Repeated launches, checked against the fp32 path."""
    task, O, A = TASKS["hum"]
    dyn, actor, v, vc = orc.make_problem(99, O, A, hidden=(256, 256), num_nets=E, num_elites=max(1, E - 2), task=task)
    model, _ = load_problem(engine, dyn, actor, v, vc)
    obs, act = orc.make_states(98, N, O, A, dyn)
    x = engine.to_device(np.concatenate([obs, act], -1))
    m32 = model.predict_ensemble_device(x, precision="fp32")[0]
    sig = engine.to_device(np.maximum(np.sqrt(dyn.var_out), 1e-2).astype(np.float32))
    for rep in range(4):
        m16 = model.predict_ensemble_device(x, precision="fp16")[0]
        engine.synchronize()
        err = float(((m16 - m32).abs() / (1e-3 * m32.abs() + 1e-3 * sig)).max())
        assert err <= 1.0, (rep, err)
