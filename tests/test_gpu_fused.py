"""The fused rollout step (policy head + FakeEnv row math + sampler rules + ModelBuffer write-out inside
the tcgen05 dynamics kernel, two launches per step) against the step-wise tensor-core path
(the default: four launches per step, raw [E,B,2D] outputs through HBM).  Both run literally the
same per-row functions (csrc/step_common.cuh, row_math.cuh) on the same GEMM results, in the same
summation order, so every field must be BIT-identical -- whatever parity the step-wise path has with the
oracle (tests/test_gpu_tc.py, test_gpu_parity_tc.py) the fused path inherits exactly."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from helpers import TASKS, GAE, load_problem, ShapeEnv, calibrated_dkl_lim

pytestmark = pytest.mark.gpu

FIELDS = ("length", "end_reason", "last_val", "last_cval", "obs", "nextobs", "act", "mu", "rew", "val", "cval",
          "logp", "cost", "dkl", "dyn_error", "term", "adv", "ret", "cadv", "cret", "cum_dkl", "path_return",
          "path_cost", "final_obs", "step_stats")


def _run(engine, cfg, B, T, O, A, obs, noise, flags, prec, **kw):
    import cmbpo_b200 as cb
    bufs = cb.RolloutBuffers(engine, B, T, O, A)
    if noise is None:
        bufs.set_inputs(obs)
    else:
        bufs.set_inputs(obs, noise.act_eps, noise.elite_pos, noise.__dict__.get("state_eps"))
    launches0 = engine.launch_count
    bufs.run(cfg, precision=prec, flags=flags, **kw)
    bufs.launches = engine.launch_count - launches0
    bufs.gae(GAE["gamma"], GAE["lam"], GAE["cost_gamma"], GAE["cost_lam"])
    engine.synchronize()
    return bufs


def _assert_identical(engine, a, b, skip=()):
    t = engine.torch
    for name in FIELDS:
        if name in skip:
            continue
        x, y = getattr(a, name), getattr(b, name)
        if name == "step_stats":        # float64 sums accumulated with atomics: the order differs
            assert t.allclose(x, y, rtol=1e-12, atol=0), name
            continue
        assert bool(t.equal(x, y)), name


@pytest.mark.parametrize("key,B,T", [("hcs", 1500, 12), ("ant", 1111, 10), ("hum", 700, 8)])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_fused_equals_stepwise_injected_noise(engine, key, B, T, prec):
    from cmbpo_b200 import _lib as L
    import cmbpo_b200 as cb
    task, O, A = TASKS[key]
    dyn, actor, v, vc = orc.make_problem(301, O, A, hidden=(512, 512), task=task)
    model, _ = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    obs, act = orc.make_states(302, B, O, A, dyn)
    noise = orc.TableNoise(303, T, B, A, len(dyn.elite_inds))
    cfg = env.env_cfg(True)
    fused = _run(engine, cfg, B, T, O, A, obs, noise, L.ROLLOUT_FUSE, prec)
    step = _run(engine, cfg, B, T, O, A, obs, noise, 0, prec)
    assert fused.launches < step.launches, (fused.launches, step.launches)
    if key == "hcs":                                   # no terminations: no compaction launches
        assert fused.launches <= 2 * (T - 1) + 6, fused.launches
    assert int(fused.length.sum().item()) > 0
    _assert_identical(engine, fused, step)


@pytest.mark.parametrize("key,hidden,B", [("hcs", (512, 512), 40000), ("ant", (256, 256), 20000), ("hum", (128, 128), 3000)])
def test_fused_equals_stepwise_philox_uncertainty(engine, key, hidden, B):
    """Philox noise, the reference's default 'uncertainty' rollout mode (paths end at many different steps),
    alive-row compaction on and off, many tiles per CTA and tiles shared by two CTAs (B >> 148 * 128)."""
    from cmbpo_b200 import _lib as L
    import cmbpo_b200 as cb
    task, O, A = TASKS[key]
    T = 20
    dyn, actor, v, vc = orc.make_problem(311, O, A, hidden=hidden, task=task)
    model, _ = load_problem(engine, dyn, actor, v, vc)
    env = cb.FakeEnv(ShapeEnv(O, A), task, model, True, True, False)
    obs, act = orc.make_states(312, B, O, A, dyn)
    lim = calibrated_dkl_lim(dyn, task, obs[:2000], act[:2000], factor=12.0)
    cfg = env.env_cfg(True)
    kw = dict(uncertainty_mode=True, dkl_lim=lim, seed=5)
    ref = _run(engine, cfg, B, T, O, A, obs, None, L.ROLLOUT_NO_COMPACT, "fp16", **kw)
    ln = ref.length.cpu().numpy()
    assert 1 < ln.mean() < T - 2 and len(np.unique(ln)) > 5
    for flags in (L.ROLLOUT_FUSE, L.ROLLOUT_FUSE | L.ROLLOUT_NO_COMPACT, 0):
        got = _run(engine, cfg, B, T, O, A, obs, None, flags, "fp16", **kw)
        # final_obs of ended paths is whatever row the compaction left behind: compare alive paths only
        _assert_identical(engine, got, ref, skip=("final_obs",))


def test_fused_state_noise_mode(engine):
    """deterministic=False (`mean + std * eps`, the injected-noise mode of BASELINE configs[1])."""
    from cmbpo_b200 import _lib as L
    task, O, A = TASKS["hcs"]
    B, T = 3000, 9
    dyn, actor, v, vc = orc.make_problem(321, O, A, hidden=(512, 512), task=task)
    load_problem(engine, dyn, actor, v, vc)
    obs, act = orc.make_states(322, B, O, A, dyn)
    cfg = L.EnvCfg(L.TERM_NO_DONE, L.COST_HCS, 0, 0, 1)          # deterministic = 0
    rng = np.random.default_rng(5)
    noise = orc.TableNoise(323, T, B, A, len(dyn.elite_inds))
    noise.state_eps = rng.standard_normal((T, B, O)).astype(np.float32)
    for nz in (noise, None):                                      # injected arrays, then Philox
        fused = _run(engine, cfg, B, T, O, A, obs, nz, L.ROLLOUT_FUSE, "fp16", seed=9)
        step = _run(engine, cfg, B, T, O, A, obs, nz, 0, "fp16", seed=9)
        _assert_identical(engine, fused, step)


def test_no_store_mode_gives_the_same_statistics(engine):
    """ROLLOUT_NO_STORE: per-path results and step statistics without the per-step fields."""
    from cmbpo_b200 import _lib as L
    task, O, A = TASKS["ant"]
    B, T = 2000, 8
    dyn, actor, v, vc = orc.make_problem(331, O, A, hidden=(512, 512), task=task)
    load_problem(engine, dyn, actor, v, vc)
    obs, act = orc.make_states(332, B, O, A, dyn)
    cfg = L.EnvCfg(L.TERM_ANTSAFE, L.COST_ANTSAFE, 0, 1, 1)
    t = engine.torch
    for prec in ("fp32", "fp16"):
        full = _run(engine, cfg, B, T, O, A, obs, None, 0, prec, seed=3)
        lean = _run(engine, cfg, B, T, O, A, obs, None, L.ROLLOUT_NO_STORE, prec, seed=3)
        for name in ("length", "end_reason", "last_val", "last_cval", "cum_dkl", "path_return", "path_cost"):
            assert bool(t.equal(getattr(full, name), getattr(lean, name))), (prec, name)
        assert t.allclose(full.step_stats, lean.step_stats, rtol=1e-12, atol=0)
        assert float(lean.rew.abs().sum().item()) == 0.0 and float(lean.nextobs.abs().sum().item()) == 0.0
