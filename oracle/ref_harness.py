"""TEST INFRASTRUCTURE ONLY.  Drives the reference's UNMODIFIED FakeEnv / ModelSampler /
ModelBuffer / CPOBuffer (imported through `oracle/ref_stubs.py`, build container only)
on the same weights and the same injected noise the oracle and the CUDA path see.

The two TF graphs are supplied by the numpy restatement (`OracleModel`, `OraclePolicy`);
everything else that runs here is the reference's own code.
"""
import numpy as np

from . import cmbpo_oracle as orc
from . import ref_stubs


class _RefNoise:
    """Adapts `TableNoise` (keyed by step / global path id) to the reference objects,
    which know nothing about `ctx`: the key is recovered from the sampler/pool state."""

    def __init__(self, noise, sampler, pool):
        self.noise, self.sampler, self.pool = noise, sampler, pool

    def _key(self):
        return self.sampler._n_episodes - 1, np.flatnonzero(self.pool.alive_paths)

    def eps_fn(self, policy, n):
        t, ids = self._key()
        assert len(ids) == n
        return self.noise.act_eps[t, ids]

    def elite_pos(self, n):
        t, ids = self._key()
        assert len(ids) == n
        return self.noise.elite_pos[t, ids]


def reference_rollout(dyn, actor, v, vc, task, start_obs, noise, max_path_length,
                      rollout_mode=False, dkl_lim=None, max_samples=None,
                      stop_alive_ratio=None, gamma=0.99, lam=0.95, cgamma=0.97, clam=0.5,
                      n_steps=None):
    """cmbpo.py:251-269 with the reference's own sampler, buffer and fake env."""
    ref = ref_stubs.load()
    B, O = start_obs.shape
    A = actor.W[-1].shape[1]
    model = orc.OracleModel(dyn)
    env = ref.FakeEnv(ref_stubs.ShapeEnv(O, A), task, model, True, True, False)
    pool = ref.ModelBuffer(B, O, A, max_path_length)
    pool.initialize({"mu": (A,), "log_std": (A,)}, gamma=gamma, lam=lam,
                    cost_gamma=cgamma, cost_lam=clam)
    smp = ref.ModelSampler(max_path_length, B, rollout_mode, logger=object())
    adapter = _RefNoise(noise, smp, pool)
    policy = orc.OraclePolicy(actor, v, vc, adapter.eps_fn)
    elites = np.asarray(model.elite_inds)
    env.random_inds = lambda size: elites[adapter.elite_pos(size)]   # fake_env.py:174-176
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
    snap = {"obs": pool.obs_buf, "act": pool.act_buf, "nextobs": pool.nextobs_buf,
            "rew": pool.rew_buf, "val": pool.val_buf, "cost": pool.cost_buf,
            "cval": pool.cval_buf, "logp": pool.logp_buf, "dyn_error": pool.dyn_error_buf,
            "adv": pool.adv_buf, "ret": pool.ret_buf, "cadv": pool.cadv_buf,
            "cret": pool.cret_buf, "term": pool.term_buf, "populated": pool.populated_mask,
            "mu": pool.pi_info_bufs["mu"], "log_std": pool.pi_info_bufs["log_std"]}
    snap = {k: np.array(a) for k, a in snap.items()}
    out, bdiag = pool.get()
    return out, bdiag, dict(diag), snap


def reference_fakeenv_step(dyn, task, obs, act, elite_pos):
    ref = ref_stubs.load()
    O, A = obs.shape[1], act.shape[1]
    model = orc.OracleModel(dyn)
    env = ref.FakeEnv(ref_stubs.ShapeEnv(O, A), task, model, True, True, False)
    elites = np.asarray(model.elite_inds)
    env.random_inds = lambda size: elites[np.asarray(elite_pos)]
    return env.step(obs, act)
