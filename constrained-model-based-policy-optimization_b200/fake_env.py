"""`FakeEnv` with the reference's constructor and `step` contract (models/fake_env.py:15-198),
served by cmbpo_fakeenv_step: ensemble forward + Gaussian head + across-member KL / variance +
elite pick + statics in one pass on the GPU.
"""
import numpy as np

from . import _lib as L
from .statics import task_ids


class FakeEnv:
    def __init__(self, true_environment, task, model, predicts_delta, predicts_rew, predicts_cost):
        self.env = true_environment
        self.obs_dim = int(np.prod(self.observation_space.shape))
        self.act_dim = int(np.prod(self.action_space.shape))
        self._task = task
        self._model = model
        if not (model.is_ensemble and model.is_probabilistic):
            raise NotImplementedError("the CUDA FakeEnv serves probabilistic ensembles (cmbpo.py:120-142)")
        if not predicts_rew:
            raise AssertionError("Please provide either static functions or predictions for rewards, "
                                 "costs and terms")      # fake_env.py:155 (REWS_BY_TASK is empty)
        self._predicts_delta, self._predicts_rew, self._predicts_cost = \
            bool(predicts_delta), bool(predicts_rew), bool(predicts_cost)
        self.input_dim = model.in_dim
        self.output_dim = model.out_dim
        self.engine = model.engine
        self.term_id, self.cost_id = task_ids(task)
        self._steps = 0

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def action_space(self):
        return self.env.action_space

    def env_cfg(self, deterministic=True):
        return L.EnvCfg(self.term_id, self.cost_id, int(self._predicts_cost), int(deterministic),
                        int(self._predicts_delta))

    def random_inds(self, size):
        """fake_env.py:174-178: consumes the global numpy stream exactly like
        np.random.choice(elite_inds, size) and returns the chosen *positions*."""
        return np.random.randint(0, len(self._model.elite_inds), size)

    def step(self, obs, act, deterministic=True, elite_pos=None, state_eps=None):
        obs = np.asarray(obs, np.float32)
        act = np.asarray(act, np.float32)
        assert len(obs.shape) == len(act.shape)
        assert obs.shape[-1] == self.obs_dim and act.shape[-1] == self.act_dim
        if obs.ndim == 3:
            raise NotImplementedError("3-D inputs (fake_env.py:84-101) are not used by ModelSampler")
        single = obs.ndim == 1
        if single:
            obs, act = obs[None], act[None]
        n = obs.shape[0]
        if elite_pos is None:
            elite_pos = self.random_inds(n)
        out = self.engine.fakeenv_step(self.env_cfg(deterministic), obs, act,
                                       elite_pos=np.asarray(elite_pos, np.int32),
                                       state_eps=state_eps, step=self._steps)
        self._steps += 1
        next_obs = out["next_obs"].cpu().numpy()
        r = out["rew"].cpu().numpy()[:, None]
        terms = out["term"].cpu().numpy().astype(bool)[:, None]
        c = out["cost"].cpu().numpy()[:, None]
        if not self._predicts_cost:
            if self.cost_id == L.COST_ZERO:
                c = np.zeros_like(terms)                  # fake_env.py:146: a bool array
            elif self.cost_id == L.COST_ANTSAFE:
                c = c.astype(np.float64)                  # statics.py:52 returns float64
        info = {"ensemble_dkl_mean": out["dkl_mean"].cpu().numpy()[0],
                "ensemble_dkl_path": out["dkl_path"].cpu().numpy(),
                "ensemble_ep_var": out["ep_var"].cpu().numpy(),
                "rew": r, "cost": c}
        if single:
            next_obs, r, terms = next_obs[0], r[0], terms[0]
            info["rew"], info["cost"] = r, c[0]
        return next_obs, r, terms, info

    def close(self):
        pass
