"""Shared test plumbing: load an oracle problem into an Engine, run the oracle rollout."""
import numpy as np

from oracle import cmbpo_oracle as orc

TASKS = {"hcs": ("HalfCheetahSafe-v2", 17, 6), "ant": ("AntSafe-v2", 29, 8),
         "hum": ("HumanoidSafe-v2", 47, 17)}
GAE = dict(gamma=0.99, lam=0.95, cost_gamma=0.97, cost_lam=0.5)


def load_problem(engine, dyn, actor, v, vc):
    import cmbpo_b200 as cb
    from cmbpo_b200 import _lib as L
    model = cb.B200PE.from_arrays(engine, L.NET_DYN, dyn, name="DynEns")
    policy = cb.B200Policy(engine)
    policy.load_actor(actor.W, actor.b, actor.log_std)
    policy.load_values(v, vc)
    return model, policy


class ShapeEnv:
    class _S:
        def __init__(self, n):
            self.shape = (n,)

    def __init__(self, O, A):
        self.observation_space, self.action_space = self._S(O), self._S(A)


def calibrated_dkl_lim(dyn, task, obs, act, factor=4.0):
    O, A = obs.shape[1], act.shape[1]
    env = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), lambda e, n: np.zeros(n, int))
    _, _, _, info = env.step(obs, act)
    return float(np.median(info["ensemble_dkl_path"]) * factor)


def margin_mask_dkl(snapshot_dkl_cum, lim, rel=1e-4):
    """paths whose cumulative KL never comes within `rel` of the limit (discrete outcome stable)."""
    return np.all(np.abs(snapshot_dkl_cum - lim) > rel * lim, axis=1)
