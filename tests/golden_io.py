"""Load tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import os

import numpy as np

from oracle import cmbpo_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def problem(z):
    def ens(name, prob):
        n = len([k for k in z.files if k.startswith(name + "_W")])
        W = [z["%s_W%d" % (name, i)] for i in range(n)]
        b = [z["%s_b%d" % (name, i)] for i in range(n)]
        return orc.Ensemble(W, b, ["swish"] * (n - 1) + [None], prob, z[name + "_mu_in"],
                            z[name + "_var_in"], z[name + "_mu_out"], z[name + "_var_out"],
                            [int(i) for i in z[name + "_elite"]])
    n = len([k for k in z.files if k.startswith("actor_W")])
    actor = orc.Actor([z["actor_W%d" % i] for i in range(n)], [z["actor_b%d" % i] for i in range(n)],
                      z["actor_log_std"])
    return ens("dyn", True), actor, ens("v", False), ens("vc", False)


class FixedNoise:
    """TableNoise rebuilt from the arrays stored in a fixture."""

    def __init__(self, act_eps, elite_pos):
        self.act_eps, self.elite_pos, self.state_eps = act_eps, elite_pos, None

    def eps_fn(self, policy, n):
        t, ids = policy.ctx
        return self.act_eps[t, ids]

    def idx_fn(self, env, n):
        t, ids = env.ctx
        return self.elite_pos[t, ids]
