"""Task tables of models/statics.py:56-70 as the ids the kernels take.

The cost / termination functions themselves are evaluated in-register by the rollout
kernels (csrc/row_math.cuh: apply_statics) with the reference's exact semantics, including
the operator-precedence behaviour of antsafe_term_fn (statics.py:24-27).
"""
from . import _lib as L

TERMS_BY_TASK = {
    "default": L.TERM_NO_DONE,
    "HalfCheetah-v2": L.TERM_NO_DONE,
    "HalfCheetahSafe-v2": L.TERM_NO_DONE,
    "AntSafe-v2": L.TERM_ANTSAFE,
}

COST_BY_TASK = {
    "HalfCheetahSafe-v2": L.COST_HCS,
    "AntSafe-v2": L.COST_ANTSAFE,
}

REWS_BY_TASK = {}


def task_ids(task):
    """fake_env.py:134-146: unknown tasks use `default` terminations and a zero cost."""
    return (TERMS_BY_TASK.get(task, TERMS_BY_TASK["default"]), COST_BY_TASK.get(task, L.COST_ZERO))
