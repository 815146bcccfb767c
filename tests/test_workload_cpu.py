"""CPU: the product-side workload generator (cmbpo_b200/workload.py, used by bench.py and tools/)
produces exactly the synthetic problems of the oracle's generator, so the CUDA arm and the CPU
baseline of bench.py run the same weights and start states."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from cmbpo_b200 import workload as wl


@pytest.mark.parametrize("task,O,A,hidden", [("HalfCheetahSafe-v2", 17, 6, (512, 512)), ("AntSafe-v2", 29, 8, (64, 64)),
                                              ("HumanoidSafe-v2", 47, 17, (32, 32))])
def test_same_problem_as_the_oracle_generator(task, O, A, hidden):
    a = orc.make_problem(3, O, A, hidden=hidden, task=task)
    b = wl.make_problem(3, O, A, hidden=hidden, task=task)
    for x, y in zip(a, b):
        for wx, wy in zip(x.W, y.W):
            assert wx.dtype == wy.dtype and np.array_equal(wx, wy)
        for bx, by in zip(x.b, y.b):
            assert np.array_equal(bx, by)
    for k in (0, 2, 3):
        for name in ("mu_in", "var_in", "mu_out", "var_out"):
            assert np.array_equal(getattr(a[k], name), getattr(b[k], name)), name
        assert list(a[k].elite_inds) == list(b[k].elite_inds) and list(a[k].acts) == list(b[k].acts)
    assert np.array_equal(a[1].log_std, b[1].log_std)
    oa, aa = orc.make_states(4, 500, O, A, a[0])
    ob, ab = wl.make_states(4, 500, O, A, b[0])
    assert np.array_equal(oa, ob) and np.array_equal(aa, ab)
