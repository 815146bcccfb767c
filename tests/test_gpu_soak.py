"""GPU: soak test -- 120 rollouts of random task shapes, ensemble sizes, widths, row counts (1 ... 50 000),
horizons, precisions and modes in one process (tools/soak.py).  A protocol dead-lock in the tcgen05
kernel traps after ~4.5 s and fails the run; every stored value must be finite."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_soak_random_shapes(seed):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak.py"), str(seed)], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "soak ok: 120 rollouts" in out.stdout
