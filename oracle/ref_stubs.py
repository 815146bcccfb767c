"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Makes the *unmodified* reference host code importable in the build container
(`/root/reference` exists only here; it does NOT exist on the GPU box), so the
numpy restatement in `oracle/cmbpo_oracle.py` can be pinned against it and
golden vectors can be generated (`oracle/gen_golden.py`).

TensorFlow 1.14, mpi4py, gym, ray, gtimer, dotmap and mujoco_py are not
installable here (Python 3.12, no network).  They are replaced by inert
`MagicMock` modules; none of them is reached by the host logic on the hot path
(FakeEnv.step, ModelSampler.sample, ModelBuffer, CPOBuffer.finish_path/get,
statics, average_dkl, discount_cumsum) -- those are pure numpy/scipy.
"""
import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("CMBPO_REFERENCE_ROOT", "/root/reference")

_STUBBED = [
    "tensorflow", "tensorflow.python", "tensorflow.python.ops",
    "tensorflow.python.ops.math_ops", "tensorflow.python.framework",
    "tensorflow.python.framework.ops", "tensorflow.contrib",
    "tensorflow.contrib.losses", "tensorflow.python.lib",
    "tensorflow.python.lib.io", "tensorflow.python.lib.io.file_io",
    "gym", "gym.spaces", "gym.envs", "gym.envs.registration", "gtimer",
    "dotmap", "ray", "wrappers", "mujoco_py", "envs.mujoco_safety_gym",
]


class _OneRankComm:
    """mpi4py.MPI.COMM_WORLD for a world of one (utilities/mpi_tools.py:43-62)."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Allreduce(self, x, buf, op=None):
        buf[...] = x

    def Bcast(self, x, root=0):
        return None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "pens"))


_installed = False


def install():
    """Idempotently install the stubs and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBBED:
        sys.modules.setdefault(name, mock.MagicMock(name=name))
    mpi4py = types.ModuleType("mpi4py")
    mpi = types.SimpleNamespace(COMM_WORLD=_OneRankComm(), SUM="sum", MIN="min", MAX="max")
    mpi4py.MPI = mpi
    sys.modules["mpi4py"] = mpi4py
    # numpy 2 dropped the aliases the reference still uses (cpobuffer.py:135, statics.py:6)
    for alias, target in (("int", int), ("bool", bool), ("float", float)):
        if not hasattr(np, alias):
            setattr(np, alias, target)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load():
    """Return a namespace with the reference's hot-path symbols."""
    install()
    from models.fake_env import FakeEnv
    from models import statics
    from models.pens.utils import average_dkl, gaussian_kl_np
    from samplers.model_sampler import ModelSampler
    from buffers.modelbuffer import ModelBuffer
    from buffers.cpobuffer import CPOBuffer
    from utilities.utils import discount_cumsum
    from utilities.mpi_tools import mpi_statistics_scalar
    return types.SimpleNamespace(
        FakeEnv=FakeEnv, statics=statics, average_dkl=average_dkl,
        gaussian_kl_np=gaussian_kl_np, ModelSampler=ModelSampler,
        ModelBuffer=ModelBuffer, CPOBuffer=CPOBuffer,
        discount_cumsum=discount_cumsum,
        mpi_statistics_scalar=mpi_statistics_scalar)


class Space:
    """Stand-in for gym.spaces.Box: FakeEnv/CPOBuffer only read `.shape`."""

    def __init__(self, dim):
        self.shape = (int(dim),)


class ShapeEnv:
    """Stand-in for the MuJoCo env FakeEnv is built around (fake_env.py:40-42, 58-64)."""

    def __init__(self, obs_dim, act_dim):
        self.observation_space = Space(obs_dim)
        self.action_space = Space(act_dim)
