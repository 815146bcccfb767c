"""`B200PE`: the model contract FakeEnv and CPOPolicy consume (models/base_model.py:3-43,
models/pens/pe.py:401-434, 648-713), served by the CUDA ensemble-forward kernels.
"""
import numpy as np

from . import _lib as L


class B200PE:
    """One ensemble living in an `Engine` slot (dynamics / V / VC).

    `predict_ensemble(x)` and `predict(x)` take and return numpy arrays like `PE` does
    (pe.py:648-713); the `*_device` variants keep everything on the GPU.
    """

    def __init__(self, engine, which, W, b, acts, probabilistic, elite_inds=(),
                 mu_in=None, var_in=None, mu_out=None, var_out=None, name="PE"):
        self.engine, self.which, self.name = engine, which, name
        self._probabilistic = bool(probabilistic)
        self.num_nets = int(W[0].shape[0])
        self._in_dim = int(W[0].shape[1])
        last = int(W[-1].shape[2])
        self._d = last // 2 if probabilistic else last
        self._model_inds = [int(i) for i in elite_inds]
        self.num_elites = len(self._model_inds)
        engine.set_network(which, W, b, acts, mu_in, var_in, mu_out, var_out,
                           probabilistic, self._model_inds)

    @classmethod
    def from_oracle_ensemble(cls, engine, which, ens, name="PE"):
        """Build from a plain container with fields W, b, acts, probabilistic, mu_in, ..."""
        return cls(engine, which, ens.W, ens.b, ens.acts, ens.probabilistic, ens.elite_inds,
                   ens.mu_in, ens.var_in, ens.mu_out, ens.var_out, name=name)

    @classmethod
    def view(cls, engine, which, name="PE"):
        """A model object over a slot whose weights were already uploaded (Engine.set_network)."""
        meta = engine.nets[which]
        self = cls.__new__(cls)
        self.engine, self.which, self.name = engine, which, name
        self._probabilistic = meta["probabilistic"]
        self.num_nets, self._in_dim, self._d = meta["E"], meta["dims"][0], meta["D"]
        self._model_inds = list(meta["elite_inds"])
        self.num_elites = len(self._model_inds)
        return self

    # --- contract properties (pe.py:401-434) ---
    @property
    def elite_inds(self):
        return self._model_inds

    @property
    def is_probabilistic(self):
        return self._probabilistic

    @property
    def is_ensemble(self):
        return self.num_nets > 1

    @property
    def in_dim(self):
        return self._in_dim

    @property
    def out_dim(self):
        # pe.py:425-430: the reference returns None for a probabilistic ensemble (missing
        # `return`); FakeEnv.output_dim is never read, so mirror it.
        return None if self._probabilistic else self._d

    @property
    def is_tf_model(self):
        return False

    # --- prediction ---
    def predict_ensemble_device(self, x, precision=None):
        return self.engine.predict_ensemble(self.which, x, precision)

    def predict_device(self, x, precision=None):
        return self.engine.predict_mean(self.which, x, precision)

    def predict_ensemble(self, inputs, *args, **kwargs):
        """pe.py:671-713: 2-D inputs -> every member on the same rows; 3-D -> member i on slice i."""
        out = self.predict_ensemble_device(np.asarray(inputs, np.float32))
        if self._probabilistic:
            return out[0].cpu().numpy(), out[1].cpu().numpy()
        return out.cpu().numpy()

    def predict(self, inputs, *args, **kwargs):
        """pe.py:648-669."""
        inputs = np.asarray(inputs, np.float32)
        assert len(inputs.shape) == 2
        out = self.predict_device(inputs)
        if self._probabilistic:
            return [out[0].cpu().numpy(), out[1].cpu().numpy()]
        return out.cpu().numpy()
