"""TEST INFRASTRUCTURE (oracle): float64 numpy restatement of one training step of the probabilistic ensemble
-- the loss tf builds in models/pens/pe.py:921-973 (`_mspe_loss`) / 840-919 (`_nll_loss`, inc_var_loss=False),
the decays of models/pens/fc.py:168-169 with the coefficients of models/pens/pe_factory.py:50-55, and
tf.train.AdamOptimizer's update.  Only tests/ may import this module.

Parity status: the reference's graph needs TensorFlow 1.14 (not installable here), so this file is a restatement
("parity unpinned" by reference-produced vectors).  It is pinned instead by tests/test_train_oracle_cpu.py: the
analytic gradient below equals central finite differences of the restated loss (with tf.stop_gradient applied
exactly where pe.py:947,960-963 applies it), and Adam equals a literal transcription of the published update rule.
"""
import numpy as np


def swish(x):                                  # fc.py:19
    return x / (1.0 + np.exp(-x))


def swish_grad(x):
    s = 1.0 / (1.0 + np.exp(-x))
    return s * (1.0 + x * (1.0 - s))


ACT = {"swish": (swish, swish_grad), "tanh": (np.tanh, lambda x: 1.0 - np.tanh(x) ** 2),
       None: (lambda x: x, lambda x: np.ones_like(x))}


def sigma(var):                                # pens/utils.py:156
    return np.maximum(np.sqrt(var), 1e-2)


def forward(W, b, acts, x, mu_in=None, var_in=None):
    """x [E, bs, in] -> list of pre-activations z_l and activations h_l (fc.py:74-95, 3-D branch)."""
    a = x if mu_in is None else (x - mu_in) / sigma(var_in)
    zs, hs = [], [a]
    for Wl, bl, act in zip(W, b, acts):
        z = np.einsum("ebi,eio->ebo", a, Wl) + bl.reshape(Wl.shape[0], 1, -1)
        zs.append(z)
        a = ACT[act][0](z)
        hs.append(a)
    return zs, hs


def mspe_parts(out, yt):
    """pe.py:945-971 on raw outputs [E, bs, 2D] and transformed targets [E, bs, D]."""
    D = yt.shape[-1]
    mean, logvar = out[..., :D], out[..., D:]
    s = (mean - yt) ** 2                       # mse_losses_logit
    v = np.exp(logvar)                         # var_pred
    q = (v - s) ** 2                           # var_losses_logit (s under stop_gradient)
    ratio = 0.05 * s.mean() / q.mean()         # under stop_gradient
    return mean, logvar, s, v, q, ratio


def mspe_loss_vector(out, yt):
    """`total_losses` [E] as `_mspe_loss` returns it."""
    mean, logvar, s, v, q, ratio = mspe_parts(out, yt)
    return s.mean(axis=(1, 2)) + (q * ratio).mean(axis=(1, 2)) + 0.05 * np.mean(logvar ** 2)


def mspe_frozen_scalar(out, yt, s0, ratio0):
    """sum_e total_losses with the stop-gradient terms frozen at (s0, ratio0): what tf differentiates."""
    D = yt.shape[-1]
    mean, logvar = out[..., :D], out[..., D:]
    s = (mean - yt) ** 2
    q = (np.exp(logvar) - s0) ** 2
    E = out.shape[0]
    return s.mean(axis=(1, 2)).sum() + (q * ratio0).mean(axis=(1, 2)).sum() + E * 0.05 * np.mean(logvar ** 2)


def mse_loss_vector(out, yt):                  # pe.py:914 with weights = 1, inv_var = 1
    return (0.5 * (out - yt) ** 2).mean(axis=(1, 2))


def grads(W, b, acts, x, y, loss, mu_in=None, var_in=None, mu_out=None, var_out=None):
    """Loss vector [E] and the gradients of train_loss = sum_e loss_e (decays excluded) for every W_l, b_l."""
    zs, hs = forward(W, b, acts, x, mu_in, var_in)
    out = zs[-1]
    yt = y if mu_out is None else (y - mu_out) / sigma(var_out)
    E, bs = out.shape[0], out.shape[1]
    D = yt.shape[-1]
    n = bs * D
    if loss == "MSPE":
        mean, logvar, s, v, q, ratio = mspe_parts(out, yt)
        lvec = mspe_loss_vector(out, yt)
        g = np.concatenate([2.0 * (mean - yt) / n, (2.0 * ratio * (v - s) * v + 0.1 * logvar) / n], axis=-1)
    else:
        lvec = mse_loss_vector(out, yt)
        g = (out - yt) / n
    gW, gb = [None] * len(W), [None] * len(W)
    for l in range(len(W) - 1, -1, -1):
        gW[l] = np.einsum("ebi,ebo->eio", hs[l], g)
        gb[l] = g.sum(axis=1)
        if l > 0:
            g = np.einsum("ebo,eio->ebi", g, W[l]) * ACT[acts[l - 1]][1](zs[l - 1])
    return lvec, gW, gb


def decay_coeffs(n_layers, decay):             # pe_factory.py:50-55
    return [decay / 4 if l == 0 else (decay if l == n_layers - 1 else decay / 2) for l in range(n_layers)]


class Adam:
    """tf.train.AdamOptimizer (Kingma & Ba, algorithm 1 with the epsilon-hat form TF documents):
    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2; x <- x - lr_t m / (sqrt(v) + eps)."""

    def __init__(self, params, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m = [np.zeros_like(p, dtype=np.float64) for p in params]
        self.v = [np.zeros_like(p, dtype=np.float64) for p in params]

    def step(self, params, grads_):
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        out = []
        for i, (p, g) in enumerate(zip(params, grads_)):
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            out.append(p - lr_t * self.m[i] / (np.sqrt(self.v[i]) + self.eps))
        return out


def train_step(W, b, acts, x, y, loss, opt, decay, **scalers):
    """One reference train_op: gradients of sum_e loss_e + sum_l wd_l * l2_loss(W_l), Adam on [W0, b0, W1, b1, ...]."""
    lvec, gW, gb = grads(W, b, acts, x, y, loss, **scalers)
    wd = decay_coeffs(len(W), decay)
    params, gs = [], []
    for l in range(len(W)):
        params += [W[l], b[l]]
        gs += [gW[l] + wd[l] * W[l], gb[l].reshape(b[l].shape)]
    new = opt.step(params, gs)
    return lvec, new[0::2], new[1::2], gW, gb
