"""CPU: pins the float64 training-step oracle (oracle/train_oracle.py) without TensorFlow -- the analytic gradient
equals central finite differences of the restated loss with the stop-gradients of pe.py:947,960-963 frozen, the
per-member loss vector has the documented structure, Adam is the published rule."""
import numpy as np

from oracle import train_oracle as tro


def _problem(seed, E=3, bs=7, Din=5, H=6, D=2, prob=True):
    rng = np.random.default_rng(seed)
    dims = [Din, H, H, 2 * D if prob else D]
    W = [rng.standard_normal((E, dims[i], dims[i + 1])) * 0.5 for i in range(3)]
    b = [rng.standard_normal((E, dims[i + 1])) * 0.1 for i in range(3)]
    x = rng.standard_normal((E, bs, Din))
    y = rng.standard_normal((E, bs, D))
    sc = dict(mu_in=rng.standard_normal((1, Din)), var_in=rng.uniform(0.5, 2, (1, Din)),
              mu_out=rng.standard_normal((1, D)), var_out=rng.uniform(0.5, 2, (1, D)))
    return W, b, ["swish", "swish", None], x, y, sc


def _scalar_loss(W, b, acts, x, y, loss, sc, frozen):
    zs, _ = tro.forward(W, b, acts, x, sc["mu_in"], sc["var_in"])
    yt = (y - sc["mu_out"]) / tro.sigma(sc["var_out"])
    if loss == "MSPE":
        return tro.mspe_frozen_scalar(zs[-1], yt, *frozen)
    return tro.mse_loss_vector(zs[-1], yt).sum()


def _check_fd(loss, prob):
    W, b, acts, x, y, sc = _problem(1, prob=prob)
    lvec, gW, gb = tro.grads(W, b, acts, x, y, loss, **sc)
    frozen = None
    if prob:
        zs, _ = tro.forward(W, b, acts, x, sc["mu_in"], sc["var_in"])
        yt = (y - sc["mu_out"]) / tro.sigma(sc["var_out"])
        _, _, s0, _, _, r0 = tro.mspe_parts(zs[-1], yt)
        frozen = (s0, r0)
    rng = np.random.default_rng(9)
    h = 1e-6
    for l in range(3):
        for arr, g in ((W[l], gW[l]), (b[l], gb[l])):
            for _ in range(12):
                idx = tuple(rng.integers(0, s) for s in arr.shape)
                old = arr[idx]
                arr[idx] = old + h
                lp = _scalar_loss(W, b, acts, x, y, loss, sc, frozen)
                arr[idx] = old - h
                lm = _scalar_loss(W, b, acts, x, y, loss, sc, frozen)
                arr[idx] = old
                fd = (lp - lm) / (2 * h)
                assert abs(fd - g[idx]) <= 1e-6 * max(1.0, abs(fd)), (l, idx, fd, g[idx])
    return lvec


def test_mspe_gradient_equals_finite_differences():
    lvec = _check_fd("MSPE", True)
    assert lvec.shape == (3,) and np.all(np.isfinite(lvec))


def test_mse_gradient_equals_finite_differences():
    _check_fd("MSE", False)


def test_mspe_loss_vector_structure():
    """total_e = mean_e s + ratio mean_e q + 0.05 mean_all logvar^2, ratio = 0.05 mean_all s / mean_all q."""
    rng = np.random.default_rng(2)
    out = rng.standard_normal((4, 6, 6))
    yt = rng.standard_normal((4, 6, 3))
    s = (out[..., :3] - yt) ** 2
    q = (np.exp(out[..., 3:]) - s) ** 2
    want = s.mean((1, 2)) + 0.05 * s.mean() / q.mean() * q.mean((1, 2)) + 0.05 * (out[..., 3:] ** 2).mean()
    np.testing.assert_allclose(tro.mspe_loss_vector(out, yt), want, rtol=1e-13)


def test_adam_first_steps_hand_computed():
    """t=1: m = 0.1 g, v = 0.001 g^2, lr_1 = lr sqrt(0.001)/0.1 -> x - lr g / (|g| + eps sqrt(1000) ...)"""
    p, g = [np.array([1.0, -2.0])], [np.array([0.5, -0.25])]
    opt = tro.Adam(p, lr=1e-3)
    new = opt.step(p, g)[0]
    lr1 = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = p[0] - lr1 * (0.1 * g[0]) / (np.sqrt(0.001 * g[0] ** 2) + 1e-8)
    np.testing.assert_allclose(new, want, rtol=1e-15)
    np.testing.assert_allclose(new, p[0] - 1e-3 * np.sign(g[0]), rtol=1e-6)     # the well-known first Adam step
    new2 = opt.step([new], g)[0]
    m2, v2 = 0.9 * 0.1 * g[0] + 0.1 * g[0], 0.999 * 0.001 * g[0] ** 2 + 0.001 * g[0] ** 2
    lr2 = 1e-3 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    np.testing.assert_allclose(new2, new - lr2 * m2 / (np.sqrt(v2) + 1e-8), rtol=1e-15)


def test_decay_coefficients():
    assert tro.decay_coeffs(3, 1e-6) == [2.5e-7, 5e-7, 1e-6]
    assert tro.decay_coeffs(5, 4.0) == [1.0, 2.0, 2.0, 2.0, 4.0]
