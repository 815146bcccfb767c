"""-m gpu: the ensemble training step (csrc/train.cu) against the float64 oracle
(oracle/train_oracle.py): per-member loss vector, gradients of every layer, three Adam steps, and the host loop
`B200PE.train` (pe.py:457-646) on a learnable synthetic regression problem.

Tolerances (stated): fp32 GEMMs -- loss rtol 2e-5, gradients within 2e-4 of the tensor's largest magnitude; TF32
tensor-core GEMMs -- loss rtol 2e-3, gradients within 5e-3 of the largest magnitude."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from oracle import train_oracle as tro

pytestmark = pytest.mark.gpu


def _setup(engine, seed, E, bs, O, A, hidden, prob=True):
    import cmbpo_b200 as cb
    from cmbpo_b200 import _lib as L
    rng = np.random.default_rng(seed)
    Din, D = O + A, O + 1
    if prob:
        dyn, _, _, _ = orc.make_problem(seed, O, A, hidden=hidden, num_nets=E, num_elites=max(1, E - 2))
        ens, which = dyn, L.NET_DYN
    else:
        _, _, v, _ = orc.make_problem(seed, O, A, hidden=(64, 64), vf_nets=E, vf_hidden=hidden)
        ens, which, Din, D = v, L.NET_V, O, 1
    for bl in ens.b:                                   # non-zero biases so that their gradients matter
        bl += 0.05 * rng.standard_normal(bl.shape).astype(np.float32)
    model = cb.B200PE.from_arrays(engine, which, ens)
    x = (rng.standard_normal((E, bs, Din)) * np.maximum(np.sqrt(ens.var_in), 1e-2) + ens.mu_in).astype(np.float32)
    y = (rng.standard_normal((E, bs, D)) * np.maximum(np.sqrt(ens.var_out), 1e-2) + ens.mu_out).astype(np.float32)
    return model, ens, which, x, y


def _f64(ens):
    return ([w.astype(np.float64) for w in ens.W], [b.astype(np.float64).reshape(b.shape[0], -1) for b in ens.b],
            dict(mu_in=ens.mu_in.astype(np.float64), var_in=ens.var_in.astype(np.float64),
                 mu_out=ens.mu_out.astype(np.float64), var_out=ens.var_out.astype(np.float64)))


@pytest.mark.parametrize("math,ltol,gtol", [("fp32", 2e-5, 2e-4), ("tf32", 2e-3, 5e-3)])
@pytest.mark.parametrize("prob,hidden", [(True, (512, 512)), (True, (200, 200, 200, 200)), (False, (128, 128))])
def test_train_step_vs_oracle(engine, math, ltol, gtol, prob, hidden):
    E, bs, O, A = (7, 256, 17, 6) if prob else (3, 256, 17, 6)
    model, ens, which, x, y = _setup(engine, 11, E, bs, O, A, hidden, prob)
    loss = "MSPE" if prob else "MSE"
    decay = 1e-2                                      # large enough to show up in the update
    model.configure_training(loss=loss, lr=1e-3, decay=decay, math=math)
    engine.train_begin(which)
    W, b, sc = _f64(ens)
    opt = tro.Adam([p for l in range(len(W)) for p in (W[l], b[l])], lr=1e-3)
    lout = engine.empty(E)
    for step in range(3):
        want_l, W2, b2, gW, gb = tro.train_step(W, b, ens.acts, x.astype(np.float64), y.astype(np.float64), loss, opt, decay, **sc)
        engine.train_step(which, x, y, model._train_cfg, lout)
        got_l = lout.cpu().numpy()
        np.testing.assert_allclose(got_l, want_l, rtol=ltol * (1 + 2 * step), err_msg="loss vector, step %d" % step)
        if step == 0:                                 # same weights on both sides: gradients are comparable
            for l in range(len(W)):
                dW, db = engine.train_grads(which, l)
                for got, want, nm in ((dW.cpu().numpy(), gW[l], "dW%d" % l), (db.cpu().numpy(), gb[l], "db%d" % l)):
                    scale = np.abs(want).max()
                    assert np.abs(got - want).max() <= gtol * scale, (nm, np.abs(got - want).max(), scale)
        W, b = W2, b2
    # after three Adam steps: every weight moved by at most 3 lr (Adam's bound), and, where the gradient is well
    # above the rounding noise, to the oracle's value
    for l in range(len(W)):
        Wd, bd = engine.get_weights(which, l)
        for got, want, g0, nm in ((Wd.cpu().numpy(), W[l], gW[l], "W%d" % l), (bd.cpu().numpy(), b[l], gb[l], "b%d" % l)):
            assert np.abs(got - want).max() <= 6.5e-3, nm
            strong = np.abs(g0) > 0.05 * np.abs(g0).max()
            # tf32: an element whose later gradients pass through zero can take one Adam step of the other sign
            tol = 3e-5 if math == "fp32" else 2.5e-3
            assert np.abs(got - want)[strong].max() <= tol, (nm, np.abs(got - want)[strong].max())


def test_train_loss_is_half_mse_on_normalised_targets(engine):
    model, ens, which, x, y = _setup(engine, 5, 7, 300, 17, 6, (512, 512))
    engine.train_begin(which)
    got = engine.train_loss(which, x, y).cpu().numpy()
    W, b, sc = _f64(ens)
    zs, _ = tro.forward(W, b, ens.acts, x.astype(np.float64), sc["mu_in"], sc["var_in"])
    yt = (y - sc["mu_out"]) / tro.sigma(sc["var_out"])
    want = (0.5 * (zs[-1][..., :18] - yt) ** 2).mean(axis=(1, 2))
    np.testing.assert_allclose(got, want, rtol=3e-5)


def test_train_loop_learns_and_repacks(engine):
    """B200PE.train: holdout loss falls on a learnable problem, elites are the best holdout members, and the
    tcgen05 prediction path sees the trained weights (it agrees with the fp32 path after training)."""
    import cmbpo_b200 as cb
    from cmbpo_b200 import _lib as L
    O, A, E = 17, 6, 7
    rng = np.random.RandomState(0)
    dyn, _, _, _ = orc.make_problem(21, O, A, hidden=(128, 128), num_nets=E, num_elites=5, gain=1.0)
    model = cb.B200PE.from_arrays(engine, L.NET_DYN, dyn)
    n = 4096
    X = rng.standard_normal((n, O + A)).astype(np.float32)
    M = (rng.standard_normal((O + A, O + 1)) * 0.3).astype(np.float32)
    Y = (np.tanh(X @ M) + 0.05 * rng.standard_normal((n, O + 1))).astype(np.float32)
    model.configure_training(loss="MSPE", lr=1e-3, decay=1e-6, use_scaler_in=True, use_scaler_out=True)
    before = model.validate(X[:512], Y[:512])
    out = model.train(X, Y, batch_size=256, max_epochs=15, holdout_ratio=0.1, rng=rng)
    after = model.validate(X[:512], Y[:512])
    assert after < 0.35 * before, (before, after)
    assert list(out.keys()) == ["PE/val_loss"] and np.isfinite(out["PE/val_loss"])
    assert len(model.elite_inds) == 5 and len(set(model.elite_inds)) == 5
    assert model._grad_updates == 15 * int(np.ceil((n - 409) / 256))
    m16, v16 = model.predict_ensemble_device(X[:1000], precision="fp16")
    m32, v32 = model.predict_ensemble_device(X[:1000], precision="fp32")
    np.testing.assert_allclose(m16.cpu().numpy(), m32.cpu().numpy(), rtol=2e-2, atol=2e-2)
    assert engine.nets[L.NET_DYN]["elite_inds"] == model.elite_inds
