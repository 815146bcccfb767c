"""GPU: start-state sampling from the device-resident archive (cmbpo_b200.DeviceArchive,
csrc/archive.cu) against the oracle restatement of CPOBuffer.epoch_batch / boltz_dist /
distributed_batch_from_archive (buffers/cpobuffer.py:385-396, 413-524) and CPOPolicy.compute_DKL
(policies/cpo_policy.py:837-845).  Index arithmetic is bit-exact; the KL is float32 with rtol 2e-4
(TF's reduction order is unspecified); the random draws are checked in distribution."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from golden_io import load

pytestmark = pytest.mark.gpu
O, A = 17, 6


def make_archive(engine, ep, seed):
    import cmbpo_b200 as cb
    rng = np.random.default_rng(seed)
    n = len(ep)
    obs = rng.standard_normal((n, O)).astype(np.float32)
    mu = (0.3 * rng.standard_normal((n, A))).astype(np.float32)
    ls = np.broadcast_to(np.linspace(-0.9, -0.2, A, dtype=np.float32), (n, A)).copy()
    ls += (0.05 * ep.clip(0)[:, None]).astype(np.float32)
    arch = cb.DeviceArchive(engine, n, O, A, seed=5)
    arch.append(obs, mu, ls, ep)
    return arch, obs, mu, ls


@pytest.fixture()
def loaded_policy(engine):
    import cmbpo_b200 as cb
    dyn, actor, v, vc = orc.make_problem(31, O, A, hidden=(64, 64))
    pol = cb.B200Policy(engine)
    pol.load_actor(actor.W, actor.b, actor.log_std)
    pol.load_values(v, vc)
    return pol, actor


def test_index_is_a_stable_sort_by_epoch(engine):
    z = load("archive_boltz.npz")
    ep = z["epoch_archive"].astype(np.int32)
    arch, *_ = make_archive(engine, ep, 1)
    sorted_idx, offs, offs_host = arch._build_index()
    s = sorted_idx.cpu().numpy()
    n_valid = int((ep >= 0).sum())
    want = np.argsort(np.where(ep >= 0, ep, ep.max() + 1), kind="stable")[:n_valid]
    assert np.array_equal(s[:n_valid], want)
    assert np.array_equal(offs_host, np.concatenate([[0], np.cumsum(np.bincount(ep[ep >= 0]))]))
    assert np.array_equal(arch.epochs_list, z["epochs_list"])
    assert arch.min_ep == int(z["min_ep"]) and arch.max_ep == int(z["max_ep"])
    assert np.array_equal(offs.cpu().numpy(), offs_host)


def test_boltz_dist_matches_reference_golden(engine):
    z = load("archive_boltz.npz")
    arch, *_ = make_archive(engine, z["epoch_archive"].astype(np.int32), 2)
    for alpha in (1, 5.0):
        got = arch.boltz_dist(z["kls"], alpha)
        want = z["btz_alpha%g" % alpha]
        assert got.dtype == want.dtype and np.array_equal(got, want)


def test_epoch_batch_rows_and_kl(engine, loaded_policy):
    pol, actor = loaded_policy
    z = load("archive_boltz.npz")
    ep = z["epoch_archive"].astype(np.int32)
    arch, obs, mu, ls = make_archive(engine, ep, 3)
    B = 4096
    idx = arch.epoch_batch_indices(B).cpu().numpy()
    eps = arch.epochs_list
    assert idx.shape == (len(eps), B)
    for k, e in enumerate(eps):
        assert (ep[idx[k]] == e).all()                                 # every row belongs to its epoch
        # uniform over the epoch's rows: mean rank near the middle (5 sigma)
        rows = np.nonzero(ep == e)[0]
        r = np.searchsorted(rows, idx[k]) / len(rows)
        assert abs(r.mean() - 0.5) < 5 * np.sqrt(1 / 12 / B) + 1 / len(rows)
    # KL per epoch on EXACTLY the sampled rows vs the oracle (same draw: reseed the counter)
    arch._draw = 0
    arch2_idx = arch.epoch_batch_indices(B).cpu().numpy()
    arch._draw = 0
    kls = arch.epoch_kls(pol, B)
    want = np.array([orc.policy_kl(actor, obs[i], mu[i], ls[i]) for i in arch2_idx], np.float64)
    assert np.allclose(kls, want, rtol=2e-4, atol=1e-6), (kls, want)


def test_distributed_batch_follows_the_boltzmann_distribution(engine, loaded_policy):
    pol, actor = loaded_policy
    z = load("archive_boltz.npz")
    ep = z["epoch_archive"].astype(np.int32)
    arch, obs, mu, ls = make_archive(engine, ep, 4)
    B = 200000
    kls = z["kls"]
    idx = arch.distributed_batch_indices(B, kls, alpha=5.0).cpu().numpy()
    assert (idx >= 0).all() and (ep[idx] >= 0).all()
    dist = orc.boltz_dist(z["epoch_archive"], kls, 5.0).astype(np.float64)
    eps = arch.epochs_list
    p_ep = np.array([dist[ep == e].sum() for e in eps])
    p_ep /= p_ep.sum()
    freq = np.array([(ep[idx] == e).mean() for e in eps])
    assert np.all(np.abs(freq - p_ep) < 5 * np.sqrt(p_ep * (1 - p_ep) / B) + 1e-4), (freq, p_ep)
    # reproducible for a (seed, draw) and different across draws
    arch._draw = 0
    a = arch.distributed_batch_indices(1000, kls, 5.0).cpu().numpy()
    b = arch.distributed_batch_indices(1000, kls, 5.0).cpu().numpy()
    arch._draw = 0
    c = arch.distributed_batch_indices(1000, kls, 5.0).cpu().numpy()
    assert np.array_equal(a, c) and not np.array_equal(a, b)


def test_sample_start_states_feeds_the_sampler(engine, loaded_policy):
    import cmbpo_b200 as cb
    pol, actor = loaded_policy
    z = load("archive_boltz.npz")
    ep = z["epoch_archive"].astype(np.int32)
    arch, obs, mu, ls = make_archive(engine, ep, 5)
    out = arch.sample_start_states(pol, 512, alpha=1)
    idx = out["indices"].cpu().numpy()
    assert np.array_equal(out["observations"].cpu().numpy(), obs[idx])
    assert np.array_equal(out["mu"].cpu().numpy(), mu[idx]) and np.array_equal(out["log_std"].cpu().numpy(), ls[idx])
    assert out["observations"].is_cuda and (out["kls"] >= 0).all()
