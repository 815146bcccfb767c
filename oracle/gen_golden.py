"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the reference's UNMODIFIED
host code (imported from /root/reference through oracle/ref_stubs.py -- build container only) on
seeded inputs.  The fixtures carry the inputs (weights included) and the reference's outputs, so
the oracle (CPU tests) and the CUDA path (GPU tests) can both be checked where /root/reference
does not exist.

The reference has no golden vectors of its own; these are self-generated
from its code and labelled as such.  The two TF graphs (PE forward, Gaussian actor) cannot be
executed (TF 1.14); wherever they are needed the numpy restatement supplies them, and the fixture
name says so ("tfrestated").

    python -m oracle.gen_golden        # from the repo root
"""
import os

import numpy as np

from . import cmbpo_oracle as orc
from . import ref_harness, ref_stubs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TASKS = {"hcs": ("HalfCheetahSafe-v2", 17, 6), "ant": ("AntSafe-v2", 29, 8),
         "hum": ("HumanoidSafe-v2", 47, 17)}
GAE = dict(gamma=0.99, lam=0.95, cgamma=0.97, clam=0.5)


def pack_problem(dyn, actor, v, vc):
    d = {}
    for name, ens in (("dyn", dyn), ("v", v), ("vc", vc)):
        for i, (W, b) in enumerate(zip(ens.W, ens.b)):
            d["%s_W%d" % (name, i)], d["%s_b%d" % (name, i)] = W, b
        for k in ("mu_in", "var_in", "mu_out", "var_out"):
            d["%s_%s" % (name, k)] = getattr(ens, k)
        d["%s_elite" % name] = np.asarray(ens.elite_inds, np.int32)
    for i, (W, b) in enumerate(zip(actor.W, actor.b)):
        d["actor_W%d" % i], d["actor_b%d" % i] = W, b
    d["actor_log_std"] = actor.log_std
    return d


def unpack_problem(z):
    def ens(name, prob):
        n = len([k for k in z.files if k.startswith(name + "_W")])
        W = [z["%s_W%d" % (name, i)] for i in range(n)]
        b = [z["%s_b%d" % (name, i)] for i in range(n)]
        acts = ["swish"] * (n - 1) + [None]
        return orc.Ensemble(W, b, acts, prob, z[name + "_mu_in"], z[name + "_var_in"],
                            z[name + "_mu_out"], z[name + "_var_out"],
                            [int(i) for i in z[name + "_elite"]])
    n = len([k for k in z.files if k.startswith("actor_W")])
    actor = orc.Actor([z["actor_W%d" % i] for i in range(n)], [z["actor_b%d" % i] for i in range(n)],
                      z["actor_log_std"])
    return ens("dyn", True), actor, ens("v", False), ens("vc", False)


def gen_statics(ref):
    rng = np.random.default_rng(0)
    rows = []
    for z in (0.1, 0.2, np.float32(0.2), 0.6, 1.0, 1.0000001, np.nan, np.inf, -np.inf):
        for q in ((0.0, 0.0), (0.93, 0.0), (0.65, 0.66), (np.nan, 0.0), (np.inf, 0.1)):
            for y in (0.0, 3.2, np.float32(3.2), 3.3, -4.0, np.nan):
                o = rng.standard_normal(29).astype(np.float32) * 0.1
                o[0], o[2], o[3], o[-1] = z, q[0], q[1], y
                rows.append(o)
    o = rng.standard_normal(29).astype(np.float32); o[0] = 0.6; o[7] = np.nan; rows.append(o)
    ant = np.array(rows, np.float32)
    act = np.zeros((len(ant), 8), np.float32)
    hcs = rng.standard_normal((64, 17)).astype(np.float32)
    hcs[:8, -1] = [0.0, 0.19999999, 0.2, 0.20000002, -0.2, -0.19999999, np.nan, np.inf]
    hact = np.zeros((64, 6), np.float32)
    with np.errstate(invalid="ignore"):
        np.savez_compressed(os.path.join(OUT, "statics.npz"), ant_obs=ant, hcs_obs=hcs,
                 ant_term=ref.statics.antsafe_term_fn(ant, act, ant),
                 ant_cost=ref.statics.antsafe_c_fn(ant, act, ant),
                 hcs_cost=ref.statics.hcs_cost_f(hcs, hact, hcs),
                 no_done=ref.statics.no_done(hcs, hact, hcs))


def gen_dkl_and_scan(ref):
    rng = np.random.default_rng(1)
    mu = rng.standard_normal((7, 50, 17)).astype(np.float32)
    std = np.exp(rng.normal(-1, 1.5, (7, 50, 17))).astype(np.float32)
    std[0, 0, 0] = 0.0          # log(0) = -inf -> clipped to -100
    std[1, 1, 1] = 1e30         # exp(2*log_std) overflows
    with np.errstate(all="ignore"):
        np.savez_compressed(os.path.join(OUT, "average_dkl.npz"), mu=mu, std=std, out=ref.average_dkl(mu, std))
    x = rng.standard_normal((40, 34)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "discount_cumsum.npz"), x=x,
             out_gae=ref.discount_cumsum(x, 0.99, 0.95, axis=-1),
             out_cgae=ref.discount_cumsum(x, 0.97, 0.5, axis=-1),
             x1=x[0], out1=ref.discount_cumsum(x[0], 0.99, 0.95, axis=-1))


def gen_cpobuffer(ref):
    rng = np.random.default_rng(2)
    O, A, n = 5, 2, 700
    buf = ref.CPOBuffer(1000, 5000, ref_stubs.Space(O), ref_stubs.Space(A))
    buf.initialize({"mu": (A,), "log_std": (A,)}, gamma=0.99, lam=0.95, cost_gamma=0.97, cost_lam=0.5)
    seg = [0, 1, 33, 34, 98, 99, 400, 700]
    obs = rng.standard_normal((n, O)).astype(np.float32)
    act = rng.standard_normal((n, A)).astype(np.float32)
    rew, val, cval, logp = (rng.standard_normal(n).astype(np.float32) for _ in range(4))
    cost = (rng.random(n) < 0.1).astype(np.float32)
    mu = rng.standard_normal((n, A)).astype(np.float32)
    ls = np.full((n, A), -0.5, np.float32)
    lv = rng.standard_normal(len(seg) - 1).astype(np.float32)
    lc = rng.standard_normal(len(seg) - 1).astype(np.float32)
    for s in range(len(seg) - 1):
        for i in range(seg[s], seg[s + 1]):
            buf.store(obs[i], act[i], obs[i], rew[i], val[i], cost[i], cval[i], logp[i],
                      {"mu": mu[i], "log_std": ls[i]}, False, 0)
        buf.finish_path(lv[s:s + 1], lc[s:s + 1])      # 1-element arrays, as CpoSampler passes them
    pre = dict(adv=buf.adv_buf[:n].copy(), ret=buf.ret_buf[:n].copy(), cadv=buf.cadv_buf[:n].copy(),
               cret=buf.cret_buf[:n].copy())
    out, diag = buf.get()
    np.savez_compressed(os.path.join(OUT, "cpobuffer.npz"), seg=np.array(seg), obs=obs, act=act, rew=rew, val=val,
             cost=cost, cval=cval, logp=logp, mu=mu, log_std=ls, last_val=lv, last_cval=lc,
             **{"pre_" + k: v for k, v in pre.items()},
             **{"out%d" % i: a for i, a in enumerate(out)},
             ret_mean=diag["poolr_ret_mean"], cret_mean=diag["poolr_cret_mean"])


def gen_fakeenv(ref):
    for key, (task, O, A) in TASKS.items():
        dyn, actor, v, vc = orc.make_problem(100, O, A, hidden=(32, 32), vf_hidden=(32, 32), a_hidden=(32, 32), task=task)
        obs, act = orc.make_states(101, 256, O, A, dyn)
        pos = np.random.default_rng(102).integers(0, len(dyn.elite_inds), 256)
        with np.errstate(all="ignore"):
            nxt, r, term, info = ref_harness.reference_fakeenv_step(dyn, task, obs, act, pos)
        np.savez_compressed(os.path.join(OUT, "fakeenv_step_%s_tfrestated.npz" % key), task=task, obs=obs, act=act,
                 elite_pos=pos, next_obs=nxt, rew=r, term=term, cost=info["cost"],
                 dkl_path=info["ensemble_dkl_path"], ep_var=info["ensemble_ep_var"],
                 dkl_mean=info["ensemble_dkl_mean"], **pack_problem(dyn, actor, v, vc))


def gen_rollout(ref):
    for key, mode, max_samples in (("hcs", False, 250), ("ant", "uncertainty", None), ("hum", "uncertainty", None)):
        task, O, A = TASKS[key]
        B, T = 48, 9
        dyn, actor, v, vc = orc.make_problem(200, O, A, hidden=(32, 32), vf_hidden=(32, 32), a_hidden=(32, 32), task=task)
        obs, act = orc.make_states(201, B, O, A, dyn)
        noise = orc.TableNoise(202, T, B, A, len(dyn.elite_inds))
        lim = None
        if mode:
            env = orc.OracleFakeEnv(O, A, task, orc.OracleModel(dyn), lambda e, n: np.zeros(n, int))
            lim = float(np.median(env.step(obs, act)[3]["ensemble_dkl_path"]) * 4)
        with np.errstate(all="ignore"):
            out, bdiag, diag, snap = ref_harness.reference_rollout(
                dyn, actor, v, vc, task, obs, noise, T, mode, lim, max_samples=max_samples,
                stop_alive_ratio=0.1, **GAE)
        np.savez_compressed(os.path.join(OUT, "rollout_%s_tfrestated.npz" % key), task=task, start_obs=obs, T=T,
                 mode=str(mode), dkl_lim=-1.0 if lim is None else lim,
                 max_samples=-1 if max_samples is None else max_samples,
                 act_eps=noise.act_eps, elite_pos=noise.elite_pos,
                 **{"out%d" % i: a for i, a in enumerate(out)},
                 **{"snap_" + k: a for k, a in snap.items()},
                 diag_keys=np.array(sorted(diag)), diag_vals=np.array([float(diag[k]) for k in sorted(diag)]),
                 poolm_batch_size=bdiag["poolm_batch_size"], poolm_ret_mean=bdiag["poolm_ret_mean"],
                 poolm_cret_mean=bdiag["poolm_cret_mean"], **pack_problem(dyn, actor, v, vc))


def make_archive_case(seed, n_rows=5000, archive_size=6000, n_ep=7):
    """Epoch labels of an archive filled by consecutive dump_to_archive calls, with unequal epoch
    sizes, a gap in the epoch numbers and an empty tail (epoch -1)."""
    rng = np.random.default_rng(seed)
    sizes = rng.multinomial(n_rows, rng.dirichlet(np.ones(n_ep) * 2))
    labels = np.concatenate([np.full(n, e + (2 if e >= 3 else 0)) for e, n in enumerate(sizes)])
    ep = np.full(archive_size, -1, dtype=int)
    ep[:n_rows] = labels
    kls = np.abs(rng.normal(0.02, 0.03, n_ep))
    kls[0] = 0.0
    return ep, kls


def gen_archive(ref):
    ep, kls = make_archive_case(300)
    buf = ref.CPOBuffer(10, len(ep), ref_stubs.Space(3), ref_stubs.Space(2))
    buf.epoch_archive[:] = ep
    out = {}
    for alpha in (1, 5.0):
        out["btz_alpha%g" % alpha] = buf.boltz_dist(kls, alpha=alpha)
    np.savez_compressed(os.path.join(OUT, "archive_boltz.npz"), epoch_archive=ep, kls=kls,
                        epochs_list=buf.epochs_list, min_ep=buf.min_ep, max_ep=buf.max_ep, **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_stubs.load()
    gen_statics(ref)
    gen_dkl_and_scan(ref)
    gen_cpobuffer(ref)
    gen_fakeenv(ref)
    gen_rollout(ref)
    gen_archive(ref)
    for f in sorted(os.listdir(OUT)):
        print("%-44s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
