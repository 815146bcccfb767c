"""Soak test: many rollouts of random shapes / modes / precisions in one process; checks that nothing
traps (a protocol dead-lock in the tcgen05 kernel traps after ~4.5 s) and every stored value is finite."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as wl

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
TASKS = [("HalfCheetahSafe-v2", 17, 6, L.TERM_NO_DONE, L.COST_HCS), ("AntSafe-v2", 29, 8, L.TERM_ANTSAFE, L.COST_ANTSAFE),
         ("HumanoidSafe-v2", 47, 17, L.TERM_NO_DONE, L.COST_ZERO)]
t0 = time.time()
n_roll = 0
for it in range(12):
    task, O, A, term, cost = TASKS[it % 3]
    hidden = [(512, 512), (256, 256), (128, 128), (200, 200), (300, 300), (64, 64)][int(rng.integers(6))]
    E = int(rng.choice([3, 5, 7]))
    dyn, actor, v, vc = wl.make_problem(it, O, A, hidden=hidden, num_nets=E, num_elites=max(1, E - 2), task=task)
    prec = ["fp16", "bf16"][it % 2]
    eng = cb.Engine(0, precision=prec)
    cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
    pol = cb.B200Policy(eng); pol.load_actor(actor.W, actor.b, actor.log_std); pol.load_values(v, vc)
    cfg = L.EnvCfg(term, cost, 0, 1, 1)
    for rep in range(10):
        B = int(rng.choice([1, 7, 127, 128, 129, 1000, 4097, 20000, 50000]))
        T = int(rng.choice([2, 3, 9, 35]))
        obs, act = wl.make_states(100 * it + rep, B, O, A, dyn)
        bufs = cb.RolloutBuffers(eng, B, T, O, A); bufs.set_inputs(obs)
        unc = bool(rng.integers(2))
        lim = 0.0
        if unc:
            out = eng.fakeenv_step(cfg, obs[:min(B, 2000)], act[:min(B, 2000)], seed=1, step=0)
            lim = float(out["dkl_path"].mean()) * float(rng.choice([0.5, 3.0, 20.0]))
        print(task, hidden, E, prec, "B=%d T=%d unc=%s lim=%.3g" % (B, T, unc, lim), flush=True)
        bufs.run(cfg, uncertainty_mode=unc, dkl_lim=lim, seed=rep)
        bufs.gae(0.99, 0.95, 0.97, 0.5)
        eng.synchronize()
        ln = bufs.length
        m = (torch.arange(T, device=ln.device)[:, None] < ln[None, :])
        for name in ("rew", "val", "cval", "adv", "cadv", "logp"):
            x = getattr(bufs, name)
            assert bool(torch.isfinite(x[m]).all()), (task, hidden, E, prec, B, T, unc, name)
        assert int(ln.max()) <= T - 1 and int(ln.min()) >= 0
        n_roll += 1
        del bufs
    eng.close()
print("soak ok: %d rollouts, %.1f s" % (n_roll, time.time() - t0))
