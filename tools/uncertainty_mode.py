"""Rollout throughput in the reference's default 'uncertainty' mode (paths are cut when their
cumulative ensemble KL reaches dkl_lim): device-resident rollout + GAE, with and without the
alive-row compaction (flags=L.ROLLOUT_NO_COMPACT)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as wl

B, T, O, A = 100000, 35, 17, 6
dyn, actor, v, vc = wl.make_problem(0, O, A, hidden=(512, 512))
eng = cb.Engine(0, precision="fp16")
model = cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
pol = cb.B200Policy(eng); pol.load_actor(actor.W, actor.b, actor.log_std); pol.load_values(v, vc)
obs, act = wl.make_states(1, B, O, A, dyn)
cfg = L.EnvCfg(L.TERM_NO_DONE, L.COST_HCS, 0, 1, 1)
# calibrate dkl_lim like cmbpo.py:198-200: a multiple of the one-step mean disagreement
out = eng.fakeenv_step(cfg, obs[:20000], act[:20000], seed=1, step=0)
base = float(out["dkl_path"].mean())
bufs = cb.RolloutBuffers(eng, B, T, O, A); bufs.set_inputs(obs)
for factor in (5.0, 15.0):
    for nc, every in (("0", "1"), ("0", "2"), ("0", "4"), ("1", "4")):
        kw = dict(flags=L.ROLLOUT_NO_COMPACT if nc == "1" else 0, compact_every=int(every))
        for i in range(3):
            bufs.run(cfg, uncertainty_mode=True, dkl_lim=base * factor, seed=i, **kw); bufs.gae(0.99, 0.95, 0.97, 0.5)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(6):
            bufs.run(cfg, uncertainty_mode=True, dkl_lim=base * factor, seed=10 + i, **kw); bufs.gae(0.99, 0.95, 0.97, 0.5)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 6
        n = int(bufs.length.sum().item())
        print("dkl_lim = %4.1f x one-step KL, compaction %s: mean path %.1f steps, %.2f ms per rollout, %.1f M transitions/s"
              % (factor, "off" if nc == "1" else "every %s" % every, n / B, dt * 1e3, n / dt / 1e6))
