"""Fused (two launches per step) against the default four-launch rollout step: ms per rollout and per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as wl

B, T, O, A = 100000, 35, 17, 6
dyn, actor, v, vc = wl.make_problem(0, O, A, hidden=(512, 512))
eng = cb.Engine(0, precision="fp16")
model = cb.B200PE.from_oracle_ensemble(eng, L.NET_DYN, dyn)
pol = cb.B200Policy(eng); pol.load_actor(actor.W, actor.b, actor.log_std); pol.load_values(v, vc)
obs, act = wl.make_states(1, B, O, A, dyn)
cfg = L.EnvCfg(L.TERM_NO_DONE, L.COST_HCS, 0, 1, 1)
bufs = cb.RolloutBuffers(eng, B, T, O, A); bufs.set_inputs(obs)
masks = [0]
for flags, name in ((L.ROLLOUT_FUSE, "fused"), (0, "stepwise")):
    for m in (masks if flags else [0]):
        for i in range(2):
            bufs.run(cfg, seed=i, flags=flags)
        eng.profile(True); eng.profile_read(0, True); eng.profile_read(3, True); eng.profile_read(2, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3):
            bufs.run(cfg, seed=10 + i, flags=flags)
        e1.record(); torch.cuda.synchronize()
        d_ms, d_n = eng.profile_read(0, True); p_ms, p_n = eng.profile_read(3, True); s_ms, s_n = eng.profile_read(2, True)
        eng.profile(False)
        print("%s skip=%3d: rollout %.2f ms | dyn %.1f us/launch  policy %.1f us/launch  rows %.1f us/launch"
              % (name, m, e0.elapsed_time(e1) / 3, 1e3 * d_ms / max(d_n, 1), 1e3 * p_ms / max(p_n, 1), 1e3 * s_ms / max(s_n, 1)), flush=True)
