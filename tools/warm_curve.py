"""Per-rollout duration over consecutive rollouts from a cold start (clock / power ramp)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as wl
B, T, O, A = 100000, 35, 17, 6
dyn, actor, v, vc = wl.make_problem(0, O, A, hidden=(512, 512))
eng = cb.Engine(0, precision="fp16")
cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
pol = cb.B200Policy(eng); pol.load_actor(actor.W, actor.b, actor.log_std); pol.load_values(v, vc)
obs, _ = wl.make_states(1, B, O, A, dyn)
bufs = cb.RolloutBuffers(eng, B, T, O, A); bufs.set_inputs(obs)
cfg = L.EnvCfg(L.TERM_NO_DONE, L.COST_HCS, 0, 1, 1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
ev[0].record()
for i in range(40):
    bufs.run(cfg, seed=i); ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
print("rollout ms:", " ".join("%.1f" % x for x in ms))
if len(sys.argv) > 1:
    eng.profile(True)
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev2[0].record()
    for i in range(10):
        bufs.run(cfg, seed=i); ev2[i + 1].record()
    torch.cuda.synchronize()
    print("with profiling on:", " ".join("%.1f" % ev2[i].elapsed_time(ev2[i + 1]) for i in range(10)))
