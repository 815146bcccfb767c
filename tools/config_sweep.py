"""Device-resident rollout + GAE throughput of the three task shapes of BASELINE.json's configs
(HalfCheetahSafe 17/6, AntSafe 29/8, HumanoidSafe 47/17), 100 k start states, maxroll 35, fp16."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as wl

CFG = {"hcs": ("HalfCheetahSafe-v2", 17, 6, L.TERM_NO_DONE, L.COST_HCS),
       "ant": ("AntSafe-v2", 29, 8, L.TERM_ANTSAFE, L.COST_ANTSAFE),
       "hum": ("HumanoidSafe-v2", 47, 17, L.TERM_NO_DONE, L.COST_ZERO)}
B, T = 100000, 35
for key, (task, O, A, term, cost) in CFG.items():
    dyn, actor, v, vc = wl.make_problem(0, O, A, hidden=(512, 512), task=task)
    eng = cb.Engine(0, precision="fp16")
    cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
    pol = cb.B200Policy(eng); pol.load_actor(actor.W, actor.b, actor.log_std); pol.load_values(v, vc)
    obs, _ = wl.make_states(1, B, O, A, dyn)
    bufs = cb.RolloutBuffers(eng, B, T, O, A)
    bufs.set_inputs(obs)
    cfg = L.EnvCfg(term, cost, 0, 1, 1)
    def one(seed):
        bufs.run(cfg, seed=seed)
        bufs.gae(0.99, 0.95, 0.97, 0.5)
    for i in range(4): one(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    counts = []
    for i in range(8):
        one(10 + i); counts.append(bufs.length.sum())       # no host sync inside the timed loop
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = sum(int(c.item()) for c in counts)
    print("%s  O=%d A=%d: %.1f M transitions/s (%.2f ms per rollout, mean path length %.1f)"
          % (key, O, A, n / dt / 1e6, dt / 8 * 1e3, n / 8 / B))
    eng.close(); del bufs
