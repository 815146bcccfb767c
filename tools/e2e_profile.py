"""Where does the end-to-end (numpy in -> numpy out) pass spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as orc   # synthetic problem generator (no test oracle in tools)

B, T, O, A = 100000, 35, 17, 6
dyn, actor, v, vc = orc.make_problem(0, O, A, hidden=(512, 512))
eng = cb.Engine(0, precision="fp16")
model = cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
policy = cb.B200Policy(eng); policy.load_actor(actor.W, actor.b, actor.log_std); policy.load_values(v, vc)
class S:
    def __init__(s, n): s.shape = (n,)
class Env: observation_space, action_space = S(O), S(A)
fenv = cb.FakeEnv(Env(), "HalfCheetahSafe-v2", model, True, True, False)
pool = cb.ModelBuffer(B, O, A, T, engine=eng)
pool.initialize({"mu": (A,), "log_std": (A,)}, gamma=0.99, lam=0.95, cost_gamma=0.97, cost_lam=0.5)
smp = cb.ModelSampler(T, B, False, logger=object(), seed=7)
smp.initialize(fenv, policy, pool)
obs, _ = orc.make_states(1, B, O, A, dyn)

def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print("  %-28s %7.1f ms" % (name, (t1 - t0) * 1e3)); return t1

for rep in range(3):
    print("pass", rep)
    t = time.perf_counter(); t00 = t
    smp.reset(obs); t = tick("reset (H2D + rollout launch)", t)
    while True:
        _, _, _, info = smp.sample(None)
        if info["alive_ratio"] <= 0.1: break
    t = tick("sample() x n", t)
    smp.finish_all_paths(); t = tick("finish_all_paths", t)
    out, diag = pool.get_device(); t = tick("get_device (stats, compact)", t)
    res = [x.cpu().numpy() for x in out] if rep == 0 else None
    pool.reset(); t = tick("D2H pageable / reset", t)
    print("  total %.1f ms" % ((t - t00) * 1e3))
