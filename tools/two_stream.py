"""Experiment: does splitting a rollout batch over two CUDA streams (two engines) recover the tail /
imbalance / launch-gap time of the step-wise rollout?  Compares 1 x 100k against 2 x 50k."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as orc   # synthetic problem generator (no test oracle in tools)

B, T, O, A = 100000, 35, 17, 6
dyn, actor, v, vc = orc.make_problem(0, O, A, hidden=(512, 512))
obs, _ = orc.make_states(1, B, O, A, dyn)
cfg = L.EnvCfg(L.TERM_NO_DONE, L.COST_HCS, 0, 1, 1)

def make(nB, stream):
    with torch.cuda.stream(stream):
        eng = cb.Engine(0, precision="fp16")
        cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
        pol = cb.B200Policy(eng); pol.load_actor(actor.W, actor.b, actor.log_std); pol.load_values(v, vc)
        bufs = cb.RolloutBuffers(eng, nB, T, O, A)
    return eng, bufs

s0 = torch.cuda.current_stream()
e1, b1 = make(B, s0); b1.set_inputs(obs)
def run_one():
    b1.run(cfg, seed=3)
for _ in range(3): run_one()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): run_one()
torch.cuda.synchronize(); t1 = time.perf_counter()
print("1 x 100k : %.2f ms per rollout" % ((t1 - t0) / 5 * 1e3))

sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ea, ba = make(B // 2, sa); eb, bb = make(B // 2, sb)
with torch.cuda.stream(sa): ba.set_inputs(obs[:B // 2])
with torch.cuda.stream(sb): bb.set_inputs(obs[B // 2:])
torch.cuda.synchronize()
def run_two():
    ba.run(cfg, seed=3, path_id_base=0)
    bb.run(cfg, seed=3, path_id_base=B // 2)
for _ in range(3): run_two()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): run_two()
torch.cuda.synchronize(); t1 = time.perf_counter()
print("2 x 50k  : %.2f ms per rollout pair (two streams)" % ((t1 - t0) / 5 * 1e3))
