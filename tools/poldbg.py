"""Debug aid: per-CTA protocol counters / event trace of the grouped (policy) tcgen05 kernel (CMBPO_TC_DEBUG=2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dyn, actor, v, vc = orc.make_problem(0, 17, 6, hidden=(64, 64))
eng = cb.Engine(0, precision="fp16")
L.check(eng.lib.cmbpo_ctx_set_debug(eng.h, int(os.environ.get("CMBPO_TC_DEBUG", "0")), int(os.environ.get("CMBPO_TC_TRACE_ONLY", "0"))))
pol = cb.B200Policy(eng)
pol.load_actor(actor.W, actor.b, actor.log_std)
pol.load_values(v, vc)
obs, act = orc.make_states(1, N, 17, 6, dyn)
x = eng.to_device(obs)
for i in range(3):
    eng.policy_act(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    eng.policy_act(x)
torch.cuda.synchronize()
print("policy_act %d rows: %.3f ms" % (N, (time.perf_counter() - t0) / 5 * 1e3))
