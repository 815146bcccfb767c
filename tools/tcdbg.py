"""Debug aid: per-CTA protocol timing of the tcgen05 ensemble kernel (CMBPO_TC_DEBUG=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cmbpo_b200 as cb
from cmbpo_b200 import _lib as L
from cmbpo_b200 import workload as orc   # synthetic problem generator (no test oracle in tools)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dyn, actor, v, vc = orc.make_problem(0, 17, 6, hidden=(512, 512))
eng = cb.Engine(0, precision="fp16")
L.check(eng.lib.cmbpo_ctx_set_debug(eng.h, int(os.environ.get("CMBPO_TC_DEBUG", "0")), int(os.environ.get("CMBPO_TC_TRACE_ONLY", "0"))))
model = cb.B200PE.from_arrays(eng, L.NET_DYN, dyn)
obs, act = orc.make_states(1, N, 17, 6, dyn)
x = eng.to_device(np.concatenate([obs, act], -1))
for i in range(3):
    m, v_ = model.predict_ensemble_device(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    m, v_ = model.predict_ensemble_device(x)
torch.cuda.synchronize()
print("predict_ensemble %d rows: %.3f ms" % (N, (time.perf_counter() - t0) / 5 * 1e3))
