"""Debug aid: where finish_all_paths (fused mode) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(__file__), "e2e_profile.py")).read().split("for rep in range")[0])
def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print("  %-28s %7.2f ms" % (name, (t1 - t0) * 1e3)); return t1
for it in range(3):
    smp.reset(obs)
    while True:
        _, _, _, info = smp.sample(max_samples=None)
        if info["alive_ratio"] <= 0: break
    torch.cuda.synchronize(); print("pass", it); t = time.perf_counter()
    bufs = pool.bufs
    if smp._alive_now > 0:
        bufs.truncate(stop_step=smp._n_episodes - 1); smp._alive_now = 0
    t = tick("truncate", t)
    pool.finish_all_device(); t = tick("gae", t)
    d = bufs.diagnostics(); t = tick("diagnostics", t)
    a = bufs.path_return.cpu().numpy(); b = bufs.path_cost.cpu().numpy(); t = tick("path arrays", t)
    pool.ptr = int(bufs.length.max().item()); t = tick("length max", t)
    out, diag = pool.get_device(); t = tick("get_device", t)
    pool.reset(); t = tick("reset", t)
