"""Debug aid: host-side timeline of the pipelined end-to-end loop of bench.py (two sampler pairs)."""
import os, sys, time
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "e2e_profile.py")).read().split("for rep in range")[0])
def make_pair(seed):
    pool = cb.ModelBuffer(B, O, A, T, engine=eng)
    pool.initialize({"mu": (A,), "log_std": (A,)}, gamma=0.99, lam=0.95, cost_gamma=0.97, cost_lam=0.5)
    smp = cb.ModelSampler(T, B, False, logger=object(), seed=seed)
    smp.initialize(fenv, policy, pool)
    return smp, pool
pairs = [make_pair(7), make_pair(8)]
log = []
def stamp(name, t0):
    t1 = time.perf_counter(); log.append((name, (t1 - t0) * 1e3)); return t1
def collect(k):
    smp, pool = pairs[k]
    t = time.perf_counter()
    smp.sample(None); t = stamp("first sample (waits rollout)", t)
    while True:
        _, _, _, info = smp.sample(None)
        if info["alive_ratio"] <= 0.1: break
    t = stamp("other samples", t)
    smp.finish_all_paths(); t = stamp("finish_all_paths", t)
    h = pool.get_async(); t = stamp("get_async", t)
    return h
def run(n):
    handles = []
    t = time.perf_counter()
    pairs[0][0].reset(obs); t = stamp("reset", t)
    for i in range(n):
        if i + 1 < n:
            t = time.perf_counter(); pairs[(i + 1) & 1][0].reset(obs); stamp("reset", t)
        handles.append(collect(i & 1))
        if len(handles) == 2:
            t = time.perf_counter(); handles.pop(0).result(); stamp("result", t)
    for h in handles:
        t = time.perf_counter(); h.result(); stamp("result", t)
run(4); log.clear()
t0 = time.perf_counter(); run(8); total = (time.perf_counter() - t0) * 1e3
from collections import defaultdict
agg = defaultdict(list)
for k, v in log: agg[k].append(v)
print("total %.1f ms for 8 batches = %.2f ms per batch" % (total, total / 8))
for k, v in agg.items(): print("  %-30s n=%2d mean %.2f ms  max %.2f" % (k, len(v), sum(v) / len(v), max(v)))
