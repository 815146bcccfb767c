"""Debug aid: cProfile of ModelSampler.finish_all_paths (fused mode)."""
import os, sys, time, cProfile, pstats
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "e2e_profile.py")).read().split("for rep in range")[0])
for rep in range(3):
    smp.reset(obs)
    while True:
        _, _, _, info = smp.sample(None)
        if info["alive_ratio"] <= 0.1: break
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter(); smp.finish_all_paths(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pr.disable()
    print("finish_all_paths %.2f ms" % (dt * 1e3))
    if rep == 2: pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
    pool.get_device(); pool.reset()
