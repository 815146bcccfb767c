#!/usr/bin/env python
"""bench.py -- model-rollout transitions/sec (+ GAE ms per 1M steps) of the CMBPO hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU: the oracle port of the reference

One "step" = one pass of the hot path over one batch of synthetic start states:
H-step model rollout (policy -> PENS ensemble -> FakeEnv statics -> ModelBuffer write-out), GAE +
cost-GAE scans, advantage statistics / normalisation.  Workload at every N: BASELINE.json
configs[1] -- HalfCheetahSafe (obs 17, act 6), 7x(512,512) swish ensemble, 100 000 start states per
GPU, maxroll 35 (34 stored steps), deterministic-mean transitions, Philox noise; weak scaling
(start states sharded across ranks, no data-path collective, one tiny all-reduce of statistics).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv and int(os.environ.get("RANK", "0")) == 0:
    # the reference arm runs on rank 0 alone with ALL host threads; torchrun pins OMP_NUM_THREADS=1
    # for its children, which must be undone before numpy loads its BLAS
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_k] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK, OBS, ACT = "HalfCheetahSafe-v2", 17, 6
HIDDEN, E = (512, 512), 7
MAXROLL = 35
GAE = dict(gamma=0.99, lam=0.95, cost_gamma=0.97, cost_lam=0.5)
METRIC, UNIT = "model-rollout transitions/sec", "transitions/s"
WORKLOAD = ("BASELINE configs[1]: HalfCheetahSafe H-step model rollout + GAE/cost-GAE + advantage normalisation")


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def host_threads():
    """Threads the numpy BLAS behind the oracle port actually uses (what `cores` reports)."""
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def flop_per_transition(O=OBS, A=ACT, h=HIDDEN[0]):
    """dynamics + actor + V,VC, un-padded dims."""
    D, Din = O + 1, O + A
    dyn = 2 * E * (Din * h + h * h + h * 2 * D)
    actor = 2 * (O * 128 + 128 * 128 + 128 * A)
    vvc = 2 * 3 * 2 * (O * 128 + 128 * 128 + 128)
    return dyn, actor, vvc


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sus=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def ncu_traffic():
    """DRAM bytes per launch of K1 from the committed `ncu --set full` summary (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r1_k1_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks
    line), through NVML in a thread (initialised before the region starts: spawning nvidia-smi takes
    longer than a short timed region on an 8-GPU box); falls back to `nvidia-smi -lms`."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag = None, None, [], False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; map through CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.idx < len(ids) and ids[self.idx].isdigit():
                    phys = int(ids[self.idx])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, b in bits.items() if any(r & b for _, _, r in self.samples))
            pmax = max(p for _, p, _ in self.samples)
            load = [s for s, p, _ in self.samples if p > 0.5 * pmax] or [s for s, _, _ in self.samples]
            return {"sm_mhz": float(np.median(load)), "sm_max_mhz": self.max_sm, "power_w_max": float(pmax),
                    "samples": len(self.samples), "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (numpy), bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_rollout_sample(B, seed=0):
    """One reference-port pass (reset -> sample x 34 -> finish_all_paths -> get) on B start states."""
    from oracle import cmbpo_oracle as orc
    dyn, actor, v, vc = orc.make_problem(seed, OBS, ACT, hidden=HIDDEN, task=TASK)
    obs, _ = orc.make_states(seed + 1, B, OBS, ACT, dyn)
    noise = orc.TableNoise(seed + 2, MAXROLL, B, ACT, len(dyn.elite_inds))
    t0 = time.perf_counter()
    out, bdiag, diag, _ = orc.run_rollout(dyn, actor, v, vc, TASK, obs, noise, MAXROLL,
                                          gamma=GAE["gamma"], lam=GAE["lam"],
                                          cgamma=GAE["cost_gamma"], clam=GAE["cost_lam"])
    dt = time.perf_counter() - t0
    return len(out[0]), dt


def cpu_gae_ms_per_1m():
    from oracle import cmbpo_oracle as orc
    rng = np.random.default_rng(0)
    B, T = 8192, 34
    r, v, c, cv = (rng.standard_normal((B, T)).astype(np.float32) for _ in range(4))
    lv, lc = (rng.standard_normal(B).astype(np.float32) for _ in range(2))
    t0 = time.perf_counter()
    orc.gae_path(r, v, c, cv, lv, lc, 0.99, 0.95, 0.97, 0.5)
    return (time.perf_counter() - t0) * 1e3 / (B * T / 1e6)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_threads()
    B = args.cpu_batch
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_rollout_sample(min(B, 200))
    times, n_tr = [], 0
    for s in range(args.steps):
        n, dt = cpu_rollout_sample(B, seed=s)
        times.append(dt); n_tr += n
    total = sum(times)
    value = n_tr / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "impl_note": "CPU port of the reference (numpy restatement of the TF "
                               "graphs; TF 1.14 is not installable), bounded sample of the workload per step",
                   "start_states_per_step": B, "maxroll": MAXROLL, "stored_steps": MAXROLL - 1, "obs": OBS,
                   "act": ACT, "ensemble": "7x(512,512) swish, 5 elites"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d start states x %d steps per pass, %d passes" % (B, MAXROLL - 1, args.steps),
                         "gae_ms_per_1M_steps": cpu_gae_ms_per_1m()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import cmbpo_b200 as cb
    from cmbpo_b200 import _lib as L
    from cmbpo_b200 import workload as wl      # synthetic weights / start states (no oracle on this arm)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = cb.Engine(local_rank, precision=args.precision)
    t = torch

    # weights: generated on rank 0, broadcast once over NCCL (NVLink), then uploaded from device
    dyn, actor, v, vc = wl.make_problem(0, OBS, ACT, hidden=HIDDEN, task=TASK)

    def bcast(a):
        x = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        if world > 1:
            if rank != 0:
                x.zero_()
            dist.broadcast(x, 0)
        return x

    def ens_dev(en):
        return dict(W=[bcast(w) for w in en.W], b=[bcast(b) for b in en.b], mu_in=bcast(en.mu_in),
                    var_in=bcast(en.var_in), mu_out=bcast(en.mu_out), var_out=bcast(en.var_out))

    dev_nets = {L.NET_DYN: ens_dev(dyn), L.NET_V: ens_dev(v), L.NET_VC: ens_dev(vc)}
    dev_actor = ([bcast(w) for w in actor.W], [bcast(b) for b in actor.b], bcast(actor.log_std))

    def load_engine(e):
        for which, en, prob in ((L.NET_DYN, dyn, True), (L.NET_V, v, False), (L.NET_VC, vc, False)):
            d = dev_nets[which]
            e.set_network(which, d["W"], d["b"], en.acts, d["mu_in"], d["var_in"], d["mu_out"],
                          d["var_out"], prob, en.elite_inds)
        e.set_actor(*dev_actor)

    load_engine(eng)

    B, T = args.batch, MAXROLL
    obs_host, _ = wl.make_states(100 + rank, B, OBS, ACT, dyn)
    obs_pinned = torch.from_numpy(obs_host).pin_memory()
    start_dev = obs_pinned.to(dev)
    bufs = cb.RolloutBuffers(eng, B, T, OBS, ACT)
    env_cfg = L.EnvCfg(L.TERM_NO_DONE, L.COST_HCS, 0, 1, 1)
    path_base = rank * B

    def reduce_fn(x):
        if world > 1:
            dist.all_reduce(x)

    def hot_path(step, start):
        """rollout -> GAE -> statistics -> normalise, everything resident in HBM."""
        bufs.start_obs = start
        bufs.run(env_cfg, seed=1234 + step, path_id_base=path_base, precision=args.precision)
        bufs.gae(GAE["gamma"], GAE["lam"], GAE["cost_gamma"], GAE["cost_lam"])
        st = eng.adv_statistics(bufs.adv, bufs.cadv, bufs.ret, bufs.cret, B, T, 1, B, bufs.length,
                                reduce_fn if world > 1 else None)
        eng.adv_normalise(bufs.adv, bufs.cadv, B, T, 1, B, bufs.length, st)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        hot_path(-1 - w, start_dev)
    barrier()

    # ---- timed region (device-resident inputs) ----
    eng.profile(True)
    eng.profile_read(L_PROF_DYN, True); eng.profile_read(L_PROF_GAE, True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    n_tr = 0
    for s in range(args.steps):
        st = hot_path(s, start_dev)
        n_tr += st["n"]          # with N > 1 the statistics are all-reduced: already the whole-job count
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    dyn_ms, dyn_n = eng.profile_read(L_PROF_DYN, True)
    gae_ms, gae_n = eng.profile_read(L_PROF_GAE, True)
    step_ms, step_n = eng.profile_read(2, True)
    pol_ms, pol_n = eng.profile_read(3, True)
    eng.profile(False)

    # ---- end to end through the public API: host start states in, host sample list out ----
    class _Space:
        def __init__(self, n):
            self.shape = (n,)

    class ShapeEnv:
        observation_space, action_space = _Space(OBS), _Space(ACT)

    # Two sampler / buffer pairs, each on its own engine (= its own CUDA stream), used alternately as a
    # caller would to keep the GPU busy: reset() only queues the rollout, so batch i+1 rolls out while
    # the host walks batch i through sample() / finish_all_paths() / get_async() and while the
    # device->host copy of batch i runs on a side stream.  Every batch's H2D (start states) and D2H
    # (sample list) are inside the timed region.
    def make_pair(k):
        stream = torch.cuda.current_stream(dev) if k == 0 else torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            e = eng if k == 0 else cb.Engine(local_rank, precision=args.precision)
            if k:
                load_engine(e)
            policy = cb.B200Policy(e)
            policy.attach_loaded(actor.log_std)
            fenv = cb.FakeEnv(ShapeEnv(), TASK, cb.B200PE.view(e, L.NET_DYN), True, True, False)
            pool = cb.ModelBuffer(B, OBS, ACT, T, engine=e)
            pool.initialize({"mu": (ACT,), "log_std": (ACT,)}, **GAE)
            pool.reduce_fn = reduce_fn if world > 1 else None
            smp = cb.ModelSampler(T, B, False, logger=object(), seed=7 + k)
            smp.path_id_base = path_base
            smp.initialize(fenv, policy, pool)
        return smp, pool, stream

    pairs = [make_pair(0), make_pair(1)]
    torch.cuda.synchronize()

    def e2e_reset(k):
        with torch.cuda.stream(pairs[k][2]):
            pairs[k][0].reset(obs_host)              # H2D of the start states, rollout queued

    def e2e_collect(k):
        smp, pool, stream = pairs[k]
        with torch.cuda.stream(stream):
            while True:
                _, _, _, info = smp.sample(None)    # the first call waits for this pair's rollout
                if info["alive_ratio"] <= 0.1:
                    break
            smp.finish_all_paths()
            return pool.get_async()

    def e2e_run(n_batches):
        n, d2h, handles = 0, 0, []
        e2e_reset(0)
        for i in range(n_batches):
            if i + 1 < n_batches:
                e2e_reset((i + 1) & 1)
            handles.append(e2e_collect(i & 1))
            if len(handles) == 2:                    # at most two sample lists in flight
                out, _ = handles.pop(0).result()
                n += len(out[0])
        for h in handles:
            out, _ = h.result()
            n += len(out[0])
        # bytes that crossed PCIe: a stride-0 axis (log_std, one row broadcast) is not copied
        d2h = sum(int(np.prod([m for m, st in zip(a.shape, a.strides) if st != 0] or [1])) * a.itemsize for a in out)
        return n, d2h

    e2e_run(4)              # warm-up: page-locked buffer sets of both pairs get allocated
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 10))
    e2e_n, d2h = e2e_run(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks / totals ----
    vec = torch.tensor([ms, e2e_s, dyn_ms / max(dyn_n, 1), gae_ms / max(gae_n, 1)], device=dev, dtype=torch.float64)
    cnt = torch.tensor([n_tr, e2e_n, launches], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        cnt[0] /= world          # n_tr was global on every rank (see the timed loop)
    ms, e2e_s, dyn_launch_ms, gae_launch_ms = (float(x) for x in vec)
    n_tr, e2e_n, launches = (float(x) for x in cnt)

    if rank == 0:
        pk = peaks()
        fdyn, fact, fvvc = flop_per_transition()
        rows_per_launch = B                      # every dynamics launch processes all B rows of one step
        achieved_tf = fdyn * rows_per_launch / (dyn_launch_ms * 1e-3) / 1e12 if dyn_launch_ms > 0 else 0.0
        tensor_path = args.precision != "fp32"
        peak_tf = pk["tf_sus"]
        gae_steps = B * (T - 1)
        gae_gbs = 32.0 * gae_steps / (gae_launch_ms * 1e-3) / 1e9 if gae_launch_ms > 0 else 0.0
        # CPU baseline: bounded sample of the same workload through the oracle port
        # CPU baseline: rank 0 at N=1 only (under torchrun OMP_NUM_THREADS is pinned to 1)
        skip_cpu = world > 1 or args.cpu_batch <= 0        # --cpu-batch 0: developer runs only
        cpu_n, cpu_dt = (0, 1.0) if skip_cpu else cpu_rollout_sample(args.cpu_batch)
        if not skip_cpu and cpu_dt < 6.0:                         # scale the bounded sample to ~12 s of CPU work
            scaled = int(min(20000, args.cpu_batch * 12.0 / max(cpu_dt, 1e-3)))
            cpu_n, cpu_dt = cpu_rollout_sample(scaled)
            args.cpu_batch = scaled
        cores = host_threads()
        line = {
            "metric": METRIC, "value": n_tr / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16": "f16 (tcgen05 kind::f16, f32 accumulate)",
                      "bf16": "bf16 (tcgen05 kind::f16, f32 accumulate)",
                      }[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "start_states_per_gpu": B, "maxroll": T, "stored_steps": T - 1, "obs": OBS, "act": ACT,
                       "ensemble": "7x(512,512) swish, 5 elites", "policy": "tanh 128-128 + 2x(3x swish 128-128-1)",
                       "mode": "deterministic-mean, Philox noise, dkl_lim=inf", "precision": args.precision,
                       "l2": "rollout buffers (%.0f MB/GPU) exceed the 126 MB L2; no explicit flush" %
                             (B * T * ((2 * OBS + 2 * ACT + 11) * 4 + 1) / 1e6),
                       "flop_per_transition": fdyn + fact + fvvc,
                       "peak_note": "tensor peak = %s bf16 GEMM, sustained (burst %.1f)" % (pk["src"], pk["tf_burst"])},
            "roofline": {"bound": "tensor", "kernel": "dynamics-ensemble GEMM chain (K1)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if tensor_path else achieved_tf / peak_tf,
                         "traffic": ncu_traffic(), "launch_ms": dyn_launch_ms, "rows_per_launch": rows_per_launch,
                         "flop_per_row": fdyn, "share_of_step": dyn_launch_ms * (T - 1) / (ms / args.steps)},
            "gae": {"ms_per_1M_steps": gae_launch_ms / (gae_steps / 1e6), "achieved_GBps": gae_gbs,
                    "peak_GBps": pk["hbm"], "frac": gae_gbs / pk["hbm"], "bytes_per_step": 32,
                    "steps_per_launch": gae_steps, "scan": "strict float64 sequential (bit-exact)"},
            "cpu_baseline": ({"value": cpu_n / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
                              "sample": "%d start states x %d steps (one pass, %.1f s)" % (args.cpu_batch, T - 1, cpu_dt)}
                             if not skip_cpu else None),
            "e2e": {"value": e2e_n / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(obs_host.nbytes), "d2h_bytes_per_step": int(d2h),
                    "api": "ModelSampler.reset/sample/finish_all_paths + ModelBuffer.get_async().result(), numpy in / numpy out, two sampler+buffer pairs alternating (batch i+1 rolls out while batch i is collected and copied)"},
            "breakdown_ms_per_step": {"dynamics_gemm_chain": dyn_ms / args.steps, "policy_pass": pol_ms / args.steps,
                                      "row_kernel": step_ms / args.steps, "gae": gae_ms / args.steps,
                                      "note": "rank 0, CUDA events around each launch"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


L_PROF_DYN, L_PROF_GAE = 0, 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp32", "fp16", "bf16"])
    ap.add_argument("--batch", type=int, default=100000, help="start states per GPU")
    ap.add_argument("--cpu-batch", type=int, default=500, help="start states of the bounded CPU sample")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print at C level while the job runs
    # (NCCL's version banner) goes to stderr; the real stdout is restored for the final print
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        os.execv(sys.executable, cmd)
    run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
