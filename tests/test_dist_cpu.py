"""CPU, world_size 2 over gloo: the N>1 host logic -- contiguous sharding and the two Allreduce steps
of mpi_statistics_scalar (utilities/mpi_tools.py:71-92) -- gives the single-process statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cmbpo_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cmbpo_b200
    from cmbpo_b200.dist import shard_bounds, make_reduce_fn, combine_pass1, combine_pass2
    rng = np.random.default_rng(0)                       # same data on every rank
    B, T = 1001, 9
    length = rng.integers(0, T, B)
    adv, cadv, ret, cret = (rng.standard_normal((B, T)).astype(np.float32) * 2 + 0.5 for _ in range(4))
    lo, hi = shard_bounds(B, rank, world)
    m = (np.arange(T)[None] < length[lo:hi, None])
    reduce_fn = make_reduce_fn()
    assert reduce_fn is not None
    # pass 1: local float64 sums of this rank's paths, all-reduced (what cmbpo_adv_stats_pass1 + NCCL do)
    s = torch.tensor([m.sum(), adv[lo:hi][m].sum(dtype=np.float64), cadv[lo:hi][m].sum(dtype=np.float64),
                      ret[lo:hi][m].sum(dtype=np.float64), cret[lo:hi][m].sum(dtype=np.float64)], dtype=torch.float64)
    reduce_fn(s)
    st = combine_pass1(s.numpy())
    d = adv[lo:hi][m] - st["adv_mean"]
    ss = torch.tensor([np.sum((d * d).astype(np.float64))], dtype=torch.float64)
    reduce_fn(ss)
    st["adv_std"] = combine_pass2(float(ss[0]), st["n"])
    q.put((rank, lo, hi, st["n"], float(st["adv_mean"]), float(st["adv_std"]), float(st["cadv_mean"])))
    dist.destroy_process_group()


def test_two_rank_statistics_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the path range without overlap
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1001
    rng = np.random.default_rng(0)
    B, T = 1001, 9
    length = rng.integers(0, T, B)
    adv, cadv = (rng.standard_normal((B, T)).astype(np.float32) * 2 + 0.5 for _ in range(2))
    m = np.arange(T)[None] < length[:, None]
    mean, std = orc.stats_scalar(adv[m])
    cmean, _ = orc.stats_scalar(cadv[m])
    for r in res:
        assert r[3] == m.sum()
        assert abs(r[4] - mean) <= 4e-7 * max(1, abs(mean))
        assert abs(r[5] - std) <= 4e-7 * std
        assert abs(r[6] - cmean) <= 4e-7 * max(1, abs(cmean))
    assert res[0][3:] == res[1][3:]            # every rank derives identical statistics


def test_shard_bounds_cover():
    from cmbpo_b200.dist import shard_bounds
    for n in (0, 1, 7, 100000, 1000003):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
