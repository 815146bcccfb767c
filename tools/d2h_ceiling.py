"""What the BOX delivers when N GPUs copy device -> page-locked host memory at the same time (plain cudaMemcpyAsync of
512 MB per rank, nothing of this repo involved): the ceiling of any numpy-out end-to-end number at N GPUs.
Run under torchrun like bench.py; rank 0 prints one JSON line."""
import json, os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dst.fill_(1)
res = {}
for label, sync in (("alone_rank0", False), ("all_ranks_together", True)):
    if label == "alone_rank0" and rank != 0:
        if world > 1:
            dist.barrier()
        continue
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if sync and world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    gbps = reps * n / (time.perf_counter() - t0) / 1e9
    if label == "alone_rank0":
        res[label] = gbps
        if world > 1:
            dist.barrier()
    else:
        t = torch.tensor([gbps], device="cuda", dtype=torch.float64)
        if world > 1:
            lst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            per = [float(x.item()) for x in lst]
        else:
            per = [gbps]
        res["per_rank_GBps"] = per
        res["aggregate_GBps"] = sum(per)
if rank == 0:
    res.update(n_gpus=world, bytes_per_copy=n, what="concurrent D2H into page-locked host memory, torch copy_, 10 copies per rank")
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
