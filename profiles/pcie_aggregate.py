"""Aggregate PCIe bandwidth of the box with every GPU copying at once (explains the N-GPU e2e figure):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_aggregate.py
"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 256 << 20
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=10):
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s1):
                host.copy_(d, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d2.copy_(host2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    gbs = reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return gbs, float(t.item())


for kind in ("d2h", "h2d", "both"):
    run(kind, 2)
    mine, total = run(kind)
    if rank == 0:
        print("%s: rank0 %.1f GB/s per direction, all %d ranks %.1f GB/s" % (kind, mine, world, total), flush=True)
if world > 1:
    dist.destroy_process_group()
