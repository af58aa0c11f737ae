#!/usr/bin/env python
"""Why does the e2e leg (201 MB fp32 x 64 channels per frame leaving the device) not scale with the GPU count?
Per-GPU pinned D2H bandwidth alone (ranks take turns) against all ranks at once, plus the CPU affinity of
every rank.  torchrun --nproc-per-node N profiles/probe_d2h.py > gpurun_out/probe_d2h.json (rank 0 prints)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 28                                   # 1 GiB of fp32
src = torch.empty(n, dtype=torch.float32, device=dev).normal_()
dst = torch.empty(n, dtype=torch.float32).pin_memory()
back = torch.empty(n, dtype=torch.float32).pin_memory()


def bw(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return reps * n * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


res = {"rank": rank, "affinity": sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]}
alone_d2h = alone_h2d = None
for r in range(world):
    barrier()
    if r == rank:
        alone_d2h = bw(lambda: dst.copy_(src, non_blocking=True))
        alone_h2d = bw(lambda: src.copy_(back, non_blocking=True))
barrier()
together_d2h = bw(lambda: dst.copy_(src, non_blocking=True))
barrier()
together_h2d = bw(lambda: src.copy_(back, non_blocking=True))
barrier()
res.update(d2h_alone_GBps=alone_d2h, h2d_alone_GBps=alone_h2d, d2h_all_ranks_GBps=together_d2h, h2d_all_ranks_GBps=together_h2d)
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    print(json.dumps({"world": world, "ranks": out,
                      "sum_d2h_all_ranks_GBps": sum(r["d2h_all_ranks_GBps"] for r in out),
                      "sum_d2h_alone_GBps": sum(r["d2h_alone_GBps"] for r in out)}, indent=1))
if world > 1:
    dist.destroy_process_group()
