#!/usr/bin/env python
"""Training forward + backward of the joint block (animating_softmax_splating.py:584-692) at the training
shape: the fused producer + splat (slr_sfs_b200.training_block) against the reference's own expression on the
Level-0 drop-in operators with torch autograd (what install_as_reference_modules() gives with zero source
changes).  CUDA events, inputs resident.   python profiles/bench_training_block.py > gpurun_out/training_block.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__

__graft_entry__.build()
import slr_sfs_b200 as pkg
from slr_sfs_b200 import training_block


def inputs(B, C, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).cuda().requires_grad_(True)
    t = dict(start_fs=mk(B, C, H, W), end_fs=mk(B, C, H, W), Z_f=mk(B, 1, H, W), Z_p=mk(B, 1, H, W),
             flow_f=(torch.rand(B, 2, H, W, generator=g) * 10 - 5).cuda().requires_grad_(True),
             flow_p=(torch.rand(B, 2, H, W, generator=g) * 10 - 5).cuda().requires_grad_(True))
    alpha = torch.rand(B, generator=g).cuda()
    return t, alpha


def level0(t, alpha):
    B, _, H, W = t["start_fs"].shape
    a4 = alpha.view(B, 1, 1, 1)
    zf = torch.clamp(t["Z_f"] - t["Z_f"].max(), min=-20.0, max=20.0)
    zp = torch.clamp(t["Z_p"] - t["Z_p"].max(), min=-20.0, max=20.0)
    ten_f = torch.cat([t["start_fs"] * zf.exp() * a4, zf.exp() * a4], 1)
    ten_p = torch.cat([t["end_fs"] * zp.exp() * (1 - a4), zp.exp() * (1 - a4)], 1)
    splat = pkg.softsplat.ModuleSoftsplat("summation")
    ones = t["start_fs"].new_ones(B, 1, H, W)
    acc = splat(tenInput=ten_f, tenFlow=t["flow_f"], tenMetric=ones) + splat(tenInput=ten_p, tenFlow=t["flow_p"], tenMetric=ones)
    return acc[:, :-1] / torch.clamp(acc[:, -1:], min=1e-8)


def fused(t, alpha):
    return training_block.joint_block_training(t["start_fs"], t["end_fs"], t["Z_f"], t["Z_p"], t["flow_f"], t["flow_p"], alpha)


def timed(fn, t, alpha, g, reps=20):
    for _ in range(3):
        for v in t.values():
            v.grad = None
        fn(t, alpha).backward(g)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd = bwd = 0.0
    for _ in range(reps):
        for v in t.values():
            v.grad = None
        e[0].record()
        out = fn(t, alpha)
        e[1].record()
        out.backward(g)
        e[2].record()
        torch.cuda.synchronize()
        fwd += e[0].elapsed_time(e[1])
        bwd += e[1].elapsed_time(e[2])
    return fwd / reps, bwd / reps


if __name__ == "__main__":
    for (B, C, H, W) in [(2, 64, 256, 256), (16, 64, 256, 256), (1, 64, 768, 1024)]:
        t, alpha = inputs(B, C, H, W)
        g = torch.randn(B, C, H, W, device="cuda")
        f0, b0 = timed(level0, t, alpha, g)
        f1, b1 = timed(fused, t, alpha, g)
        torch.cuda.reset_peak_memory_stats()
        fused(t, alpha).backward(g)
        m1 = torch.cuda.max_memory_allocated()
        torch.cuda.reset_peak_memory_stats()
        level0(t, alpha).backward(g)
        m0 = torch.cuda.max_memory_allocated()
        print(json.dumps({"shape": [B, C, H, W], "level0_ms": {"forward": round(f0, 4), "backward": round(b0, 4)},
                          "fused_ms": {"forward": round(f1, 4), "backward": round(b1, 4)},
                          "speedup_fwd_bwd": round((f0 + b0) / (f1 + b1), 3),
                          "peak_MB": {"level0": round(m0 / 1e6, 1), "fused": round(m1 / 1e6, 1)}}))
