#!/usr/bin/env python
"""The operator-level summation splat (FunctionSoftsplat 'summation' forward) by fp32 atomics
(slr_softsplat_sum_fwd) and through the gather pipeline (slr_softsplat_sum_fwd_gather), with the reference's
own kernel on the same GPU where oracle/_ref holds it for the shape.  CUDA events, inputs resident.
    python profiles/bench_forward_op.py > gpurun_out/forward_op.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__

__graft_entry__.build()
import slr_sfs_b200 as pkg
from slr_sfs_b200 import workloads


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if __name__ == "__main__":
    try:
        from oracle import refgpu
        ref = refgpu if refgpu.available() else None
    except Exception:
        ref = None
    for (B, C, H, W) in [(1, 65, 768, 1024), (1, 65, 512, 512), (2, 65, 256, 256), (1, 65, 1536, 2048)]:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B, C, H, W, generator=g).cuda()
        # displacement of a mid-clip frame of the benchmark scene (smooth, half of the pixels static)
        _, _, m = workloads.scene(H, W, 4, "A", seed=0)
        from slr_sfs_b200.euler_integration_manipulator import euler_integration
        flow = euler_integration(m.cuda(), 30)[0].repeat(B, 1, 1, 1).contiguous()
        row = {"shape": [B, C, H, W]}
        for name, thresh in (("atomics_ms", 1 << 62), ("gather_ms", 0)):
            pkg.softsplat.GATHER_MIN_ELEMENTS = thresh
            row[name] = round(timed(lambda: pkg.FunctionSoftsplat(x, flow, None, "summation")), 4)
        if ref is not None and (B, C, H, W) in ref.baked_shapes():
            row["reference_kernel_ms"] = round(timed(lambda: ref.softsplat_sum(x, flow)), 4)
        row["speedup_vs_atomics"] = round(row["atomics_ms"] / row["gather_ms"], 3)
        print(json.dumps(row))
