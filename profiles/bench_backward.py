#!/usr/bin/env python
"""the backward kernels at speed.  Times slr_softsplat_grad_input / slr_softsplat_grad_flow
(and the forward) against the reference's own CUDA kernels (oracle/_ref, built by oracle/build.py) on the same
GPU, at the training shape (B=2 per GPU, 65 channels, 256x256) and at the inference shape.  CUDA events,
20 repetitions after 3 warm-ups; inputs larger than nothing here fit L2 at 256^2 -- flagged in the output.

    python profiles/bench_backward.py > gpurun_out/backward.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0      # us


def main():
    import __graft_entry__
    __graft_entry__.build()
    from slr_sfs_b200 import _lib
    from oracle import refgpu
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    for shape in [(2, 65, 256, 256), (1, 65, 768, 1024)]:
        B, C, H, W = shape
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, C, H, W, generator=g).cuda()
        flow = (torch.rand(B, 2, H, W, generator=g) * 6 - 3).cuda()
        gout = torch.randn(B, C, H, W, generator=g).cuda()
        gin, gflow, out = torch.empty_like(x), torch.empty_like(flow), torch.empty_like(x)
        s = _lib.current_stream(x.device)
        plane = 4.0 * B * H * W
        ours = {
            "forward": timeit(lambda: _lib.call("slr_softsplat_sum_fwd", _lib.ptr(x), _lib.ptr(flow), _lib.ptr(out), B, C, H, W, 1, s)),
            "grad_input": timeit(lambda: _lib.call("slr_softsplat_grad_input", _lib.ptr(flow), _lib.ptr(gout), _lib.ptr(gin), B, C, H, W, s)),
            "grad_flow": timeit(lambda: _lib.call("slr_softsplat_grad_flow", _lib.ptr(x), _lib.ptr(flow), _lib.ptr(gout), _lib.ptr(gflow), B, C, H, W, s)),
        }
        ref = {}
        if refgpu.available() and shape in refgpu.baked_backward_shapes():
            ref["grad_input"] = timeit(lambda: refgpu.softsplat_backward(x, flow, gout, True, False))
            ref["grad_flow"] = timeit(lambda: refgpu.softsplat_backward(x, flow, gout, False, True))
        if refgpu.available() and shape in refgpu.baked_shapes():
            ref["forward"] = timeit(lambda: refgpu.softsplat_sum(x, flow))
        alg = {"forward": plane * (2 * C + 2), "grad_input": plane * (2 * C + 2), "grad_flow": plane * (2 * C + 4)}
        for k, us in ours.items():
            rows.append({"shape": list(shape), "kernel": k, "ours_us": us, "reference_us": ref.get(k),
                         "speedup": None if k not in ref else ref[k] / us,
                         "algorithmic_GBps": alg[k] / us / 1e3, "frac_of_hbm_peak": alg[k] / us / 1e3 / peak,
                         "note": "reference timings include its zero-filled output allocation (new_zeros), as its backward() does; "
                                 "working set %.0f MB %s L2" % (alg[k] / 1e6, "fits" if alg[k] < 100e6 else "exceeds")})
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
