#!/usr/bin/env python
"""Tuning sweep of the clip pipeline's knobs on one B200, in ONE process:

    SLR_GATHER_SHAPE   CTA shape of rowgather_kernel, frames x row pairs (2x2, 1x4; read per call)
    batch              frames per expand / gather launch (a JointSplat attribute)
    main_priority      run the caller's stream (the gather) at high priority

The timed step is bench.py's `value` step (configs[1]: 768x1024x64, N = 60, motion A, inputs
resident, CUDA events).  Run under gpurun:
    python profiles/sweep_variants.py > gpurun_out/sweep_variants.jsonl

profiles/r01/sweep_variants*.jsonl were written by the wider version of this script at commit
187de93, when the library still had the knobs that were measured and then removed (expand claim
mode, shared-memory carve-out, gather shared-memory padding, side-stream priority, CTA shapes
2x4 / 4x4 / 4x1): `git show 187de93:profiles/sweep_variants.py`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__

__graft_entry__.build()
import slr_sfs_b200 as pkg
from slr_sfs_b200 import _lib, workloads

H, W, C, N = 768, 1024, 64, 60
DEFAULT = {"shape": os.environ.get("SLR_GATHER_SHAPE", "2x2"), "batch": pkg.JointSplat.batch, "main_priority": 0}


def main():
    feat, Z, m = workloads.scene(H, W, C, os.environ.get("SWEEP_MOTION", "A"), seed=0)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    bufs = [torch.empty(48, C, H, W, device="cuda") for _ in range(2)]
    high = torch.cuda.Stream(priority=-1)

    def main_stream(v):
        return torch.cuda.stream(high if v.get("main_priority") else torch.cuda.current_stream())

    def clip(v, pipeline=True):
        js = pkg.JointSplat(feat, Z, m, inputs_event=False)
        js.batch, js.pipeline = v["batch"], pipeline
        js.prepare_clip(0, N - 1)
        for i, b0 in enumerate(range(0, N, 48)):
            nb = min(48, N - b0)
            js.frames(0, N - 1, b0, nb, out=bufs[i & 1][:nb])

    def measure(v, steps=10, reps=2, pipeline=True):
        os.environ["SLR_GATHER_SHAPE"] = v["shape"]
        best = None
        with main_stream(v):
            for _ in range(2):
                clip(v, pipeline)
            for _ in range(reps):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    clip(v, pipeline)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / (steps * N)
                best = ms if best is None else min(best, ms)
        return 1000.0 / best

    def kernels(v):
        os.environ["SLR_GATHER_SHAPE"] = v["shape"]
        clip(v, False)
        _lib.kernel_timing(True)
        for _ in range(3):
            clip(v, False)
        t = _lib.kernel_timing(False)
        return {k: round(ms / (3 * N), 5) for k, (ms, n) in t.items()}

    def emit(row):
        print(json.dumps(row), flush=True)

    base = measure(DEFAULT)
    emit(dict(DEFAULT, frames_per_s=base, what="default", kernels_ms_per_frame=kernels(DEFAULT)))
    best, top = dict(DEFAULT), base
    for shape in ("2x2", "1x4"):
        for batch in (8, 12, 16):
            for mp in (0, -1):
                v = dict(shape=shape, batch=batch, main_priority=mp)
                if v == DEFAULT:
                    continue
                fps = measure(v)
                emit(dict(v, frames_per_s=fps, what="variant"))
                if fps > top * 1.005:
                    top, best = fps, v
    fps = measure(best, steps=20)
    again = measure(DEFAULT, steps=20)
    emit(dict(best, frames_per_s=fps, what="best", kernels_ms_per_frame=kernels(best)))
    emit(dict(DEFAULT, frames_per_s=again, what="default again"))
    emit({"best": best if fps > again * 1.005 else DEFAULT, "gain": fps / again})


if __name__ == "__main__":
    main()
