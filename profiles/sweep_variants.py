#!/usr/bin/env python
"""Tuning sweep of the clip pipeline's run-time knobs on one B200, in ONE process (the library
reads SLR_GATHER_SHAPE / SLR_EXPAND_CLAIM per call, the batch is a JointSplat attribute):

    SLR_GATHER_SHAPE   CTA shape of rowgather_kernel, frames x row pairs   (1x4, 2x4, 4x4, 2x2, 4x1)
    SLR_EXPAND_CLAIM   how expand_kernel claims list slots                 (atomic, store)
    SLR_SMEM_CARVEOUT  shared-memory carve-out all clip kernels ask for, %  (-1 = driver default)
    SLR_GATHER_PAD_SMEM unused dynamic shared memory per gather CTA (caps its CTAs per SM under a carve-out)
    SLR_SIDE_PRIORITY  priority of the side stream (table / bins / expand)  (0, -1 = high)
    batch              frames per expand / gather launch

One knob at a time against the defaults, then the best value of each combined.  The timed step is
bench.py's `value` step (configs[1]: 768x1024x64, N = 60, motion A, inputs resident, CUDA events).
Run under gpurun:  python profiles/sweep_variants.py > gpurun_out/sweep_variants.jsonl
The winner is printed last as {"best": {...}}; per-kernel times (single stream) for the default and
the winner are in "kernels_ms_per_frame"."""
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
DEFAULT = {"shape": os.environ.get("SLR_GATHER_SHAPE", "1x4"), "claim": os.environ.get("SLR_EXPAND_CLAIM", "atomic"),
           "carveout": int(os.environ.get("SLR_SMEM_CARVEOUT", "-1")), "priority": int(os.environ.get("SLR_SIDE_PRIORITY", "0")),
           "pad": int(os.environ.get("SLR_GATHER_PAD_SMEM", "0")), "main_priority": 0,
           "batch": pkg.JointSplat.batch}


def apply(v):
    os.environ["SLR_GATHER_SHAPE"] = v["shape"]
    os.environ["SLR_EXPAND_CLAIM"] = v["claim"]
    os.environ["SLR_SMEM_CARVEOUT"] = str(v["carveout"])
    os.environ["SLR_GATHER_PAD_SMEM"] = str(v["pad"])
    if os.environ.get("SLR_SIDE_PRIORITY") != str(v["priority"]):
        os.environ["SLR_SIDE_PRIORITY"] = str(v["priority"])
        torch.cuda.synchronize()
        pkg.JointSplat._shared.clear()          # a new side stream with the new priority


def main():
    feat, Z, m = workloads.scene(H, W, C, os.environ.get("SWEEP_MOTION", "A"), seed=0)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    bufs = [torch.empty(48, C, H, W, device="cuda") for _ in range(2)]

    high = torch.cuda.Stream(priority=-1)

    def main_stream(v):
        # main_priority -1: the caller's stream (the gather) is a high-priority one, so that the side
        # stream's CTAs only fill what the gather leaves free
        return torch.cuda.stream(high if v.get("main_priority") else torch.cuda.current_stream())

    def clip(v, pipeline=True):
        js = pkg.JointSplat(feat, Z, m, inputs_event=False)
        js.batch, js.pipeline = v["batch"], pipeline
        js.prepare_clip(0, N - 1)
        for i, b0 in enumerate(range(0, N, 48)):
            nb = min(48, N - b0)
            js.frames(0, N - 1, b0, nb, out=bufs[i & 1][:nb])

    def measure(v, steps=10, reps=2, pipeline=True):
        apply(v)
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
        apply(v)
        clip(v, False)
        _lib.kernel_timing(True)
        for _ in range(3):
            clip(v, False)
        t = _lib.kernel_timing(False)
        return {k: round(ms / (3 * N), 5) for k, (ms, n) in t.items()}

    def emit(row):
        print(json.dumps(row), flush=True)

    stages = os.environ.get("SWEEP_STAGES", "knobs,overlap,room,main").split(",")
    base = measure(DEFAULT)
    emit(dict(DEFAULT, frames_per_s=base, what="default", kernels_ms_per_frame=kernels(DEFAULT)))
    best, top = dict(DEFAULT), base
    if "knobs" in stages:
        for knob, values in (("shape", ["2x4", "4x4", "2x2", "4x1"]), ("claim", ["store"]), ("batch", [6, 8, 10, 16, 20]),
                             ("carveout", [25, 50, 75]), ("priority", [-1])):
            knob_top = base
            for val in values:
                v = dict(DEFAULT, **{knob: val})
                fps = measure(v)
                emit(dict(v, frames_per_s=fps, what=knob))
                if fps > knob_top * 1.005:
                    knob_top, best[knob] = fps, val
        top = measure(best)
        emit(dict(best, frames_per_s=top, what="best of each"))
    if "overlap" in stages:      # the overlap knobs interact: carve-out x priority on top of the best shape / claim / batch
        for carve in (-1, 25, 50, 75):
            for prio in (0, -1):
                v = dict(best, carveout=carve, priority=prio)
                if v == best:
                    continue
                fps = measure(v)
                emit(dict(v, frames_per_s=fps, what="overlap"))
                if fps > top * 1.005:
                    top, best = fps, v
    if "room" in stages:         # cap the gather at 3 CTAs per SM so that a side-stream CTA fits beside it
        for carve, pad in ((50, 33 << 10), (50, 28 << 10), (75, 40 << 10), (75, 33 << 10)):
            for prio in (0, -1):
                v = dict(best, carveout=carve, pad=pad, priority=prio)
                fps = measure(v)
                emit(dict(v, frames_per_s=fps, what="room for the side stream"))
                if fps > top * 1.005:
                    top, best = fps, v
    if "main" in stages:         # the gather's stream at high priority, for the default and the multi-frame CTA shapes
        for shape in ("1x4", "2x2", "4x1"):
            for mp in (0, -1):
                for batch in (12, 16):
                    v = dict(DEFAULT, shape=shape, main_priority=mp, batch=batch)
                    fps = measure(v)
                    emit(dict(v, frames_per_s=fps, what="main stream priority"))
                    if fps > top * 1.005:
                        top, best = fps, v
    fps = measure(best, steps=20)
    again = measure(DEFAULT, steps=20)
    emit(dict(best, frames_per_s=fps, what="combined", kernels_ms_per_frame=kernels(best)))
    emit(dict(DEFAULT, frames_per_s=again, what="default again"))
    emit({"best": best if fps > again * 1.005 else DEFAULT, "gain": fps / again})


if __name__ == "__main__":
    main()
