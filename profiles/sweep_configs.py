#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[4]: the 2-layer (67-channel) joint block and the
resolution sweep, timed with CUDA events (inputs resident, outputs stay on the device).
Run under gpurun:  python profiles/sweep_configs.py > gpurun_out/sweep_configs.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__

__graft_entry__.build()
import slr_sfs_b200 as pkg
from slr_sfs_b200 import workloads

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def run(H, W, C=64, N=60, two_layer=False, batch=12, reps=3):
    feat, Z, m = workloads.scene(H, W, C, "A", seed=0)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    tail = None
    clamp = (0.0, 1.0)
    if two_layer:       # use_alpha0_as_blending_weight layout: 64 + (a_f e^A, e^A) + e^Z = 67 channels
        a_f, a_bg = (t.cuda() for t in workloads.two_layer_extras(H, W))
        A = torch.sigmoid(a_f) / torch.clamp(torch.sigmoid(a_f) + a_bg, min=1e-8)
        tail = torch.cat([a_f * A.exp(), A.exp()], 1).contiguous()
        clamp = (1.0 / 600.0, 599.0 / 600.0)
    nbuf = min(N, 2 * batch)
    out = torch.empty(nbuf, C, H, W, device="cuda")

    def clip():
        js = pkg.JointSplat(feat, Z, m, tail=tail)
        js.batch = batch
        js.prepare_clip(0, N - 1)
        for b0 in range(0, N, nbuf):
            nb = min(nbuf, N - b0)
            js.frames(0, N - 1, b0, nb, out=out[:nb], want_aux=two_layer, want_mask=two_layer, alpha_clamp=clamp)

    for _ in range(2):
        clip()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        clip()
    e1.record()
    torch.cuda.synchronize()
    ms_frame = e0.elapsed_time(e1) / (reps * N)
    P = H * W
    n_a, n_w, n_extra = (2, 3, 1) if two_layer else (0, 1, 0)
    whole = 4.0 * P * (4 * C + 3 + n_a + 2 * n_w + n_extra)          # SURVEY 8d whole-path bytes per frame (two passes)
    floor = 4.0 * P * (2 * C + 3 + n_a + n_extra)                    # one pass: inputs once + outputs once
    return {"H": H, "W": W, "C": C, "two_layer": two_layer, "batch": batch, "frames_per_s": 1000.0 / ms_frame,
            "us_per_frame": 1000.0 * ms_frame, "one_pass_floor_MB_per_frame": floor / 1e6,
            "frac_one_pass_floor": floor / ms_frame / 1e6 / PEAK,
            "two_pass_MB_per_frame": whole / 1e6, "frac_two_pass": whole / ms_frame / 1e6 / PEAK,
            "workspace_MB_per_frame_of_batch": pkg._lib.load().slr_clip_workspace_bytes(H, W, 1) / 1e6}


if __name__ == "__main__":
    rows = [run(768, 1024), run(768, 1024, two_layer=True)]
    for (H, W, b) in [(256, 256, 12), (512, 512, 12), (1024, 1024, 12), (1536, 2048, 12), (1536, 2048, 3)]:
        rows.append(run(H, W, batch=b))
    if "--batches" in sys.argv:      # small frames are launch-bound: fewer, larger batches (kMaxFrames = 64)
        for (H, W) in [(256, 256), (512, 512)]:
            for b in (20, 30, 60):
                rows.append(run(H, W, batch=b))
    for r in rows:
        print(json.dumps(r))
