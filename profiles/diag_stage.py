#!/usr/bin/env python
"""How many (tile, frame) pairs of the benchmark scene the staged gather leaves to the L1 gather, per batch."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__
__graft_entry__.build()
import slr_sfs_b200 as pkg
from slr_sfs_b200 import _lib, workloads

H, W, C, N = 768, 1024, 64, 60
motion = sys.argv[1] if len(sys.argv) > 1 else "A"
dev = torch.device("cuda")
feat, Z, m = (t.to(dev) for t in workloads.scene(H, W, C, motion, seed=0))
js = pkg.JointSplat(feat, Z, m)
js.pipeline = False
rows = []
for t0 in range(0, N, 12):
    out = js.frames(0, N - 1, t0, 12)
    torch.cuda.synchronize()
    st = js._shared_state()
    ws = st["ws"][st["turn"]]
    stats = (ctypes.c_uint32 * 6)()
    _lib.call("slr_clip_stats_host", _lib.ptr(ws), ws.numel() * 4, H, W, 12, stats, _lib.current_stream(dev))
    rows.append({"t0": t0, "flagged": stats[0], "full_heavy": stats[1], "excess_pairs": stats[2], "l1_fallback": stats[4],
                 "tile_frames": stats[5], "fallback_frac": stats[4] / stats[5]})
print(json.dumps(rows, indent=1))
