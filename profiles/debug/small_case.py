"""Tiny gather run for compute-sanitizer:  compute-sanitizer python profiles/debug/small_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import slr_sfs_b200 as pkg
from slr_sfs_b200 import workloads

for (H, W, C, n) in [(20, 20, 6, 1), (40, 70, 8, 5), (96, 128, 16, 12)]:
    feat, Z, m = workloads.scene(H, W, C, "A", seed=3)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    out = js.frames(0, 11, 0, n)
    torch.cuda.synchronize()
    print("ok", H, W, C, n, float(out.abs().sum()))
