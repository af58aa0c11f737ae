#!/usr/bin/env python
"""Small runs of every new kernel path for compute-sanitizer (memcheck / initcheck):
    compute-sanitizer --tool memcheck  python profiles/sanitize_drive.py
    compute-sanitizer --tool initcheck python profiles/sanitize_drive.py
The direct index never clears its lists: initcheck is the proof that the gather only reads cells that were written."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__

__graft_entry__.build()
import oracle
import slr_sfs_b200 as pkg
from slr_sfs_b200 import training_block, workloads


def rel(a, b):
    s = float(np.sqrt(np.mean(b.astype(np.float64) ** 2)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), s)))


dev = torch.device("cuda")
H, W, C, N = 72, 104, 9, 7
feat, Z, m = workloads.scene(H, W, C, "A", seed=3)
js = pkg.JointSplat(feat.to(dev), Z.to(dev), m.to(dev))
out, aux, mask, nnz = js.frames(0, N - 1, 0, N, want_aux=True, want_mask=True, want_nnz=True)
torch.cuda.synchronize()
for t in (0, 3, N - 1):
    assert rel(out[t:t + 1].cpu().numpy(), oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, t, N - 1))) <= 1e-4
print("clip ok")
# convergent flow: deep lists, excess pairs, and (one frame) the whole-batch fallback
ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
for flow in (np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys])[None].astype(np.float32),
             np.stack([-(xs - W / 2) * 0.45, -(ys - H / 2) * 0.2])[None].astype(np.float32)):
    js = pkg.JointSplat(feat.to(dev), Z.to(dev), torch.from_numpy(flow).to(dev))
    for (t0, n) in ((1, 1), (1, 2)):
        got = js.frames(0, 2, t0, n).cpu().numpy()
        assert rel(got[:1], oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), flow, (0, t0, 2))) <= 1e-4
print("convergent ok")
# operator-level splat through the gather, and the training block
pkg.softsplat.GATHER_MIN_ELEMENTS = 0
x = torch.randn(2, 5, H, W, device=dev)
fl = (torch.rand(2, 2, H, W, device=dev) * 8 - 4)
fl[:, :, :10] = 0
y = pkg.FunctionSoftsplat(x, fl, None, "summation")
assert rel(y.cpu().numpy(), oracle.softsplat_sum(x.cpu().numpy(), fl.cpu().numpy())) <= 1e-4
print("splat ok")
t = [torch.randn(2, 6, H, W, device=dev, requires_grad=True) for _ in range(2)] + \
    [torch.randn(2, 1, H, W, device=dev, requires_grad=True) for _ in range(2)] + \
    [(torch.rand(2, 2, H, W, device=dev) * 6 - 3).requires_grad_(True) for _ in range(2)]
gen = training_block.joint_block_training(*t, torch.tensor([0.3, 0.8], device=dev))
gen.sum().backward()
torch.cuda.synchronize()
print("training block ok")
print("SANITIZE DRIVE DONE")
