"""TEST INFRASTRUCTURE: run in a process started with LD_PRELOAD=libasan (tests/test_emu_asan.py).
Drives the AddressSanitizer build of the emulated library over the whole clip pipeline -- ragged
shapes, both index modes, both rowgather CTA shapes, the heavy paths -- and the operator-level
entry points, with every buffer (inputs, outputs, scene, table, workspace) allocated at exactly
the size the C ABI asks for, so that an out-of-bounds access of any kernel aborts the process."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

import emu
from emu import build
from slr_sfs_b200 import _lib as binding, workloads

L = ctypes.CDLL(build.build(asan=True))
for name, argtypes in binding.SIGNATURES.items():
    fn = getattr(L, name)
    fn.argtypes = argtypes
    fn.restype = binding._OTHER_RESTYPE.get(name, ctypes.c_int)
emu._lib = L
emu.aligned = lambda nbytes, align=256: np.zeros(nbytes, dtype=np.uint8)     # exact size: overruns are visible

for (H, W, C, kind, n) in [(17, 37, 5, "A", 3), (24, 40, 4, "B", 2), (9, 33, 3, "C", 4), (40, 72, 6, "sink", 1)]:
    feat, Z, m = [t.numpy() for t in workloads.scene(H, W, C, "A" if kind == "sink" else kind, seed=1)]
    if kind == "sink":
        ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
        m = np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys])[None].astype(np.float32)
    # index modes: the direct index (insert_kernel, default) and the bin pipeline (bin_fill + expand_kernel)
    for (mode, shape) in (("ldg", "1x4"), ("ldg", "2x2"), ("bins", "1x4"), ("bins", "2x2")):
        os.environ["SLR_GATHER_MODE"] = mode
        os.environ["SLR_GATHER_SHAPE"] = shape
        sc = emu.Scene(feat, Z, m)
        sc.frames(0, 7, 1, n, want_mask=True)
        sc.frames(0, 7, 1, n, table=sc.table(0, 7, 0, 8))
        if kind == "sink":
            print("sink", mode, sc.stats, flush=True)
    rng = np.random.default_rng(0)
    inp = rng.standard_normal((2, 3, H, W)).astype(np.float32)
    flow = rng.uniform(-6, 6, (2, 2, H, W)).astype(np.float32)
    a, b, gf = np.empty_like(inp), np.empty_like(inp), np.empty_like(flow)
    emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(a), 2, 3, H, W, 1, None)
    emu.call("slr_softsplat_grad_input", emu.p(flow), emu.p(inp), emu.p(b), 2, 3, H, W, None)
    emu.call("slr_softsplat_grad_flow", emu.p(inp), emu.p(flow), emu.p(a), emu.p(gf), 2, 3, H, W, None)
    emu.call("slr_maxwarpnorm", emu.p(inp), emu.p(flow), emu.p(a), emu.p(b), 2, 3, H, W, None)
    disp = np.empty((2, H, W), dtype=np.float32)
    emu.call("slr_euler", emu.p(m), -1.0, 9, emu.p(disp), None, H, W, None)
    print("ok", H, W, kind, sc.stats, flush=True)
print("ASAN DRIVE DONE")
