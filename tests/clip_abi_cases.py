"""One C-ABI call sequence, two back ends: the CUDA library on device buffers (tests/test_gpu_*)
and the CPU emulation of the same sources on numpy buffers (tests/test_emu_*).  Keeping the body
shared means the GPU test's own logic has already run (on the CPU) before it ever sees a GPU."""
import ctypes

import numpy as np

import oracle
from conftest import rel_err

TOL = 1e-4


class EmuBackend:
    def __init__(self):
        import emu
        self.emu = emu
        self.lib = emu.lib()
        self.stream = None

    def call(self, name, *args):
        self.emu.call(name, *args)

    def dev(self, a):
        return np.ascontiguousarray(a, dtype=np.float32)

    def empty(self, *shape):
        return np.full(shape, np.nan, dtype=np.float32)

    def scratch(self, nbytes):
        return self.emu.aligned(nbytes)

    def ptr(self, a):
        return None if a is None else ctypes.c_void_p(a.ctypes.data)

    def host(self, a):
        return np.asarray(a)

    def sync(self):
        pass


class CudaBackend:
    def __init__(self):
        import torch
        from slr_sfs_b200 import _lib
        self.torch, self._lib = torch, _lib
        self.lib = _lib.load()
        self.stream = _lib.current_stream(torch.device("cuda"))

    def call(self, name, *args):
        self._lib.call(name, *args)

    def dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()

    def empty(self, *shape):
        return self.torch.full(shape, float("nan"), device="cuda")

    def scratch(self, nbytes):
        return self.torch.empty((nbytes + 3) // 4, dtype=self.torch.float32, device="cuda")

    def ptr(self, a):
        return None if a is None else ctypes.c_void_p(a.data_ptr())

    def host(self, a):
        return a.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()


def clip_table_bin_stats(be, H=40, W=72, C=6, start=1, end=12, t0=2, n_table=9, batches=((2, 4), (6, 3), (9, 2), (4, 1))):
    """slr_clip_table once, then slr_clip_bin + expand + gather + heavy per batch, against the oracle
    and against the per-batch slr_clip_plan path; slr_clip_stats_host on a convergent flow."""
    from slr_sfs_b200 import workloads
    feat, Z, m = [t.numpy() for t in workloads.scene(H, W, C, "A", seed=17)]
    d_feat, d_Z, d_m = be.dev(feat), be.dev(Z), be.dev(m)
    lib, s = be.lib, be.stream
    zmax = be.empty(1)
    scene = be.scratch(lib.slr_scene_bytes(C, 0, H, W))
    be.call("slr_reduce_max", be.ptr(d_Z), Z.size, be.ptr(zmax), s)
    be.call("slr_scene_prep", be.ptr(d_feat), be.ptr(d_Z), be.ptr(zmax), None, 0, be.ptr(scene), C, H, W, s)
    tb_bytes = lib.slr_clip_table_bytes(H, W, n_table)
    table = be.scratch(tb_bytes)
    be.call("slr_clip_table", be.ptr(d_m), H, W, start, end, t0, n_table, be.ptr(table), tb_bytes, s)
    for (b0, n) in batches:
        ws_bytes = lib.slr_clip_workspace_bytes(H, W, n)
        ws = be.scratch(ws_bytes)
        out, mask = be.empty(n, C, H, W), be.empty(n, 1, H, W)
        args = (C, 0, H, W, start, end, b0, n, 0.0, 1.0)
        be.call("slr_clip_bin", be.ptr(table), tb_bytes, H, W, n_table, b0 - t0, n, be.ptr(ws), ws_bytes, s)
        be.call("slr_clip_expand", be.ptr(scene), be.ptr(d_m), *args, be.ptr(ws), ws_bytes, s)
        for entry in ("slr_clip_gather", "slr_clip_heavy"):
            be.call(entry, be.ptr(scene), be.ptr(d_m), *args, be.ptr(out), None, be.ptr(mask), None, be.ptr(ws), ws_bytes, s)
        ref = be.empty(n, C, H, W)
        ws2 = be.scratch(ws_bytes)
        be.call("slr_clip_frames", be.ptr(scene), be.ptr(d_m), *args, be.ptr(ref), None, None, None, be.ptr(ws2), ws_bytes, s)
        be.sync()
        got, ref = be.host(out), be.host(ref)
        assert rel_err(got, ref) <= 1e-5, (b0, n)              # same chains, continued instead of restarted
        for i in range(n):
            want = oracle.joint_splat_baseline(feat, Z, m, (start, b0 + i, end))
            assert rel_err(got[i:i + 1], want) <= TOL, (b0, i)
            assert np.all(got[i:i + 1][want == 0.0] == 0.0)
            covered = np.abs(want).sum(1, keepdims=True) != 0
            assert np.mean(be.host(mask)[i:i + 1].astype(bool) != covered) < 1e-2
    # argument errors are reported, not crashes
    for bad in (dict(f0=n_table, n=1), dict(f0=0, n=65)):
        try:
            be.call("slr_clip_bin", be.ptr(table), tb_bytes, H, W, n_table, bad["f0"], bad["n"], be.ptr(ws), ws_bytes, s)
        except Exception as exc:
            assert "slr_clip_bin" in str(exc)
        else:
            raise AssertionError("slr_clip_bin accepted %r" % (bad,))

    # a one-step sink: the tile that receives everything is flagged, its pairs beyond the list
    # depth are counted, and with one frame in the batch the excess list (2 * P entries) overflows
    # (bin pipeline: whole-tile reductions; direct index: the batch is redone by scatter + divide)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    sink = np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys])[None].astype(np.float32)
    d_sink = be.dev(sink)
    ws_bytes = lib.slr_clip_workspace_bytes(H, W, 1)
    ws = be.scratch(ws_bytes)
    out = be.empty(1, C, H, W)
    be.call("slr_clip_frames", be.ptr(scene), be.ptr(d_sink), C, 0, H, W, 0, 2, 1, 1, 0.0, 1.0,
            be.ptr(out), None, None, None, be.ptr(ws), ws_bytes, s)
    stats = (ctypes.c_uint32 * 6)()
    be.call("slr_clip_stats_host", be.ptr(ws), ws_bytes, H, W, 1, stats, s)
    assert stats[0] >= 1 and stats[2] > stats[3] == 2 * H * W and stats[5] == 5 * 3
    # whole-tile heavy tiles exist with the bin pipeline only (the direct index redoes the batch by scatter + divide)
    import os
    assert (stats[1] >= 1) == (os.environ.get("SLR_GATHER_MODE") in ("bins", "staged"))
    want = oracle.joint_splat_baseline(feat, Z, sink, (0, 1, 2))
    assert rel_err(be.host(out), want) <= TOL
