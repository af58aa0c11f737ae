"""TEST INFRASTRUCTURE: the CUDA sources of slr-sfs_b200/csrc compiled for the CPU (see build.py
and include/cuda_runtime.h) and driven through the same C ABI with numpy buffers."""
import contextlib
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    """The emulation build of libslr_splat, bound with the argument types the product binding
    (slr_sfs_b200/_lib.py: SIGNATURES) declares for the real one."""
    global _lib
    if _lib is None:
        from slr_sfs_b200 import _lib as binding
        L = ctypes.CDLL(_build.build())
        for name, argtypes in binding.SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = binding._OTHER_RESTYPE.get(name, ctypes.c_int)
        _lib = L
    return _lib


def _bind(path):
    from slr_sfs_b200 import _lib as binding
    L = ctypes.CDLL(path)
    for name, argtypes in binding.SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = binding._OTHER_RESTYPE.get(name, ctypes.c_int)
    return L


@contextlib.contextmanager
def variant(tag, defines):
    """Run the enclosed calls against a compile-time variant of the library (build.build_variant)."""
    global _lib
    lib()
    saved, _lib = _lib, _bind(_build.build_variant(tag, defines))
    try:
        yield
    finally:
        _lib = saved


def shfl_up_calls():
    """Lane-level __shfl_up_sync calls executed so far by the current library (emu::kShflUp = 2)."""
    fn = lib().emu_collective_calls
    fn.restype = ctypes.c_ulonglong
    return int(fn(2))


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError("%s -> %d: %s" % (name, rc, L.slr_last_error_string().decode()))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def aligned(nbytes, align=256):
    """Zero-initialised byte buffer whose data pointer is `align`-aligned (cudaMalloc gives 256)."""
    raw = np.zeros(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


class Scene:
    """What synthesis.JointSplat does on the device, on host buffers: Z.max(), scene buffer,
    then batches of frames through the clip entry points."""

    def __init__(self, feat, Z, motion, tail=None, z_mode="max"):
        self.feat, self.Z, self.motion = f32(feat), f32(Z), f32(motion)
        self.tail = None if tail is None else f32(tail)
        self.n_tail = 0 if tail is None else self.tail.shape[1]
        _, self.C, self.H, self.W = self.feat.shape
        self.zsub = None
        if z_mode == "max":
            self.zsub = np.zeros(1, dtype=np.float32)
            call("slr_reduce_max", p(self.Z), self.Z.size, p(self.zsub), None)
        n = lib().slr_scene_bytes(self.C, self.n_tail, self.H, self.W)
        self.scene = aligned(n)
        call("slr_scene_prep", p(self.feat), p(self.Z), p(self.zsub), p(self.tail), self.n_tail,
             p(self.scene), self.C, self.H, self.W, None)

    @classmethod
    def from_buffer(cls, scene, motion, C, H, W):
        """Over an already prepared scene buffer (what sharding.SceneExchange delivers): no features, no Z."""
        self = cls.__new__(cls)
        self.feat = self.Z = self.tail = self.zsub = None
        self.n_tail, self.C, self.H, self.W = 0, C, H, W
        self.motion = f32(motion)
        assert scene.ctypes.data % 32 == 0
        self.scene = scene
        return self

    def table(self, start, end, t0, n):
        """slr_clip_table for frames t0 .. t0+n-1; returns the handle frames(..., table=) takes."""
        nb = lib().slr_clip_table_bytes(self.H, self.W, n)
        buf = aligned(nb)
        buf[:] = 0x5A
        call("slr_clip_table", p(self.motion), self.H, self.W, start, end, t0, n, p(buf), nb, None)
        return dict(buf=buf, bytes=nb, t0=t0, n=n, start=start, end=end)

    def frames(self, start, end, t0, n, alpha_clamp=(0.0, 1.0), want_aux=False, want_mask=False, split=False,
               table=None, want_nnz=False):
        C, H, W = self.C, self.H, self.W
        out = np.full((n, C, H, W), np.nan, dtype=np.float32)
        aux = np.full((n, self.n_tail + 1, H, W), np.nan, dtype=np.float32) if want_aux else None
        mask = np.full((n, 1, H, W), np.nan, dtype=np.float32) if want_mask else None
        nnz = np.full((n, 1, H, W), np.nan, dtype=np.float32) if want_nnz else None
        nb = lib().slr_clip_workspace_bytes(H, W, n)
        ws = aligned(nb)
        ws[:] = 0xA5          # the library must not rely on a zeroed workspace
        args = (C, self.n_tail, H, W, start, end, t0, n, alpha_clamp[0], alpha_clamp[1])
        if table is not None:
            assert (table["start"], table["end"]) == (start, end)
            call("slr_clip_bin", p(table["buf"]), table["bytes"], H, W, table["n"], t0 - table["t0"], n, p(ws), nb, None)
        if split or table is not None:
            if table is None:
                call("slr_clip_plan", p(self.motion), H, W, start, end, t0, n, p(ws), nb, None)
            call("slr_clip_expand", p(self.scene), p(self.motion), *args, p(ws), nb, None)
            for entry in ("slr_clip_gather", "slr_clip_heavy"):
                call(entry, p(self.scene), p(self.motion), *args, p(out), p(aux), p(mask), p(nnz), p(ws), nb, None)
        else:
            call("slr_clip_frames", p(self.scene), p(self.motion), *args, p(out), p(aux), p(mask), p(nnz), p(ws), nb, None)
        st = (ctypes.c_uint32 * 6)()
        call("slr_clip_stats_host", p(ws), nb, H, W, n, st, None)
        self.stats = dict(flagged=st[0], full=st[1], excess=st[2], excess_cap=st[3], fallback=st[4], tiles=st[5])
        res = (out,) + ((aux,) if want_aux else ()) + ((mask,) if want_mask else ()) + ((nnz,) if want_nnz else ())
        return res if len(res) > 1 else out
