"""ctypes binding of csrc/libslr_splat.so (C ABI: include/slr_splat.h).

The CUDA library is the product.  There is no CPU or PyTorch fallback: if the
shared library is missing or a call fails, this module raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libslr_splat.so")

_f32p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_flt = ctypes.c_float
_strm = ctypes.c_void_p

# name -> argtypes; every function returns int (0 = ok) unless listed in _OTHER_RESTYPE.
SIGNATURES = {
    "slr_version": [],
    "slr_last_error_string": [],
    "slr_softsplat_sum_fwd": [_f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _int, _strm],
    "slr_softsplat_grad_input": [_f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _strm],
    "slr_softsplat_grad_flow": [_f32p, _f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _strm],
    "slr_maxsplat_fwd": [_f32p, _f32p, _f32p, _flt, _i64, _i64, _i64, _i64, _strm],
    "slr_maxwarpnorm": [_f32p, _f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _strm],
    "slr_euler": [_f32p, _flt, _int, _f32p, _f32p, _i64, _i64, _strm],
    "slr_euler_grad_motion": [_f32p, _flt, _int, _f32p, _f32p, _i64, _i64, _strm],
    "slr_softsplat_gather_scratch_bytes": [_i64, _i64, _i64],
    "slr_softsplat_sum_fwd_gather": [_f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _f32p, ctypes.c_size_t, _strm],
    "slr_producer_splat_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _int, _strm],
    "slr_producer_splat_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, _strm],
    "slr_reduce_max": [_f32p, _i64, _f32p, _strm],
    "slr_joint_scatter": [_f32p, _f32p, _f32p, _f32p, _int, _f32p, _f32p, _flt, _f32p, _i64, _i64, _i64, _strm],
    "slr_joint_scatter_weights": [_f32p, _f32p, _f32p, _f32p, _int, _f32p, _f32p, _flt, _flt, _f32p, _i64, _i64, _i64, _strm],
    "slr_normalize": [_f32p, _f32p, _f32p, _i64, _i64, _i64, _flt, _i64, _i64, _strm],
    "slr_scene_bytes": [_i64, _int, _i64, _i64],
    "slr_scene_core_bytes": [_i64, _int, _i64, _i64],
    "slr_scene_quilt": [_f32p, _i64, _int, _i64, _i64, _strm],
    "slr_scene_prep": [_f32p, _f32p, _f32p, _f32p, _int, _f32p, _i64, _i64, _i64, _strm],
    "slr_clip_workspace_bytes": [_i64, _i64, _int],
    "slr_clip_plan": [_f32p, _i64, _i64, _int, _int, _int, _int, _f32p, ctypes.c_size_t, _strm],
    "slr_clip_expand": [_f32p, _f32p, _i64, _int, _i64, _i64, _int, _int, _int, _int, _flt, _flt,
                        _f32p, ctypes.c_size_t, _strm],
    "slr_clip_gather": [_f32p, _f32p, _i64, _int, _i64, _i64, _int, _int, _int, _int, _flt, _flt,
                        _f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_size_t, _strm],
    "slr_clip_heavy": [_f32p, _f32p, _i64, _int, _i64, _i64, _int, _int, _int, _int, _flt, _flt,
                       _f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_size_t, _strm],
    "slr_clip_frames": [_f32p, _f32p, _i64, _int, _i64, _i64, _int, _int, _int, _int, _flt, _flt,
                        _f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_size_t, _strm],
    "slr_clip_table_bytes": [_i64, _i64, _int],
    "slr_clip_table": [_f32p, _i64, _i64, _int, _int, _int, _int, _f32p, ctypes.c_size_t, _strm],
    "slr_clip_bin": [_f32p, ctypes.c_size_t, _i64, _i64, _int, _int, _int, _f32p, ctypes.c_size_t, _strm],
    "slr_frame_sink_u8": [_f32p, _f32p, _i64, _i64, _i64, _i64, _i64, _flt, _flt, _int, _strm],
    "slr_clip_stats_host": [_f32p, ctypes.c_size_t, _i64, _i64, _int, ctypes.POINTER(ctypes.c_uint32), _strm],
}
_OTHER_RESTYPE = {"slr_last_error_string": ctypes.c_char_p, "slr_scene_bytes": ctypes.c_size_t, "slr_scene_core_bytes": ctypes.c_size_t,
                  "slr_clip_workspace_bytes": ctypes.c_size_t, "slr_clip_table_bytes": ctypes.c_size_t,
                  "slr_softsplat_gather_scratch_bytes": ctypes.c_size_t}

_lib = None
_lock = threading.Lock()


class SlrError(RuntimeError):
    pass


def load():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SlrError(
                    "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  slr_sfs_b200 has no CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, argtypes in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if the ABI and the binding drift apart
                fn.argtypes = argtypes
                fn.restype = _OTHER_RESTYPE.get(name, ctypes.c_int)
            _lib = lib
    return _lib


def call(name, *args):
    """Invoke an int-returning entry point and raise SlrError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.slr_last_error_string()
        raise SlrError("%s failed with status %d: %s" % (name, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def on_device(t):
    """True for tensors the library can work on (CUDA tensors).  The operator modules raise
    NotImplementedError otherwise, like the reference (softsplat.py:418-419)."""
    return bool(t.is_cuda)


def current_stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------
# instrumentation used by bench.py: launch counting and per-entry-point timing
# ---------------------------------------------------------------------------
# kernels launched by one call of each entry point (memsets are not kernels)
KERNELS_PER_CALL = {
    "slr_softsplat_sum_fwd": 1, "slr_softsplat_grad_input": 1, "slr_softsplat_grad_flow": 1,
    "slr_maxsplat_fwd": 2, "slr_maxwarpnorm": 3, "slr_euler": 1, "slr_euler_grad_motion": 1, "slr_reduce_max": 2,
    "slr_joint_scatter": 1, "slr_joint_scatter_weights": 1, "slr_normalize": 1, "slr_scene_prep": 1, "slr_scene_quilt": 1, "slr_clip_frames": 9,
    "slr_clip_plan": 3, "slr_clip_table": 2, "slr_clip_bin": 1, "slr_clip_expand": 1, "slr_clip_gather": 1, "slr_clip_heavy": 4, "slr_frame_sink_u8": 1,
    "slr_producer_splat_fwd": 1, "slr_producer_splat_bwd": 1,
    # per batch element: scene_prep, flow_table, static_lanes, slot_fill, bind_batch, insert, rowgather, heavy_excess,
    # heavy_finish, overflow x 3 (counted for one element; bench.py's level0 leg uses B = 1)
    "slr_softsplat_sum_fwd_gather": 12,
}
# the direct index (default SLR_GATHER_MODE): slr_clip_bin = slot_fill + bind_batch, slr_clip_heavy = excess + finish + the
# three overflow kernels, slr_clip_plan = euler_table + static_lanes + slot_fill + bind_batch; staged: + stagegather
_KERNELS_DIRECT = {"slr_clip_bin": 2, "slr_clip_heavy": 5, "slr_clip_plan": 4, "slr_clip_frames": 11}


def _kernels_per_call(name):
    mode = os.environ.get("SLR_GATHER_MODE", "ldg")
    if mode not in ("bins", "staged") and name in _KERNELS_DIRECT:
        return _KERNELS_DIRECT[name]
    if mode == "staged" and name in ("slr_clip_gather", "slr_clip_frames"):
        return KERNELS_PER_CALL[name] + 1
    return KERNELS_PER_CALL.get(name, 0)
_launches = 0
_timing = None          # None, or list of (name, start_event, end_event)
_timing_only = None     # None = every entry point, else the set of names to bracket


def launch_count():
    return _launches


def kernel_timing(enable, only=None):
    """enable=True: start bracketing every entry-point call (or just the ones named in `only`)
    with CUDA events on the current stream.  enable=False: stop, synchronise and return
    {entry point: (total ms, calls)}."""
    global _timing, _timing_only
    if enable:
        _timing = []
        _timing_only = None if only is None else set(only)
        return None
    import torch
    recs, _timing = _timing or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in recs:
        tot, n = out.get(name, (0.0, 0))
        out[name] = (tot + e0.elapsed_time(e1), n + 1)
    return out


def algorithmic_bytes(name, C, P, n_tail=0):
    """Algorithmic HBM bytes of one call (DESIGN.md 'Kernels')."""
    plane = 4 * P
    if name == "slr_joint_scatter":      # reads C feature + Z + 2 motion planes, writes C+1+n_tail accumulators
        return plane * (2 * C + 4 + 2 * n_tail)
    if name == "slr_normalize":          # reads C+1 accumulators, writes C
        return plane * (2 * C + 1)
    if name == "slr_euler":
        return plane * 4
    if name in ("slr_clip_frames", "slr_clip_gather"):        # per FRAME: same unit as slr_joint_scatter (section 8d, 2C+4 planes)
        return plane * (2 * C + 4 + 2 * n_tail)
    return 0


_plain_call = call


def call(name, *args):  # noqa: F811  (instrumented wrapper around the plain call)
    global _launches
    if _timing is not None and (_timing_only is None or name in _timing_only):
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _plain_call(name, *args)
        e1.record()
        _timing.append((name, e0, e1))
    else:
        _plain_call(name, *args)
    _launches += _kernels_per_call(name)
