"""CPU oracle for the SLR-SFS frame-synthesis hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of bench.py may import this package, and only as the
checker (or the timed CPU baseline) -- never from slr-sfs_b200/.

numpy front end over two shared libraries:
  * ``oracle/liboracle.so``  -- our C restatement (oracle/slr_oracle.c), every
    function citing the reference lines it follows;
  * ``oracle/_ref/libref_softsplat.so`` -- the reference's OWN kernel text compiled
    for the CPU by oracle/build.py (present when built in the container that has
    /root/reference; it travels to the GPU box prebuilt).  Exposed as ``ref_*``.

Plus numpy restatements of the Python-level pieces of the reference:
  * ``function_softsplat``   models/softsplat.py:665-690  (FunctionSoftsplat)
  * ``max_warp_norm``        models/softsplat.py:576-624  (_FunctionMaximumWarpNormsplat)
  * ``euler``                models/projection/euler_integration_manipulator.py:7-56
  * ``joint_splat_baseline`` models/animating_softmax_splating.py:847-924
  * ``joint_splat_2layer``   models/animating_softmax_splating_2layers_alpha_seperate.py:921-1045
"""
import ctypes
import os

import numpy as np

from . import build as _build

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_long = ctypes.c_long

_lib = None
_ref = None


def _oracle_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build_oracle())
    return _lib


def ref_available():
    """True when the CPU build of the reference's own kernels can be loaded."""
    return _ref_lib(required=False) is not None


def _ref_lib(required=True):
    global _ref
    if _ref is None:
        path = _build.build_ref()
        if path is None or not os.path.exists(path):
            if required:
                raise RuntimeError("oracle/_ref is not built (needs /root/reference once)")
            return None
        _ref = ctypes.CDLL(path)
    return _ref


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(_f32p)


def _dims(a):
    assert a.ndim == 4
    return [_i64(int(s)) for s in a.shape]


def _ldims(a):
    return [_long(int(s)) for s in a.shape]


def _check(inp, flow):
    assert inp.ndim == 4 and flow.ndim == 4
    assert flow.shape[1] == 2 and flow.shape[0] == inp.shape[0]
    assert flow.shape[2:] == inp.shape[2:]


# ----------------------------------------------------------------------------
# our restatement (liboracle.so)
# ----------------------------------------------------------------------------
def softsplat_sum(inp, flow, out=None):
    """Summation splat, softsplat.py:157-202.  ``out`` (accumulated into) or zeros."""
    inp, flow = _c(inp), _c(flow)
    _check(inp, flow)
    if out is None:
        out = np.zeros_like(inp)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == inp.shape
    _oracle_lib().orc_softsplat_sum_fwd(_p(inp), _p(flow), _p(out), *_dims(inp))
    return out


def softsplat_sum_f64(inp, flow):
    """Same scatter, double accumulators (tolerance yardstick)."""
    inp, flow = _c(inp), _c(flow)
    _check(inp, flow)
    out = np.zeros(inp.shape, dtype=np.float64)
    _oracle_lib().orc_softsplat_sum_fwd_f64acc(_p(inp), _p(flow), out.ctypes.data_as(_f64p), *_dims(inp))
    return out


def softsplat_grad_input(flow, gout):
    flow, gout = _c(flow), _c(gout)
    _check(gout, flow)
    gin = np.zeros_like(gout)
    _oracle_lib().orc_softsplat_grad_input(_p(flow), _p(gout), _p(gin), *_dims(gout))
    return gin


def softsplat_grad_flow(inp, flow, gout):
    inp, flow, gout = _c(inp), _c(flow), _c(gout)
    _check(inp, flow)
    assert gout.shape == inp.shape
    gflow = np.zeros_like(flow)
    _oracle_lib().orc_softsplat_grad_flow(_p(inp), _p(flow), _p(gout), _p(gflow), *_dims(inp))
    return gflow


def maxsplat(inp, flow, init):
    inp, flow = _c(inp), _c(flow)
    _check(inp, flow)
    out = np.full(inp.shape, init, dtype=np.float32)
    _oracle_lib().orc_maxsplat_fwd(_p(inp), _p(flow), _p(out), *_dims(inp))
    return out


def max_warp_norm(inp, flow):
    """_FunctionMaximumWarpNormsplat, softsplat.py:576-624: max-splat into a -1000
    filled buffer (:590), then gather the max back to each source starting from the
    source's own value (:607)."""
    inp, flow = _c(inp), _c(flow)
    warped = maxsplat(inp, flow, -1000.0)
    out = inp.copy()
    _oracle_lib().orc_inversesplat(_p(warped), _p(flow), _p(out), *_dims(inp))
    return out


def euler(motion, T):
    """euler_integration(motion, T) -> (displacements [1,2,H,W], visible [1,1,H,W]),
    euler_integration_manipulator.py:7-56."""
    motion = _c(motion)
    assert motion.ndim == 4 and motion.shape[0] == 1 and motion.shape[1] == 2
    H, W = motion.shape[2:]
    T = int(np.asarray(T).reshape(-1)[0])
    disp = np.zeros((1, 2, H, W), dtype=np.float32)
    vis = np.ones((1, 1, H, W), dtype=np.float32)
    _oracle_lib().orc_euler(_p(motion), ctypes.c_int(T), _p(disp), _p(vis), _i64(H), _i64(W))
    return disp, vis


def euler_grad_motion(motion, T, grad_disp):
    """d(sum(displacements * grad_disp)) / d(motion) as autograd derives it from the reference's
    euler_integration (euler_integration_manipulator.py:36-55); see orc_euler_grad_motion."""
    motion, grad_disp = _c(motion), _c(grad_disp)
    assert motion.ndim == 4 and motion.shape[0] == 1 and motion.shape[1] == 2 and grad_disp.shape == motion.shape
    H, W = motion.shape[2:]
    out = np.zeros_like(motion)
    _oracle_lib().orc_euler_grad_motion(_p(motion), ctypes.c_int(int(T)), _p(grad_disp), _p(out), _i64(H), _i64(W))
    return out


# ----------------------------------------------------------------------------
# the reference's own kernels on the CPU (oracle/_ref)
# ----------------------------------------------------------------------------
def ref_max_threads():
    return int(_ref_lib().ref_max_threads())


def ref_softsplat_sum(inp, flow, threads=1):
    inp, flow = _c(inp), _c(flow)
    _check(inp, flow)
    out = np.zeros_like(inp)
    _ref_lib().ref_softsplat_fwd(_p(inp), _p(flow), _p(out), *_ldims(inp), ctypes.c_int(threads))
    return out


def ref_softsplat_grad_input(inp, flow, gout, threads=1):
    inp, flow, gout = _c(inp), _c(flow), _c(gout)
    gin = np.zeros_like(inp)
    _ref_lib().ref_softsplat_grad_input(_p(inp), _p(flow), _p(gout), _p(gin), *_ldims(inp), ctypes.c_int(threads))
    return gin


def ref_softsplat_grad_flow(inp, flow, gout, threads=1):
    inp, flow, gout = _c(inp), _c(flow), _c(gout)
    gflow = np.zeros_like(flow)
    _ref_lib().ref_softsplat_grad_flow(_p(inp), _p(flow), _p(gout), _p(gflow), *_ldims(inp), ctypes.c_int(threads))
    return gflow


def ref_max_warp_norm(inp, flow, threads=1):
    inp, flow = _c(inp), _c(flow)
    warped = np.full(inp.shape, -1000.0, dtype=np.float32)
    _ref_lib().ref_maxsplat_fwd(_p(inp), _p(flow), _p(warped), *_ldims(inp), ctypes.c_int(threads))
    out = inp.copy()
    _ref_lib().ref_inversesplat(_p(warped), _p(flow), _p(out), *_ldims(inp), ctypes.c_int(threads))
    return out


# ----------------------------------------------------------------------------
# Python-level pieces of the reference, restated in numpy fp32
# ----------------------------------------------------------------------------
def function_softsplat(inp, flow, metric, mode, splat=softsplat_sum):
    """FunctionSoftsplat, softsplat.py:665-690.  Non-summation modes divide by the
    last channel after replacing exact zeros with 1 (:684) -- no epsilon."""
    assert metric is None or metric.shape[1] == 1
    assert mode in ("summation", "average", "linear", "softmax")
    inp = _c(inp)
    if mode == "average":
        inp = np.concatenate([inp, np.ones_like(inp[:, :1])], 1)
    elif mode == "linear":
        inp = np.concatenate([inp * metric, metric], 1)
    elif mode == "softmax":
        e = np.exp(_c(metric))
        inp = np.concatenate([inp * e, e], 1)
    out = splat(inp, flow)
    if mode != "summation":
        norm = out[:, -1:].copy()
        norm[norm == 0.0] = 1.0
        out = out[:, :-1] / norm
    return out


def blend_alpha(t, start, end):
    """alpha = 1 - (mid-start)/(end-start+1) in fp32, animating_softmax_splating.py:860."""
    a = np.float32(t - start) / np.float32(end - start + 1)
    return np.float32(1.0) - a


def joint_splat_baseline(feat, Z, motion, index, splat=softsplat_sum, z_mode="max"):
    """AnimatingSoftmaxSplating.forward_flow joint block,
    animating_softmax_splating.py:847-924, returning gen_fs [1,C,H,W]."""
    feat, Z, motion = _c(feat), _c(Z), _c(motion)
    start, mid, end = [int(v) for v in index]
    fwd, _ = euler(motion, mid - start)                      # :847
    bwd, _ = euler(-motion, end - mid + 1)                   # :848
    if z_mode == "v2":                                        # :849-851
        Zn = Z - max_warp_norm(Z, fwd)
    elif z_mode == "v1":                                      # :852-853
        Zn = Z
    else:                                                     # :855
        Zn = Z - Z.max()
    alpha = blend_alpha(mid, start, end)                      # :860
    eZ = np.exp(Zn)
    in_f = np.concatenate([feat * eZ * alpha, eZ * alpha], 1)             # :862
    acc = splat(in_f, fwd)                                                # :884
    one_m = np.float32(1.0) - alpha
    in_p = np.concatenate([feat * eZ * one_m, eZ * one_m], 1)             # :895
    acc_p = splat(in_p, bwd)                                              # :916
    gen = acc[:, :-1] + acc_p[:, :-1]                                     # :920
    norm = acc[:, -1:] + acc_p[:, -1:]                                    # :921
    norm = np.maximum(norm, np.float32(1e-8))                             # :923
    return gen / norm                                                     # :924


def _sigmoid(x):
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def joint_splat_2layer(feat, Z, a_fluid, a_bg_sigmoid, motion, index, alpha0=True,
                       splat=softsplat_sum):
    """AnimatingSoftmaxSplatingJoint.forward_flow joint block,
    animating_softmax_splating_2layers_alpha_seperate.py:921-1045.
    ``a_fluid`` is the raw alpha-encoder fluid channel (:946), ``a_bg_sigmoid`` the
    background alpha after the sigmoid (:948).  Returns (gen_fs, alpha_fluid, mask)."""
    feat, Z, motion = _c(feat), _c(Z), _c(motion)
    a_fluid, a_bg = _c(a_fluid), _c(a_bg_sigmoid)
    start, mid, end = [int(v) for v in index]
    fwd, _ = euler(motion, mid - start)                                   # :921
    bwd, _ = euler(-motion, end - mid + 1)                                # :922
    alpha = blend_alpha(mid, start, end)                                  # :950
    alpha = np.float32(min(max(alpha, np.float32(1.0 / 600.0)), np.float32(599.0 / 600.0)))  # :952
    Zn = Z - Z.max()                                                      # :961
    eZ = np.exp(Zn)
    one_m = np.float32(1.0) - alpha
    if alpha0:                                                            # :963-972
        s = _sigmoid(a_fluid)
        A = s / np.maximum(s + a_bg, np.float32(1e-8))
        eA = np.exp(A)
        in_f = np.concatenate([feat * eZ * alpha, a_fluid * eA * alpha, eA * alpha, eZ * alpha], 1)
        in_p = np.concatenate([feat * eZ * one_m, a_fluid * eA * one_m, eA * one_m, eZ * one_m], 1)
        n_tail = 3
    else:                                                                 # :974-976
        in_f = np.concatenate([feat * eZ * alpha, a_fluid * eZ * alpha, eZ * alpha], 1)
        in_p = np.concatenate([feat * eZ * one_m, a_fluid * eZ * one_m, eZ * one_m], 1)
        n_tail = 2
    acc = splat(in_f, fwd) + splat(in_p, bwd)                             # :987-1036
    gen = acc[:, :-n_tail]
    alpha_fluid = acc[:, -n_tail:-n_tail + 1]
    norm = np.maximum(acc[:, -1:], np.float32(1e-8))                      # :1038
    mask = (norm > np.float32(1e-8)).astype(np.float32)                   # :1039
    gen = gen / norm                                                      # :1040
    if alpha0:
        a_norm = np.maximum(acc[:, -2:-1], np.float32(1e-8))              # :1042
        alpha_fluid = alpha_fluid / a_norm                                # :1043
    else:
        alpha_fluid = alpha_fluid / norm                                  # :1045
    return gen, alpha_fluid, mask


def warp_flow_block(image, forward_flow, backward_flow, index, splat=softsplat_sum):
    """AnimatingSoftmaxSplating.warp_flow, animating_softmax_splating.py:1064-1138: the RGB twin of the
    joint block with Z = 1, alpha without the "+ 1" (:1064), precomputed flows, e^(Z - max) forward
    and e^Z backward (:1067, :1103).  Returns PredImg."""
    image = _c(image)
    start, mid, end = [int(v) for v in index]
    alpha = np.float32(1.0) - np.float32(mid - start) / np.float32(end - start)          # :1064
    Z = np.ones_like(image[:, :1])                                                       # :1066
    Zn = Z - Z.max()                                                                     # :1067
    in_f = np.concatenate([image * np.exp(Zn) * alpha, np.exp(Zn) * alpha], 1)           # :1068
    acc = splat(in_f, _c(forward_flow))                                                  # :1094
    one_m = np.float32(1.0) - alpha
    in_p = np.concatenate([image * np.exp(Z) * one_m, np.exp(Z) * one_m], 1)             # :1103
    acc_p = splat(in_p, _c(backward_flow))                                               # :1128
    gen = acc[:, :-1] + acc_p[:, :-1]                                                    # :1133
    norm = np.maximum(acc[:, -1:] + acc_p[:, -1:], np.float32(1e-8))                     # :1134-1137
    return gen / norm                                                                    # :1138


def producer_splat(fs, zn, flow, alpha, splat=softsplat_sum):
    """One direction of the TRAINING forward: tenInput = cat([fs * zn.exp() * alpha, zn.exp() * alpha], 1)
    (animating_softmax_splating.py:606 / :651) through the summation splat (:629-632 / :672-676).
    fs [B,C,H,W], zn [B,1,H,W] (already normalised and clamped), flow [B,2,H,W], alpha [B]."""
    a = np.asarray(alpha, np.float32).reshape(-1, 1, 1, 1)
    ez = np.exp(zn.astype(np.float32)).astype(np.float32)
    ten = np.concatenate([(fs * ez) * a, ez * a], 1).astype(np.float32)
    return splat(ten, flow), ten


def producer_splat_grads(fs, zn, flow, alpha, grad_acc):
    """(d_fs, d_zn, d_flow) of sum(producer_splat(...) * grad_acc): the chain rule autograd applies in the
    reference -- the splat's two backward kernels (softsplat.py:204-326) on tenInput, then the cat / mul / exp
    nodes -- written out in numpy on top of the pinned restatements of those kernels."""
    a = np.asarray(alpha, np.float32).reshape(-1, 1, 1, 1)
    _, ten = producer_splat(fs, zn, flow, alpha)
    ez = np.exp(zn.astype(np.float32)).astype(np.float32)
    g_ten = softsplat_grad_input(flow, grad_acc)                  # [B, C+1, H, W]
    d_flow = softsplat_grad_flow(ten, flow, grad_acc)
    C = fs.shape[1]
    ga = g_ten * a
    d_fs = (ga[:, :C] * ez).astype(np.float32)
    d_ez = (ga[:, :C].astype(np.float64) * fs).sum(1, keepdims=True) + ga[:, C:]
    d_zn = (d_ez * ez).astype(np.float32)
    return d_fs, d_zn, d_flow


def joint_block_training(start_fs, end_fs, Z_f, Z_p, flow_f, flow_p, alpha, splat=softsplat_sum):
    """The softmax-splatter branch of AnimatingSoftmaxSplating.forward between the Euler integration and the
    decoder (animating_softmax_splating.py:584-692), default options: Z - Z.max(), clamp to [-20, 20]."""
    a = np.asarray(alpha, np.float32).reshape(-1)
    zf = np.clip(Z_f - Z_f.max(), -20.0, 20.0).astype(np.float32)
    zp = np.clip(Z_p - Z_p.max(), -20.0, 20.0).astype(np.float32)
    acc = producer_splat(start_fs, zf, flow_f, a, splat)[0] + producer_splat(end_fs, zp, flow_p, np.float32(1.0) - a, splat)[0]
    return acc[:, :-1] / np.maximum(acc[:, -1:], np.float32(1e-8))
