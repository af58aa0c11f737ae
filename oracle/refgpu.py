"""The reference's own summation-splat kernel on the GPU (R-GPU row of BASELINE.md).
TEST / BENCH INFRASTRUCTURE ONLY -- never imported by slr-sfs_b200/.

oracle/build.py::build_ref_gpu() compiles the reference's kernel text -- templated
per shape by the reference's own cupy_kernel() -- with nvcc into
oracle/_ref/libref_softsplat_gpu.so.  ``reference_frame`` drives it the way
AnimatingSoftmaxSplating.forward_flow does (models/animating_softmax_splating.py:847-924):
euler_integration restated with the same eager torch ops as the reference
(models/projection/euler_integration_manipulator.py:18-56, including its boolean-mask
indexing), torch glue, two kernel launches, in-place adds, clamp, divide.
"""
import ctypes
import os

import torch

from . import build as _build

_lib = None


def available():
    return _load(required=False) is not None


def _load(required=True):
    global _lib
    if _lib is None:
        path = _build.build_ref_gpu()
        if path is None or not os.path.exists(path):
            if required:
                raise RuntimeError("oracle/_ref/libref_softsplat_gpu.so is not built")
            return None
        _lib = ctypes.CDLL(path)
        _lib.refgpu_softsplat_fwd.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_long] * 4 + [ctypes.c_void_p]
        _lib.refgpu_softsplat_bwd.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_long] * 4 + [ctypes.c_void_p]
    return _lib


def baked_shapes():
    return list(_build.REF_GPU_SHAPES)


def softsplat_sum(inp, flow):
    """_FunctionSoftsplat.forward (softsplat.py:390-423) with the reference kernel."""
    assert inp.is_cuda and inp.is_contiguous() and flow.is_contiguous()
    out = inp.new_zeros(inp.shape)
    B, C, H, W = inp.shape
    rc = _load().refgpu_softsplat_fwd(inp.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, H, W,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("reference kernel launch failed (%d); shape %s baked in? %s" % (rc, tuple(inp.shape), baked_shapes()))
    return out


def baked_backward_shapes():
    return list(_build.REF_GPU_BWD_SHAPES)


def softsplat_backward(inp, flow, gout, need_input=True, need_flow=True):
    """_FunctionSoftsplat.backward (softsplat.py:427-477) with the reference kernels:
    (gradInput | None, gradFlow | None), zero-initialised like the reference (:440-441)."""
    assert inp.is_cuda and inp.is_contiguous() and flow.is_contiguous() and gout.is_contiguous()
    B, C, H, W = inp.shape
    gin = inp.new_zeros(inp.shape) if need_input else None
    gflow = inp.new_zeros(flow.shape) if need_flow else None
    rc = _load().refgpu_softsplat_bwd(inp.data_ptr(), flow.data_ptr(), gout.data_ptr(),
                                      None if gin is None else gin.data_ptr(), None if gflow is None else gflow.data_ptr(),
                                      B, C, H, W, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("reference backward launch failed (%d); shape %s baked in? %s" % (rc, tuple(inp.shape), baked_backward_shapes()))
    return gin, gflow


def euler_integration(motion, destination_frame):
    """The reference's eager-op Euler loop, restated op for op (euler_integration_manipulator.py:18-56)."""
    b, c, height, width = motion.shape
    dev = motion.device
    y, x = torch.meshgrid([torch.linspace(0, height - 1, height, device=dev),
                           torch.linspace(0, width - 1, width, device=dev)], indexing="ij")
    coord = torch.stack([x, y], dim=0).long()
    destination_coords = coord.clone().float()
    displacements = torch.zeros(1, 2, height, width, device=dev)
    invalid_mask = torch.zeros(1, height, width, device=dev).bool()
    for _ in range(1, int(destination_frame) + 1):
        destination_coords = destination_coords + motion[0][:, torch.round(destination_coords[1]).long(),
                                                            torch.round(destination_coords[0]).long()]
        oob_x = torch.logical_or(destination_coords[0] > (width - 1), destination_coords[0] < 0)
        oob_y = torch.logical_or(destination_coords[1] > (height - 1), destination_coords[1] < 0)
        invalid_mask = torch.logical_or(oob_x.unsqueeze(0), invalid_mask)
        invalid_mask = torch.logical_or(oob_y.unsqueeze(0), invalid_mask)
        destination_coords[invalid_mask.expand_as(destination_coords)] = coord[invalid_mask.expand_as(destination_coords)].float()
        displacements = (destination_coords - coord.float()).unsqueeze(0)
        displacements[invalid_mask.unsqueeze(0).repeat(1, 2, 1, 1)] = torch.max(torch.Tensor([height, width])) + 1
    return displacements


def reference_frame(feat, Z, motion, index):
    """One forward_flow joint block (animating_softmax_splating.py:847-924) on the GPU with the
    reference kernel; returns gen_fs."""
    start, mid, end = index
    fwd = euler_integration(motion, mid - start)
    bwd = euler_integration(-motion, end - mid + 1)
    Zn = Z - Z.max()
    alpha = (1.0 - torch.tensor(float(mid - start)) / torch.tensor(float(end - start + 1))).view(1, 1, 1, 1).to(feat.device)
    in_f = torch.cat([feat * Zn.exp() * alpha, Zn.exp() * alpha], 1)
    gen_f = softsplat_sum(in_f, fwd)
    gen_fs = gen_f[:, :-1]
    norm = gen_f[:, -1:]
    in_p = torch.cat([feat * Zn.exp() * (1 - alpha), Zn.exp() * (1 - alpha)], 1)
    gen_p = softsplat_sum(in_p, bwd)
    gen_fs += gen_p[:, :-1]
    norm += gen_p[:, -1:]
    norm = torch.clamp(norm, min=1e-8)
    return gen_fs / norm


def reference_frame_2layer(feat, Z, a_fluid, a_bg_sigmoid, motion, index, alpha0=True):
    """One joint block of AnimatingSoftmaxSplatingJoint.forward_flow
    (models/animating_softmax_splating_2layers_alpha_seperate.py:921-1045) on the GPU with the
    reference kernel, restated op for op: ``a_fluid`` is the raw fluid channel of the alpha
    encoder (:946), ``a_bg_sigmoid`` the background alpha after the sigmoid (:948).
    Returns (gen_fs, alpha_fluid, alpha_fluid_mask)."""
    start, mid, end = index
    fwd = euler_integration(motion, mid - start)                                   # :921
    bwd = euler_integration(-motion, end - mid + 1)                                # :922
    alpha = (1.0 - torch.tensor(float(mid - start)) / torch.tensor(float(end - start + 1))).view(1, 1, 1, 1).to(feat.device)
    alpha = torch.clamp(alpha, min=1.0 / 600.0, max=599.0 / 600.0)                 # :952
    Zn = Z - Z.max()                                                               # :961
    if alpha0:                                                                     # :963-972
        norm0 = torch.clamp(torch.sigmoid(a_fluid) + a_bg_sigmoid, min=1e-8)
        A = torch.sigmoid(a_fluid) / norm0
        in_f = torch.cat([feat * Zn.exp() * alpha, a_fluid * A.exp() * alpha, A.exp() * alpha, Zn.exp() * alpha], 1)
        in_p = torch.cat([feat * Zn.exp() * (1 - alpha), a_fluid * A.exp() * (1 - alpha), A.exp() * (1 - alpha),
                          Zn.exp() * (1 - alpha)], 1)
        n_tail = 3
    else:                                                                          # :974-976
        in_f = torch.cat([feat * Zn.exp() * alpha, a_fluid * Zn.exp() * alpha, Zn.exp() * alpha], 1)
        in_p = torch.cat([feat * Zn.exp() * (1 - alpha), a_fluid * Zn.exp() * (1 - alpha), Zn.exp() * (1 - alpha)], 1)
        n_tail = 2
    acc = softsplat_sum(in_f.contiguous(), fwd)                                    # :987
    acc_p = softsplat_sum(in_p.contiguous(), bwd)                                  # :1024
    acc += acc_p                                                                   # :1028-1036
    gen_fs = acc[:, :-n_tail]
    alpha_fluid = acc[:, -n_tail:-n_tail + 1]
    norm = torch.clamp(acc[:, -1:], min=1e-8)                                      # :1038
    mask = (norm > 1e-8).float()                                                   # :1039
    gen_fs = gen_fs / norm                                                         # :1040
    if alpha0:
        alpha_fluid = alpha_fluid / torch.clamp(acc[:, -2:-1], min=1e-8)           # :1042-1043
    else:
        alpha_fluid = alpha_fluid / norm                                           # :1045
    return gen_fs, alpha_fluid, mask
