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
