"""The joint block of the TRAINING forward as a fused, differentiable operator.

``AnimatingSoftmaxSplating.forward`` (models/animating_softmax_splating.py:577-692) builds, per
direction, ``tenInput = cat([fs * Z_norm.exp() * alpha, Z_norm.exp() * alpha], 1)`` (:606, :651), splats
it with ``ModuleSoftsplat('summation')`` (:629-632, :672-676), adds the two results and divides by the
clamped last channel (:684-692); autograd then walks back through the splat's two backward kernels
and the cat / exp / mul nodes.  The two directions use different feature sets (``start_fs`` / ``end_fs``)
and importances (``Z_f`` / ``Z_p``), and alpha differs per sample of the batch.

``producer_splat`` is that producer + splat as ONE kernel forward (tenInput never exists) and ONE
kernel backward (the four gradient corners of a source pixel are read once and feed d(fs), d(Z) and
d(flow) together).  ``joint_block_training`` strings the two directions and the normalisation
together with exactly the reference's operations around them (Z normalisation, clamp, division), all
differentiable: drop it into ``forward`` in place of lines :586-692.
"""
import torch

from . import _lib


def _c(t):
    assert t.is_cuda and t.dtype == torch.float32, "CUDA fp32 tensors only (no CPU path)"
    return t.contiguous()


class _FunctionProducerSplat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fs, zn, flow, alpha, acc):
        fs, zn, flow, alpha = _c(fs), _c(zn), _c(flow), _c(alpha)
        B, C, H, W = fs.shape
        assert zn.shape == (B, 1, H, W) and flow.shape == (B, 2, H, W) and alpha.numel() == B
        accumulate = acc is not None
        if acc is None:
            acc = fs.new_empty(B, C + 1, H, W)
        else:
            assert acc.shape == (B, C + 1, H, W) and acc.is_contiguous()
            ctx.mark_dirty(acc)
        with torch.cuda.device(fs.device):
            _lib.call("slr_producer_splat_fwd", _lib.ptr(fs), _lib.ptr(zn), _lib.ptr(flow), _lib.ptr(alpha), _lib.ptr(acc),
                      B, C, H, W, 1 if accumulate else 0, _lib.current_stream(fs.device))
        ctx.save_for_backward(fs, zn, flow, alpha)
        ctx.accumulate = accumulate
        return acc

    @staticmethod
    def backward(ctx, grad_acc):
        fs, zn, flow, alpha = ctx.saved_tensors
        B, C, H, W = fs.shape
        grad_acc = _c(grad_acc)
        need = ctx.needs_input_grad
        d_fs = torch.empty_like(fs) if need[0] else None
        d_zn = torch.empty_like(zn) if need[1] else None
        d_flow = torch.empty_like(flow) if need[2] else None
        if d_fs is not None or d_zn is not None or d_flow is not None:
            with torch.cuda.device(fs.device):
                _lib.call("slr_producer_splat_bwd", _lib.ptr(fs), _lib.ptr(zn), _lib.ptr(flow), _lib.ptr(alpha), _lib.ptr(grad_acc),
                          _lib.ptr(d_fs), _lib.ptr(d_zn), _lib.ptr(d_flow), B, C, H, W, _lib.current_stream(fs.device))
        # alpha comes from the frame indices (:584-585): no gradient; the accumulator passes its gradient through
        return d_fs, d_zn, d_flow, None, (grad_acc if ctx.accumulate else None)


def producer_splat(fs, zn, flow, alpha, acc=None):
    """splat(cat([fs * zn.exp() * alpha, zn.exp() * alpha], 1), flow) -> [B, C+1, H, W], differentiable in
    fs, zn and flow.  ``alpha``: [B] tensor (any shape with B elements).  ``acc``: add into this accumulator
    (the second direction) instead of allocating a new one."""
    return _FunctionProducerSplat.apply(fs, zn, flow, alpha.reshape(-1), acc)


def normalise_importance(Z, mode="max", clamp=True):
    """The reference's importance normalisation in front of the producer (:592-605): 'max' Z - Z.max(),
    'v1' Z as it is, 'v3' sigmoid(Z) * 20; then clamp to [-20, 20] unless ``no_clamp_Z``."""
    if mode == "max":
        Z = Z - Z.max()
    elif mode == "v3":
        Z = torch.sigmoid(Z) * 20
    else:
        assert mode == "v1", mode
    return torch.clamp(Z, min=-20.0, max=20.0) if clamp else Z


def joint_block_training(start_fs, end_fs, Z_f, Z_p, flow_f, flow_p, alpha, z_mode="max", clamp=True):
    """gen_fs [B,C,H,W] of the training forward (:586-692, softmax-splatter branch): forward splat of
    ``start_fs`` with ``Z_f`` along ``flow_f`` weighted alpha, backward splat of ``end_fs`` with ``Z_p`` along
    ``flow_p`` weighted 1 - alpha, summed, divided by the clamped accumulated weight.  ``alpha``: [B] (or
    [B,1,1,1]) as the reference computes it from the frame indices (:584-585)."""
    alpha = alpha.reshape(-1).to(torch.float32)
    acc = producer_splat(start_fs, normalise_importance(Z_f, z_mode, clamp), flow_f, alpha)
    # the backward direction's importance is never v3 in the reference (:644-647)
    zp_mode = "max" if z_mode == "v3" else z_mode
    acc = producer_splat(end_fs, normalise_importance(Z_p, zp_mode, clamp), flow_p, 1.0 - alpha, acc)
    return acc[:, :-1] / torch.clamp(acc[:, -1:], min=1e-8)
