"""Drop-in replacement for the reference's
``models/projection/euler_integration_manipulator.py``.

    euler_integration(motion, destination_frame, return_all_frames=False)
        -> (displacements, visible_pixels)                      reference :7-56
    EulerIntegration(opt=None)(motion, destination_frame,
        return_all_frames=False, show_visible_pixels=False)     reference :58-71

The reference runs ~20 tiny torch kernels per step plus boolean-mask indexing
(a host sync each) and an H2D scalar copy (:36-55).  Here the whole chain of T
steps runs in registers inside one kernel launch with no host synchronisation;
results are bit-identical (same fp32 adds in the same order, round-half-to-even).
"""
import torch
import torch.nn as nn

from . import _lib


def _as_int(destination_frame):
    # the models pass a python int or a 1-element LongTensor (used in range(), reference :36)
    if torch.is_tensor(destination_frame):
        assert destination_frame.numel() == 1
        return int(destination_frame.reshape(-1)[0].item())
    return int(destination_frame)


def _contiguous_f32(motion):
    if motion.dtype != torch.float32 or not motion.is_contiguous():
        motion = motion.float().contiguous()
    return motion


def _integrate(motion, T):
    """One slr_euler launch: displacements [1,2,H,W], visible [1,1,H,W] after T steps."""
    height, width = motion.shape[2:]
    displacements = motion.new_empty(1, 2, height, width)
    visible_pixels = motion.new_empty(1, 1, height, width)
    with torch.cuda.device(motion.device):
        _lib.call("slr_euler", _lib.ptr(motion), 1.0, T, _lib.ptr(displacements), _lib.ptr(visible_pixels),
                  height, width, _lib.current_stream(motion.device))
    return displacements, visible_pixels


class _FunctionEuler(torch.autograd.Function):
    """The chain with the reference's autograd behaviour: the displacements are differentiable
    with respect to the motion VALUES sampled along each chain (reference :37-38; the rounded
    sample positions carry no gradient), pixels that end invalid hold the sentinel constant
    (:53-55) and pass no gradient.  This is what --train_motion needs
    (models/animating_softmax_splating.py:514-580: the flow regressor's output goes through
    EulerIntegration into the splat, whose gradFlow must reach the regressor)."""

    @staticmethod
    def forward(ctx, motion, T):
        motion = _contiguous_f32(motion.detach())
        ctx.save_for_backward(motion)
        ctx.T = T
        displacements, visible_pixels = _integrate(motion, T)
        ctx.mark_non_differentiable(visible_pixels)
        return displacements, visible_pixels

    @staticmethod
    def backward(ctx, grad_displacements, _grad_visible):
        motion, = ctx.saved_tensors
        height, width = motion.shape[2:]
        grad = _contiguous_f32(grad_displacements)
        grad_motion = motion.new_empty(1, 2, height, width)
        with torch.cuda.device(motion.device):
            _lib.call("slr_euler_grad_motion", _lib.ptr(motion), 1.0, ctx.T, _lib.ptr(grad), _lib.ptr(grad_motion),
                      height, width, _lib.current_stream(motion.device))
        return grad_motion, None


def euler_integration(motion, destination_frame, return_all_frames=False):
    """Repeatedly integrate the Eulerian motion field; see module docstring.

    Returns displacements [1,2,H,W] and visible_pixels [1,1,H,W] on ``motion``'s
    device (with ``return_all_frames`` -- broken in the reference, :31,:50 -- the
    leading dimension is destination_frame+1, one entry per step count).  Differentiable
    with respect to ``motion`` like the reference's eager ops (see _FunctionEuler).
    """
    assert (motion.dim() == 4)
    b, c, height, width = motion.shape
    assert (b == 1), 'Function only implemented for batch = 1'
    assert (c == 2), f'Input motion field should be Bx2xHxW. Given tensor is: {motion.shape}'
    if not _lib.on_device(motion):
        raise NotImplementedError()    # the reference hard-codes device='cuda' (:24-35)
    T = _as_int(destination_frame)
    assert T >= 0
    steps = list(range(T + 1)) if return_all_frames else [T]
    if motion.requires_grad and torch.is_grad_enabled():
        results = [_FunctionEuler.apply(motion, t) for t in steps]
    else:
        plain = _contiguous_f32(motion.detach())
        results = [_integrate(plain, t) for t in steps]
    if len(results) == 1:
        return results[0]
    return torch.cat([r[0] for r in results], 0), torch.cat([r[1] for r in results], 0)


class EulerIntegration(nn.Module):
    def __init__(self, opt=None):
        super().__init__()
        self.opt = opt

    def forward(self, motion, destination_frame, return_all_frames=False, show_visible_pixels=False):
        displacements = torch.zeros(motion.shape).to(motion.device)
        # the reference allocates this on the CPU (:64), forcing a D2H copy per sample;
        # keeping it on the device is a deliberate difference
        visible_pixels = torch.zeros(motion.shape[0], 1, motion.shape[2], motion.shape[3], device=motion.device)
        for b in range(motion.shape[0]):
            displacements[b:b + 1], visible_pixels[b:b + 1] = euler_integration(motion[b:b + 1], destination_frame[b])
        if show_visible_pixels:
            return displacements, visible_pixels
        else:
            return displacements
