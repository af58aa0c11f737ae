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


def euler_integration(motion, destination_frame, return_all_frames=False):
    """Repeatedly integrate the Eulerian motion field; see module docstring.

    Returns displacements [1,2,H,W] and visible_pixels [1,1,H,W] on ``motion``'s
    device (with ``return_all_frames`` -- broken in the reference, :31,:50 -- the
    leading dimension is destination_frame+1, one entry per step count).
    """
    assert (motion.dim() == 4)
    b, c, height, width = motion.shape
    assert (b == 1), 'Function only implemented for batch = 1'
    assert (c == 2), f'Input motion field should be Bx2xHxW. Given tensor is: {motion.shape}'
    if not motion.is_cuda:
        raise NotImplementedError()    # the reference hard-codes device='cuda' (:24-35)
    T = _as_int(destination_frame)
    assert T >= 0
    motion = motion.detach()
    if motion.dtype != torch.float32 or not motion.is_contiguous():
        motion = motion.float().contiguous()
    steps = list(range(T + 1)) if return_all_frames else [T]
    displacements = motion.new_empty(len(steps), 2, height, width)
    visible_pixels = motion.new_empty(len(steps), 1, height, width)
    with torch.cuda.device(motion.device):
        stream = _lib.current_stream(motion.device)
        for i, t in enumerate(steps):
            _lib.call("slr_euler", _lib.ptr(motion), 1.0, t, _lib.ptr(displacements[i]),
                      _lib.ptr(visible_pixels[i]), height, width, stream)
    return displacements, visible_pixels


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
