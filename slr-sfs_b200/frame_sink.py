"""Frame sink: decoded frames -> 8-bit images in host memory without the float frame
ever crossing PCIe.

The reference (test_animating/test_v1_4eval_rawsize.py:240-242, 284-286) resizes each decoded frame to
the raw size, moves the FLOAT image to the host (12 bytes per pixel), scales, converts RGB -> BGR and
lets cv2.imwrite round and saturate -- per frame, synchronously, on one core.  ``FrameSink.push`` runs
``slr_frame_sink_u8`` (resize + scale + round + saturate + swap + interleave, one kernel) on the stream
that produced the frames and queues the 3-bytes-per-pixel result into a ring of pinned host buffers on a
copy stream; ``pop`` hands the finished images out in order.  Nothing blocks until a slot is needed again.
"""
import collections

import torch

from . import _lib


def to_u8(frames, out_size=None, mul=0.5, add=0.5, bgr=True, out=None):
    """frames [n,3,H,W] fp32 on the device -> [n,out_H,out_W,3] uint8 on the device (current stream)."""
    assert frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous() and frames.dim() == 4 and frames.shape[1] == 3
    n, _, H, W = frames.shape
    Ho, Wo = (H, W) if out_size is None else out_size
    if out is None:
        out = torch.empty(n, Ho, Wo, 3, dtype=torch.uint8, device=frames.device)
    assert out.shape == (n, Ho, Wo, 3) and out.dtype == torch.uint8 and out.is_contiguous()
    with torch.cuda.device(frames.device):
        _lib.call("slr_frame_sink_u8", _lib.ptr(frames), _lib.ptr(out), n, H, W, Ho, Wo, float(mul), float(add), int(bool(bgr)),
                  _lib.current_stream(frames.device))
    return out


class FrameSink:
    """Ring of ``slots`` pinned host buffers of ``group`` frames each."""

    def __init__(self, H, W, device, out_size=None, group=8, slots=3, mul=0.5, add=0.5, bgr=True):
        self.device = torch.device(device)
        self.out_size = (H, W) if out_size is None else tuple(out_size)
        self.group, self.mul, self.add, self.bgr = group, mul, add, bgr
        Ho, Wo = self.out_size
        self.stage = [torch.empty(group, Ho, Wo, 3, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self.host = [torch.empty(group, Ho, Wo, 3, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.landed = [None] * slots           # event: the slot's D2H copy has finished
        self.pending = collections.deque()     # (slot, n, first frame index)
        self.turn = 0
        self.count = 0

    def push(self, frames):
        """Queue frames [n<=group,3,H,W] (valid on the current stream).  Returns immediately unless
        every slot still holds images nobody has popped."""
        n = frames.shape[0]
        assert n <= self.group
        slot = self.turn
        self.turn = (self.turn + 1) % len(self.stage)
        assert all(p[0] != slot for p in self.pending), "FrameSink: every slot is full -- pop() before pushing more"
        main = torch.cuda.current_stream(self.device)
        if self.landed[slot] is not None:
            main.wait_event(self.landed[slot])             # the device staging buffer is being copied out
        to_u8(frames, self.out_size, self.mul, self.add, self.bgr, out=self.stage[slot][:n])
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[slot][:n].copy_(self.stage[slot][:n], non_blocking=True)
            self.landed[slot] = torch.cuda.Event()
            self.landed[slot].record(self.copy_stream)
        self.pending.append((slot, n, self.count))
        self.count += n
        return self.count - n

    def pop(self):
        """(first frame index, images [n,out_H,out_W,3] uint8 numpy view of the pinned slot) of the oldest
        queued group, waiting for its copy if necessary; None when nothing is queued.  The view is
        valid until the slot comes round again."""
        if not self.pending:
            return None
        slot, n, first = self.pending.popleft()
        self.landed[slot].synchronize()
        return first, self.host[slot][:n].numpy()
