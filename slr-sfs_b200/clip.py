"""Clip-level driver: the reference's frame loop (test_animating/test_v1_4eval_rawsize.py:233-239:
``for t in range(N): forward_flow(batch)``) for one rank's block of frames, in groups that a
consumer (the decoder; a D2H copy; a checksum) takes over while the next group is synthesised.

bench.py times exactly this object, and tests/test_gpu_bench_path.py checks its output frame
by frame at the benchmark size -- the measured path and the verified path are the same code."""
import torch

from .synthesis import JointSplat


class ClipRunner:
    """Two output buffers of ``group`` frames each, reused for every scene; ``run`` fills them
    alternately and hands each finished group to ``on_frames``.

    ``on_frames(frames [k,C,H,W], t0) -> event | None`` runs on the current stream right after the
    group was queued; the event it returns (if any) says when the buffer may be overwritten."""

    def __init__(self, C, H, W, device, group):
        self.C, self.H, self.W, self.device, self.group = C, H, W, torch.device(device), int(group)
        self.bufs = [torch.empty(self.group, C, H, W, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.free = [None, None]        # events after which a buffer may be overwritten
        self.turn = 0

    def run(self, js, start, end, lo, hi, on_frames=None, group=None):
        """Frames lo .. hi-1 of the clip [start, end] of scene ``js`` (a JointSplat).  Returns the
        last group (a view of one of the two buffers)."""
        assert isinstance(js, JointSplat) and (js.C, js.H, js.W) == (self.C, self.H, self.W)
        group = min(group or self.group, self.group)
        out = None
        if hi <= lo:
            return out
        js.prepare_clip(start, end, lo, hi - lo)          # both Euler chains once for the whole block
        for b0 in range(lo, hi, group):
            nb = min(group, hi - b0)
            slot = self.turn = self.turn ^ 1
            if self.free[slot] is not None:
                torch.cuda.current_stream(self.device).wait_event(self.free[slot])
                self.free[slot] = None
            out = js.frames(start, end, b0, nb, out=self.bufs[slot][:nb])
            if on_frames is not None:
                self.free[slot] = on_frames(out, b0)
        return out
