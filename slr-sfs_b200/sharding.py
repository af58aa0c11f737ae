"""Multi-GPU partitioning of the frame loop (one process per GPU).

Frames of a clip are independent given the per-scene constants (features, Z,
motion): the reference's loop ``for t in range(N): forward_flow(batch)``
(test_animating/test_v1_4eval_rawsize.py:233-239) has no carried state.  So the
path shards over frames with NO data-path collective; the only communication is
one broadcast per scene of the shared inputs from the rank that ran the encoder
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def frame_block(n_frames, rank, world):
    """Contiguous block [lo, hi) of `rank`; block sizes differ by at most one
    (60 frames over 8 ranks: 8,8,8,8,7,7,7,7)."""
    assert 0 <= rank < world and n_frames >= 0
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_scene(tensors, src, device=None, group=None):
    """Broadcast the per-scene inputs from rank `src`.  On `src`, `tensors` is the
    tuple of tensors; elsewhere it is a tuple of (shape, dtype) specs or of
    tensors to be overwritten.  Returns the tuple of tensors on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return tuple(tensors)
    rank = dist.get_rank(group)
    out = []
    for t in tensors:
        if not torch.is_tensor(t):
            shape, dtype = t
            t = torch.empty(shape, dtype=dtype, device=device)
        elif rank != src and device is not None and t.device != torch.device(device):
            t = torch.empty_like(t, device=device)
        dist.broadcast(t, src=src, group=group)
        out.append(t)
    return tuple(out)


def synthesize_sharded(make_frames, n_frames, group=None):
    """Run ``make_frames(lo, hi)`` for this rank's frame block and return (lo, hi, result).
    Results stay on their rank (a 64-channel frame is 201 MB at 768x1024; only decoded
    RGB should ever be gathered)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = frame_block(n_frames, rank, world)
    return lo, hi, (make_frames(lo, hi) if hi > lo else None)


def all_gather_frames(local, n_frames, group=None):
    """Gather small per-frame tensors (e.g. decoded RGB or checksums) from every rank's
    block into one [n_frames, ...] tensor, in frame order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [frame_block(n_frames, r, world) for r in range(world)]
    tail = local.shape[1:]
    longest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((longest,) + tuple(tail), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
