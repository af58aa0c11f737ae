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


def frame_block(n_frames, rank, world, rotate=0):
    """Contiguous block [lo, hi) of `rank`; block sizes differ by at most one
    (60 frames over 8 ranks: 8,8,8,8,7,7,7,7).  ``rotate`` (e.g. the scene index) shifts which
    ranks get the longer blocks, so that over `world` scenes every rank synthesises the same
    number of frames (60 per 8 scenes instead of 64 on four ranks and 56 on the others)."""
    assert 0 <= rank < world and n_frames >= 0
    base, rem = divmod(n_frames, world)
    r = (rank + rotate) % world
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


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


def all_gather_frames(local, n_frames, group=None, rotate=0):
    """Gather small per-frame tensors (e.g. decoded RGB or checksums) from every rank's
    block (``frame_block(..., rotate)``) into one [n_frames, ...] tensor, in frame order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [frame_block(n_frames, r, world, rotate) for r in range(world)]
    tail = local.shape[1:]
    longest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((longest,) + tuple(tail), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    order = sorted(range(world), key=lambda r: sizes[r][0])
    return torch.cat([parts[r][: sizes[r][1] - sizes[r][0]] for r in order], 0)


class SceneExchange:
    """The per-scene broadcast of BASELINE.json configs[3], one scene ahead of the synthesis.

    What travels is the core of the PREPARED scene (synthesis.JointSplat.prepare_scene: pre-weighted,
    channel-interleaved features and e^Z -- the same 204.5 MB as features + Z at 768x1024x64; the
    receivers rebuild the staged copy behind it locally) plus the motion field, so Z.max() and the scene prep run once per scene on its owner instead of
    once per rank.  Two preallocated slots; the broadcast of scene s+1 is issued on a communication
    stream while the frames of scene s are synthesised, and a slot is refilled only after the
    events of its last users.  Every rank must call ``post`` for the same scenes in the same order.

    ``prepare(inputs, scene_out)`` fills a slot's scene buffer on the owner (default: JointSplat);
    ``broadcast(tensor, src)`` defaults to torch.distributed.broadcast -- both replaceable, which
    is how the gloo test drives this class on CPU tensors."""

    def __init__(self, C, H, W, n_tail, device, scene_numel, core_numel=None, stream=None, prepare=None, broadcast=None,
                 group=None):
        self.C, self.H, self.W, self.n_tail, self.device = C, H, W, n_tail, torch.device(device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.on_gpu = self.device.type == "cuda"
        self.comm = stream if stream is not None else (torch.cuda.Stream(device=self.device) if self.on_gpu else None)
        self.slots = [(torch.empty(scene_numel, dtype=torch.float32, device=self.device),
                       torch.empty(1, 2, H, W, dtype=torch.float32, device=self.device)) for _ in range(2)]
        self.core_numel = scene_numel if core_numel is None else core_numel     # leading elements that travel
        self.users = [[], []]          # events after which a slot may be refilled
        self.posted = 0
        self._prepare = prepare or self._prepare_with_joint_splat
        self._broadcast = broadcast or (lambda t, src: dist.broadcast(t, src=src, group=self.group))

    def _prepare_with_joint_splat(self, inputs, scene_out):
        from .synthesis import JointSplat
        feat, Z, motion = inputs[:3]
        tail = inputs[3] if len(inputs) > 3 else None
        JointSplat(feat, Z, motion, tail=tail, inputs_event=False, scene_buffer=scene_out).prepare_scene()

    def post(self, owner, inputs=None):
        """Start the exchange of the next scene: on `owner` prepare it from ``inputs`` =
        (features, Z, motion[, tail]) -- or a callable returning them, evaluated on the communication
        stream -- into the slot, then broadcast slot + motion.  Returns a ticket for ``take``."""
        slot = self.posted & 1
        self.posted += 1
        scene_buf, motion_buf = self.slots[slot]
        ctx = torch.cuda.stream(self.comm) if self.on_gpu else _null()
        with ctx:
            if self.on_gpu:
                for ev in self.users[slot]:
                    self.comm.wait_event(ev)
            self.users[slot] = []
            if self.rank == owner:
                assert inputs is not None, "the owner of a scene posts its inputs"
                if callable(inputs):        # e.g. the H2D copies of the inputs: queued on the communication stream too
                    inputs = inputs()
                self._prepare(inputs, scene_buf)
                motion_buf.copy_(inputs[2].reshape(1, 2, self.H, self.W))
            if self.world > 1:
                self._broadcast(scene_buf[:self.core_numel], owner)
                self._broadcast(motion_buf, owner)
            ready = None
            if self.on_gpu:
                ready = torch.cuda.Event()
                ready.record(self.comm)
        return slot, ready, owner

    def take(self, ticket):
        """The exchanged scene as (scene_buffer, motion, ready_event, core_only).  ``core_only``: this
        rank received only the first ``core_numel`` elements (JointSplat.from_scene_buffer rebuilds
        the rest).  Users of the buffers must be registered with ``used`` so that the slot is not
        refilled under them."""
        slot, ready, owner = ticket
        scene_buf, motion_buf = self.slots[slot]
        return scene_buf, motion_buf, ready, (self.rank != owner and self.core_numel < scene_buf.numel())

    def used(self, ticket, event):
        """The slot's contents are needed until `event`."""
        if event is not None:
            self.users[ticket[0]].append(event)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
