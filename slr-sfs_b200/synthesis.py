"""The joint-splat block of the reference models as one fused operator.

What ``forward_flow`` does between the encoder and the decoder
(models/animating_softmax_splating.py:847-924; 2-layer variant
models/animating_softmax_splating_2layers_alpha_seperate.py:921-1045):

    D+ = euler_integration( flow, t - start)         D- = euler_integration(-flow, end - t + 1)
    Zn = Z - Z.max()                                  alpha = 1 - (t - start) / (end - start + 1)
    acc = splat([fs*e^Zn*alpha, e^Zn*alpha], D+) + splat([fs*e^Zn*(1-alpha), e^Zn*(1-alpha)], D-)
    gen_fs = acc[:, :-1] / clamp(acc[:, -1:], 1e-8)

The reference runs this as ~40 eager torch ops, two Euler integrations restarted
from zero and two cupy launches per frame.  Here a frame is a handful of launches
of the sm_100a library, nothing is synchronised with the host, and no
intermediate ``tenInput`` tensor is materialised.

Two algorithms, same results up to fp32 summation order:
  * ``frames`` / ``frame``   -- gather pipeline (csrc/clip_plan.cu, csrc/clip_gather.cu), the fast path;
  * ``frame_scatter``        -- atomic scatter + normalise (csrc/splat_ops.cu), the
                                design BASELINE.json sketches, kept as measured baseline.
"""
import os
import weakref

import numpy as np
import torch

from . import _lib

EPS = 1e-8


def blend_alpha(start, mid, end):
    """alpha of animating_softmax_splating.py:860, computed in fp32 like the reference."""
    a = np.float32(mid - start) / np.float32(end - start + 1)
    return float(np.float32(1.0) - a)


def _index_triplet(index):
    if torch.is_tensor(index):
        index = index.reshape(-1).tolist()
    start, mid, end = [int(v) for v in index]
    return start, mid, end


def _req(t, name):
    assert t.is_cuda, "%s must be a CUDA tensor (no CPU path)" % name
    assert t.dtype == torch.float32 and t.is_contiguous(), "%s must be contiguous fp32" % name
    return t


class _BufferPool:
    """Per device: the big buffers that one stream writes and another reads (scene buffer, clip
    table), recycled between JointSplat objects WITHOUT going through the caching allocator.  A
    block freed there after record_stream() is not handed out again until the recorded streams
    have drained, so a host that runs ahead of the GPU keeps getting fresh cudaMallocs -- each a
    device-wide synchronisation -- in the middle of the pipeline (measured: occasional bench runs
    35 % slower).  An entry remembers, per stream, the last event after which its contents are no
    longer needed; whoever takes it over makes its writing stream wait for those on the device."""
    # Free entries kept per kind.  Two is what the pipeline needs (the scene being gathered and the next one whose
    # table / scene buffer the side stream builds): the pool then reaches its steady state within two scenes.  With
    # more, a host that runs ahead of the GPU keeps GROWING the pool (every entry still busy -> a fresh 0.4-0.8 GB
    # cudaMalloc, a device-wide synchronisation) well into a run: measured as one bench run in five 25 % slow.
    KEEP = 2

    def __init__(self, device):
        self.device = device
        self.free = {}          # kind -> free entries, oldest first

    def acquire(self, kind, nbytes):
        """An entry whose previous users have finished if there is one (the side stream can then
        run ahead into the next scene), else a new one while fewer than KEEP are waiting, else the
        one released longest ago."""
        entries = self.free.setdefault(kind, [])
        fits = [e for e in entries if e["buf"].numel() * 4 >= nbytes]
        pick = next((e for e in fits if all(ev.query() for ev in e["last"].values())), None)
        if pick is None and fits and len(entries) >= self.KEEP:
            pick = fits[0]
        if pick is not None:
            entries.remove(pick)
            return pick
        # a fresh block from the caching allocator may still have queued users on the stream that
        # allocates it (the current one); whoever writes the entry first -- possibly on another
        # stream -- is ordered behind them like behind any previous user (take_over)
        buf = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
        return {"kind": kind, "last": {torch.cuda.current_stream(self.device): _event_now(self.device)}, "buf": buf}

    def release(self, entry):
        entries = self.free.setdefault(entry["kind"], [])
        entries.append(entry)
        if len(entries) > self.KEEP:
            drop = min(entries, key=lambda e: e["buf"].numel())     # the smallest (the oldest among equals)
            entries.remove(drop)
            for stream in drop["last"]:           # back to the allocator, with the streams that used it
                drop["buf"].record_stream(stream)

    @staticmethod
    def take_over(entry, writer):
        """`writer` is about to overwrite the entry: order it behind the previous contents' users."""
        for ev in entry["last"].values():
            writer.wait_event(ev)
        entry["last"] = {}

    @staticmethod
    def used(entry, stream, event):
        """The entry's contents are needed until `event` (recorded on `stream`)."""
        entry["last"][stream] = event


def _event_now(device):
    """An event recorded now on the current stream of `device`."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    return ev


def _release_all(pool, entries):
    for e in entries:
        pool.release(e)
    del entries[:]


class JointSplat:
    """Frame synthesiser for one scene: features [1,C,H,W], importance Z [1,1,H,W]
    and Eulerian motion [1,2,H,W] are fixed, frames t = start..end are produced on
    demand.  ``z_mode``: 'max' (Z - Z.max(), :855), 'v1' (Z as is, :853), 'v2' (use_softmax_splatter_v2,
    :849-851: Z - maximum_warp_norm_splater(Z, forward flow of the frame) -- the importance then depends
    on the frame, so the scene buffer is rebuilt per frame: correct, not fast; no shipped script sets it).

    ``tail`` ([1,n_tail,H,W], optional) carries the 2-layer model's extra
    pre-weighted channels (a_f*e^A, e^A with use_alpha0_as_blending_weight,
    2layers...py:967-972; a_f*e^Zn without, :974-976); they are blended with
    alpha / 1-alpha and splatted like the features and come back un-normalised
    positions C..C+n_tail-1 of the accumulator.
    """

    #: frames per slr_clip_bin / expand / gather launch (the Euler chains are integrated once per run of frames
    #: whatever the batch).  Measured with the direct index at 768x1024x64: 8 / 12 / 16 / 20 / 24 -> 7 569 / 7 765 /
    #: 7 834 / 7 893 / 7 861 frames/s (profiles/r02/batch_sweep.jsonl).  Capped so that one batch workspace stays
    #: below 15 % of the device memory (two exist: 1536x2048 then runs batches of 10).
    batch = int(os.environ.get("SLR_BATCH", "20"))
    #: overlap plan + expand of the next batch with the gather of the current one (two streams)
    pipeline = True

    # per device: the side stream, the two ping-pong workspaces and the events that say when a
    # workspace may be overwritten.  Shared by all JointSplat objects so that the index building
    # of the next scene overlaps the gather of the previous one.
    _shared = {}

    def __init__(self, features, Z, motion, z_mode="max", tail=None, inputs_event=None, scene_buffer=None):
        """``inputs_event``: a CUDA event after which the inputs are valid.  None (default): one is
        recorded now on the current stream (whatever produced the inputs was queued there);
        False: the inputs are already complete (lets the side stream start at once).
        ``scene_buffer``: caller-owned storage (``scene_buffer_numel`` floats) for the prepared scene
        instead of a pooled one -- e.g. a broadcast slot (sharding.SceneExchange)."""
        assert features.dim() == 4 and features.shape[0] == 1
        self.feat = _req(features.detach(), "features")
        self.C, self.H, self.W = features.shape[1:]
        self.Z = _req(Z.detach().reshape(1, 1, self.H, self.W), "Z")
        self.tail = None if tail is None else _req(tail.detach(), "tail")
        self.n_tail = 0 if tail is None else tail.shape[1]
        assert z_mode in ("max", "v1", "v2")
        self.z_mode = z_mode
        self._init_common(motion, inputs_event, scene_buffer)

    @staticmethod
    def scene_buffer_numel(C, n_tail, H, W):
        """fp32 elements of a prepared scene buffer (slr_scene_bytes / 4)."""
        return (_lib.load().slr_scene_bytes(C, n_tail, H, W) + 3) // 4

    @staticmethod
    def scene_core_numel(C, n_tail, H, W):
        """fp32 elements of the leading part of a scene buffer from which the rest can be rebuilt
        (slr_scene_core_bytes / 4): what has to travel when a prepared scene is broadcast."""
        return (_lib.load().slr_scene_core_bytes(C, n_tail, H, W) + 3) // 4

    @classmethod
    def from_scene_buffer(cls, scene_buffer, motion, C, H, W, n_tail=0, ready_event=None, core_only=False):
        """A synthesiser over an ALREADY PREPARED scene buffer (built by another JointSplat's
        ``prepare_scene`` -- possibly on another rank and broadcast): no features, no Z, no scene
        prep; only the gather path (``frames`` / ``frame``) is available.  ``ready_event``: after
        which the buffer and ``motion`` are valid (None: recorded now on the current stream;
        False: already complete).  ``core_only``: only the first ``scene_core_numel`` elements are
        valid (that is all a broadcast needs to carry); the staged copy behind them is rebuilt here
        (slr_scene_quilt) before the first frame."""
        self = cls.__new__(cls)
        self.feat = self.Z = self.tail = None
        self.C, self.H, self.W, self.n_tail = int(C), int(H), int(W), int(n_tail)
        self.z_mode = "v1"                      # no Z.max() to compute: e^(Z - max) is inside the buffer
        self._init_common(motion, ready_event, scene_buffer)
        if not core_only:
            self._scene_ready = self._inputs_ready or _event_now(self.device)
        return self

    def _init_common(self, motion, inputs_event, scene_buffer):
        self.motion = _req(motion.detach().reshape(1, 2, self.H, self.W), "motion")
        self.device = motion.device
        if inputs_event is None:
            with torch.cuda.device(self.device):
                inputs_event = torch.cuda.Event()
                inputs_event.record(torch.cuda.current_stream(self.device))
        self._inputs_ready = inputs_event or None
        self._zsub = None
        self._zsub_alloc = None        # event: the allocating stream's earlier work on the block is done
        self._zsub_ready = None        # event: Z.max() is in _zsub (computed once, never rewritten)
        self._scene = None
        self._scene_ready = None       # event: the scene buffer is built
        self._table = None             # cached clip table (see _clip_table)
        self._scene_entry = None       # pool entries behind _scene / _table["buf"]
        self._pooled = []              # everything to give back when this object dies
        self._finalizer = None
        if scene_buffer is not None:
            need = self.scene_buffer_numel(self.C, self.n_tail, self.H, self.W)
            assert scene_buffer.is_cuda and scene_buffer.dtype == torch.float32 and scene_buffer.is_contiguous() \
                and scene_buffer.numel() >= need and scene_buffer.data_ptr() % 16 == 0
            self._scene = scene_buffer
            # caller-owned: same bookkeeping as a pool entry (users' events), never handed to the pool
            self._scene_entry = {"kind": "external", "last": {}, "buf": scene_buffer}

    def prepare_scene(self):
        """Build Z.max() and the scene buffer NOW on the current stream (normally done lazily by the
        first ``frames`` call, on the side stream) and return the buffer -- what the owner of a scene
        broadcasts to the other ranks."""
        with torch.cuda.device(self.device):
            self._allocate(scene=True)
            self._prepare()
        return self._scene

    def _wait_inputs(self, stream):
        if self._inputs_ready is not None:
            stream.wait_event(self._inputs_ready)

    def _prepare(self):
        """Z.max() and the pre-weighted, channel-interleaved scene buffer: each built once, on the
        stream that first needs it; users on other streams are ordered behind its event.  Z.max()
        is never recomputed once valid (readers of `zsub` on other streams may still be queued)."""
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            s = _lib.current_stream(self.device)
            if self._zsub is not None:
                if self._zsub_ready is None:
                    self._wait_inputs(cur)
                    cur.wait_event(self._zsub_alloc)
                    self._zsub.record_stream(cur)
                    _lib.call("slr_reduce_max", _lib.ptr(self.Z), self.Z.numel(), _lib.ptr(self._zsub), s)
                    self._zsub_ready = _event_now(self.device)
                else:
                    cur.wait_event(self._zsub_ready)
                    self._zsub.record_stream(cur)
            if self._scene is not None:
                if self._scene_ready is None:
                    self._wait_inputs(cur)
                    _BufferPool.take_over(self._scene_entry, cur)
                    if self.feat is None:       # a received scene core (from_scene_buffer(core_only=True))
                        if os.environ.get("SLR_GATHER_MODE") == "staged":      # only the staged gather reads the quilted copy
                            _lib.call("slr_scene_quilt", _lib.ptr(self._scene), self.C, self.n_tail, self.H, self.W, s)
                    else:
                        _lib.call("slr_scene_prep", _lib.ptr(self.feat), _lib.ptr(self.Z), _lib.ptr(self._zsub),
                                  _lib.ptr(self.tail), self.n_tail, _lib.ptr(self._scene), self.C, self.H, self.W, s)
                    self._scene_ready = _event_now(self.device)
                    _BufferPool.used(self._scene_entry, cur, self._scene_ready)
                else:
                    cur.wait_event(self._scene_ready)

    def _allocate(self, scene):
        """Allocate what _prepare fills."""
        if self.z_mode == "max" and self._zsub is None:
            self._zsub = torch.empty(1, dtype=torch.float32, device=self.device)
            self._zsub_alloc = _event_now(self.device)
        if scene and self._scene is None:
            n = _lib.load().slr_scene_bytes(self.C, self.n_tail, self.H, self.W)
            self._scene_entry = self._from_pool("scene", n)
            self._scene = self._scene_entry["buf"]

    def _from_pool(self, kind, nbytes):
        pool = self._shared_state()["pool"]
        entry = pool.acquire(kind, nbytes)
        self._pooled.append(entry)
        if self._finalizer is None:
            self._finalizer = weakref.finalize(self, _release_all, pool, self._pooled)
        return entry

    @property
    def zsub(self):
        """Device scalar Z.max() (or None for z_mode 'v1'), valid on the current stream."""
        with torch.cuda.device(self.device):
            self._allocate(scene=False)
            self._prepare()
        return self._zsub

    def _shared_state(self):
        st = JointSplat._shared.get(self.device)
        if st is None:
            st = JointSplat._shared[self.device] = {"side": torch.cuda.Stream(device=self.device),
                                                    "ws": {}, "free": {}, "turn": 0,
                                                    "pool": _BufferPool(self.device)}
        return st

    def _fitting_batch(self):
        """``self.batch``, lowered where one batch workspace would exceed 15 % of the device memory."""
        batch = max(1, int(self.batch))
        try:
            total = torch.cuda.get_device_properties(self.device).total_memory
        except Exception:          # not a CUDA device (the emulated library in tests/)
            return batch
        per_frame = _lib.load().slr_clip_workspace_bytes(self.H, self.W, 1)
        return max(1, min(batch, int(0.15 * total // max(per_frame, 1))))

    def _scratch(self, st, n, slot, side):
        need = _lib.load().slr_clip_workspace_bytes(self.H, self.W, n)
        ws = st["ws"].get(slot)
        if ws is None or ws.numel() * 4 < need:
            # allocated by the current (main) stream and marked as used by the side stream, so the
            # caching allocator does not hand the old block out while queued work still uses it
            ws = st["ws"][slot] = torch.empty((need + 3) // 4, dtype=torch.float32, device=self.device)
            ws.record_stream(side)
            side.wait_event(_event_now(self.device))     # behind whatever the allocating stream still had queued on the block
        return ws, ws.numel() * 4

    def _clip_table(self, start, end, t0, n, side):
        """The motion-only part of the plan (slr_clip_table: both Euler chains, landing coordinates,
        bin sizes) for frames t0 .. t0+n-1 of the clip [start, end], built once on `side` and cached:
        later calls for any sub-range of it only cut their batches from it (slr_clip_bin)."""
        tb = self._table
        mode = os.environ.get("SLR_GATHER_MODE", "ldg")      # what a table holds depends on how the lists are built
        if tb is not None and tb["clip"] == (start, end) and tb["t0"] <= t0 and t0 + n <= tb["t0"] + tb["n"] \
                and tb["mode"] == mode:
            return tb
        nbytes = _lib.load().slr_clip_table_bytes(self.H, self.W, n)
        if tb is not None:                      # the old table goes back to the pool right away
            self._pooled.remove(tb["entry"])
            self._shared_state()["pool"].release(tb["entry"])
            self._table = None
        entry = self._from_pool("table", nbytes)
        buf = entry["buf"]
        with torch.cuda.stream(side):
            self._wait_inputs(side)
            _BufferPool.take_over(entry, side)
            _lib.call("slr_clip_table", _lib.ptr(self.motion), self.H, self.W, start, end, t0, n,
                      _lib.ptr(buf), nbytes, _lib.current_stream(self.device))
            ready = torch.cuda.Event()
            ready.record(side)
        _BufferPool.used(entry, side, ready)
        tb = self._table = {"clip": (start, end), "t0": t0, "n": n, "buf": buf, "bytes": nbytes,
                            "ready": ready, "entry": entry, "mode": mode}
        return tb

    def prepare_clip(self, start, end, t0=None, n=None):
        """Optional hint: the caller is going to ask for frames t0 .. t0+n-1 (default: the whole
        clip) of [start, end], in any number of frames() / frame() calls.  Builds the clip table
        for that range now, so that every Euler chain is integrated once for the range instead of
        once per call."""
        t0 = start if t0 is None else t0
        n = end - t0 + 1 if n is None else n
        with torch.cuda.device(self.device):
            st = self._shared_state()
            side = st["side"] if self.pipeline else torch.cuda.current_stream(self.device)
            self._clip_table(start, end, t0, n, side)

    def frames(self, start, end, t0, n, out=None, want_aux=False, want_mask=False, alpha_clamp=(0.0, 1.0), want_nnz=False):
        """Frames t0..t0+n-1 of the clip [start, end]: gen_fs [n,C,H,W]
        (+ aux [n,n_tail+1,H,W] raw tail/norm sums, + mask [n,1,H,W], + nnz [n,1,H,W]: the number of
        non-zero channels of gen_fs per pixel, see decoder_entry).  Asynchronous like any
        torch op: results are ordered on the current stream.

        The Euler chains of the requested range are integrated once (slr_clip_table, cached: see
        prepare_clip); the rest is issued in batches of ``self.batch`` frames.  With
        ``self.pipeline`` the latency-bound part (scene prep, table, slr_clip_bin, slr_clip_expand)
        goes to a side stream with its own workspace and runs beside the bandwidth-bound gather of
        the previous batch -- of this call or of an earlier one (the previous scene)."""
        H, W, C = self.H, self.W, self.C
        if out is None:
            out = torch.empty(n, C, H, W, dtype=torch.float32, device=self.device)
        assert out.shape == (n, C, H, W) and out.is_contiguous() and out.device == self.device
        if self.z_mode == "v2":
            return self._frames_v2(start, end, t0, n, out, want_aux, want_mask, alpha_clamp, want_nnz)
        aux = torch.empty(n, self.n_tail + 1, H, W, dtype=torch.float32, device=self.device) if want_aux else None
        mask = torch.empty(n, 1, H, W, dtype=torch.float32, device=self.device) if want_mask else None
        nnz = torch.empty(n, 1, H, W, dtype=torch.float32, device=self.device) if want_nnz else None
        batch = self._fitting_batch()
        batch = -(-n // -(-n // batch))          # equal batches: 30 frames at batch 20 are 15 + 15, not 20 + 10
        batches = [(b0, min(batch, n - b0)) for b0 in range(0, n, batch)]
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            st = self._shared_state()
            two_streams = bool(self.pipeline)
            side = st["side"] if two_streams else main
            self._allocate(scene=True)
            scene = self._scene
            tb = self._clip_table(start, end, t0, n, side)
            for (b0, nb) in batches:
                slot = st["turn"] = st["turn"] ^ 1
                args = (C, self.n_tail, H, W, start, end, t0 + b0, nb, alpha_clamp[0], alpha_clamp[1])
                ws, ws_bytes = self._scratch(st, batch, slot, side)
                with torch.cuda.stream(side):
                    for ev in st["free"].get(slot, ()):
                        side.wait_event(ev)
                    side.wait_event(tb["ready"])
                    self._prepare()
                    s = _lib.current_stream(self.device)
                    _lib.call("slr_clip_bin", _lib.ptr(tb["buf"]), tb["bytes"], H, W, tb["n"], t0 + b0 - tb["t0"], nb,
                              _lib.ptr(ws), ws_bytes, s)
                    _lib.call("slr_clip_expand", _lib.ptr(scene), _lib.ptr(self.motion), *args,
                              _lib.ptr(ws), ws_bytes, s)
                    if two_streams:
                        ready = torch.cuda.Event()
                        ready.record(side)
                if two_streams:
                    main.wait_event(ready)
                for entry in ("slr_clip_gather", "slr_clip_heavy"):
                    _lib.call(entry, _lib.ptr(scene), _lib.ptr(self.motion), *args, _lib.ptr(out[b0:]),
                              None if aux is None else _lib.ptr(aux[b0:]),
                              None if mask is None else _lib.ptr(mask[b0:]),
                              None if nnz is None else _lib.ptr(nnz[b0:]), _lib.ptr(ws), ws_bytes,
                              _lib.current_stream(self.device))
                done = torch.cuda.Event()
                done.record(main)
                st["free"][slot] = (done,)
                # the batch's side-stream work precedes `ready`, which main waited for: `done` covers both
                _BufferPool.used(self._scene_entry, main, done)
                _BufferPool.used(tb["entry"], main, done)
        res = (out,) + ((aux,) if want_aux else ()) + ((mask,) if want_mask else ()) + ((nnz,) if want_nnz else ())
        return res if len(res) > 1 else out

    def _z_minus_warped_max(self, steps):
        """Z - maximum_warp_norm_splater(Z, euler_integration(motion, steps)) on the current stream
        (animating_softmax_splating.py:849-851; softsplat.py:576-624)."""
        H, W, dev = self.H, self.W, self.device
        disp = torch.empty(1, 2, H, W, dtype=torch.float32, device=dev)
        scratch = torch.empty(1, 1, H, W, dtype=torch.float32, device=dev)
        zmax = torch.empty(1, 1, H, W, dtype=torch.float32, device=dev)
        s = _lib.current_stream(dev)
        _lib.call("slr_euler", _lib.ptr(self.motion), 1.0, steps, _lib.ptr(disp), None, H, W, s)
        _lib.call("slr_maxwarpnorm", _lib.ptr(self.Z), _lib.ptr(disp), _lib.ptr(scratch), _lib.ptr(zmax), 1, 1, H, W, s)
        return self.Z - zmax

    def _frames_v2(self, start, end, t0, n, out, want_aux, want_mask, alpha_clamp, want_nnz):
        """use_softmax_splatter_v2: the scene buffer depends on the frame (see the class docstring).
        Frame by frame on the current stream: warped max of Z, scene prep, one-frame batch cut from
        the shared clip table."""
        H, W, C, dev = self.H, self.W, self.C, self.device
        aux = torch.empty(n, self.n_tail + 1, H, W, dtype=torch.float32, device=dev) if want_aux else None
        mask = torch.empty(n, 1, H, W, dtype=torch.float32, device=dev) if want_mask else None
        nnz = torch.empty(n, 1, H, W, dtype=torch.float32, device=dev) if want_nnz else None
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            st = self._shared_state()
            self._wait_inputs(main)
            if self._scene is None:
                nbytes = _lib.load().slr_scene_bytes(C, self.n_tail, H, W)
                self._scene_entry = self._from_pool("scene", nbytes)
                self._scene = self._scene_entry["buf"]
            _BufferPool.take_over(self._scene_entry, main)
            tb = self._clip_table(start, end, t0, n, main)
            main.wait_event(tb["ready"])
            ws, ws_bytes = self._scratch(st, 1, "v2", main)
            for ev in st["free"].get("v2", ()):
                main.wait_event(ev)
            s = _lib.current_stream(dev)
            for i in range(n):
                t = t0 + i
                z_eff = self._z_minus_warped_max(t - start)
                _lib.call("slr_scene_prep", _lib.ptr(self.feat), _lib.ptr(z_eff), None, _lib.ptr(self.tail), self.n_tail,
                          _lib.ptr(self._scene), C, H, W, s)
                args = (C, self.n_tail, H, W, start, end, t, 1, alpha_clamp[0], alpha_clamp[1])
                _lib.call("slr_clip_bin", _lib.ptr(tb["buf"]), tb["bytes"], H, W, tb["n"], t - tb["t0"], 1, _lib.ptr(ws), ws_bytes, s)
                _lib.call("slr_clip_expand", _lib.ptr(self._scene), _lib.ptr(self.motion), *args, _lib.ptr(ws), ws_bytes, s)
                for entry in ("slr_clip_gather", "slr_clip_heavy"):
                    _lib.call(entry, _lib.ptr(self._scene), _lib.ptr(self.motion), *args, _lib.ptr(out[i:]),
                              None if aux is None else _lib.ptr(aux[i:]), None if mask is None else _lib.ptr(mask[i:]),
                              None if nnz is None else _lib.ptr(nnz[i:]), _lib.ptr(ws), ws_bytes, s)
            done = _event_now(dev)
            st["free"]["v2"] = (done,)
            _BufferPool.used(self._scene_entry, main, done)
            _BufferPool.used(tb["entry"], main, done)
        res = (out,) + ((aux,) if want_aux else ()) + ((mask,) if want_mask else ()) + ((nnz,) if want_nnz else ())
        return res if len(res) > 1 else out

    def frame(self, index, **kw):
        """gen_fs [1,C,H,W] for index = (start, t, end) -- the per-frame call of the
        reference's loop (test_v1_4eval_rawsize.py:233-239)."""
        start, mid, end = _index_triplet(index)
        tb = self._table
        if start <= mid <= end and (tb is None or tb["clip"] != (start, end)):
            # the reference's loop asks for every t of the clip, one call each: integrate the
            # chains for the whole clip on the first call instead of restarting them per frame
            self.prepare_clip(start, end)
        return self.frames(start, end, mid, 1, **kw)

    # -- scatter variant: Euler x2 -> atomic scatter of both directions -> normalise
    def accumulate_scatter(self, index, alpha=None):
        assert self.feat is not None, "the scatter variant needs the features (not available from_scene_buffer)"
        start, mid, end = _index_triplet(index)
        if alpha is None:
            alpha = blend_alpha(start, mid, end)
        H, W, C = self.H, self.W, self.C
        dev = self.device
        disp = torch.empty(2, 2, H, W, dtype=torch.float32, device=dev)
        acc = torch.empty(1, C + self.n_tail + 1, H, W, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            s = _lib.current_stream(dev)
            _lib.call("slr_euler", _lib.ptr(self.motion), 1.0, mid - start, _lib.ptr(disp[0]), None, H, W, s)
            _lib.call("slr_euler", _lib.ptr(self.motion), -1.0, end - mid + 1, _lib.ptr(disp[1]), None, H, W, s)
            z, zsub = self.Z, self.zsub
            if self.z_mode == "v2":
                self._wait_inputs(torch.cuda.current_stream(dev))
                z, zsub = self._z_minus_warped_max(mid - start), None
            _lib.call("slr_joint_scatter", _lib.ptr(self.feat), _lib.ptr(z), _lib.ptr(zsub),
                      _lib.ptr(self.tail), self.n_tail, _lib.ptr(disp[0]), _lib.ptr(disp[1]), alpha,
                      _lib.ptr(acc), C, H, W, s)
        return acc

    def normalize(self, acc, n_out=None, norm_ch=None, want_mask=False):
        n_acc = acc.shape[1]
        n_out = self.C if n_out is None else n_out
        norm_ch = n_acc - 1 if norm_ch is None else norm_ch
        out = torch.empty(1, n_out, self.H, self.W, dtype=torch.float32, device=self.device)
        mask = torch.empty(1, 1, self.H, self.W, dtype=torch.float32, device=self.device) if want_mask else None
        with torch.cuda.device(self.device):
            _lib.call("slr_normalize", _lib.ptr(acc), _lib.ptr(out), _lib.ptr(mask), n_out, norm_ch, n_acc,
                      EPS, self.H, self.W, _lib.current_stream(self.device))
        return (out, mask) if want_mask else out

    def frame_scatter(self, index, alpha=None):
        """gen_fs [1,C,H,W] for index = (start, t, end)."""
        return self.normalize(self.accumulate_scatter(index, alpha))


def warp_flow_block(image, forward_flow, backward_flow, index):
    """The RGB twin of the joint block: ``AnimatingSoftmaxSplating.warp_flow``
    (models/animating_softmax_splating.py:1064-1138).  ``image`` [1,3,H,W] is splatted with the
    PRECOMPUTED displacement fields ``forward_flow`` = flow_f[mid - start] and ``backward_flow`` =
    flow_p[end - mid] ([1,2,H,W] each), Z = 1 everywhere:

        alpha = 1 - (mid - start) / (end - start)                       no "+ 1" here (:1064)
        acc   = splat([img * e^(Z - Z.max()) * alpha, e^(Z - Z.max()) * alpha], forward_flow)
              + splat([img * e^Z * (1 - alpha),        e^Z * (1 - alpha)],        backward_flow)   (:1067, :1103)
        out   = acc[:, :3] / clamp(acc[:, 3:], 1e-8)

    The forward direction carries e^0, the backward one e^1 -- the reference's own asymmetry, kept.
    One fused scatter (slr_joint_scatter_weights) + slr_normalize; returns PredImg [1,3,H,W]."""
    start, mid, end = _index_triplet(index)
    image = _req(image.detach(), "image")
    assert image.dim() == 4 and image.shape[0] == 1
    C, H, W = image.shape[1:]
    fwd = _req(forward_flow.detach().reshape(2, H, W), "forward_flow")
    bwd = _req(backward_flow.detach().reshape(2, H, W), "backward_flow")
    alpha = np.float32(1.0) - np.float32(mid - start) / np.float32(end - start)
    w_fwd = float(alpha)                                                  # e^(1 - 1) * alpha
    w_bwd = float(np.exp(np.float32(1.0)) * (np.float32(1.0) - alpha))    # e^1 * (1 - alpha)
    dev = image.device
    ones = torch.ones(1, 1, H, W, dtype=torch.float32, device=dev)
    zmax = torch.ones(1, dtype=torch.float32, device=dev)
    acc = torch.empty(1, C + 1, H, W, dtype=torch.float32, device=dev)
    out = torch.empty(1, C, H, W, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        s = _lib.current_stream(dev)
        _lib.call("slr_joint_scatter_weights", _lib.ptr(image), _lib.ptr(ones), _lib.ptr(zmax), None, 0,
                  _lib.ptr(fwd), _lib.ptr(bwd), w_fwd, w_bwd, _lib.ptr(acc), C, H, W, s)
        _lib.call("slr_normalize", _lib.ptr(acc), _lib.ptr(out), None, C, C, C + 1, EPS, H, W, s)
    return out
