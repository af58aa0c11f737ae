"""CPU: the host-side orchestration of slr_sfs_b200.synthesis.JointSplat (batching, clip-table
caching, workspace ping-pong, per-frame calls) run against the CPU emulation of the library.

The product has no CPU mode: this test swaps, for its own duration only, the ctypes handle for
the emulation build (tests/emu) and torch.cuda's streams / events for inert stand-ins, so that
the very Python code that drives the B200 can be checked for index and bookkeeping errors in a
container without a GPU.  Asynchrony itself (stream ordering) is what tests/test_gpu_joint.py
covers on the device."""
import contextlib

import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err

emu = pytest.importorskip("emu", reason="tests/emu")
TOL = 1e-4


class _Stream:
    cuda_stream = 0

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def query(self):
        return True


@pytest.fixture()
def host_pkg(monkeypatch):
    import slr_sfs_b200
    from slr_sfs_b200 import _lib, synthesis
    calls = []
    real_call = _lib._plain_call

    def logged(name, *args):
        calls.append(name)
        return real_call(name, *args)

    monkeypatch.setattr(_lib, "_lib", emu.lib())
    monkeypatch.setattr(_lib, "_plain_call", logged)
    monkeypatch.setattr(_lib, "current_stream", lambda device: None)
    monkeypatch.setattr(synthesis, "_req", lambda t, name: t.contiguous())
    main = _Stream()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: main)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None, priority=0: _Stream())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setattr(synthesis.JointSplat, "_shared", {})
    slr_sfs_b200.calls = calls
    return slr_sfs_b200


def _scene(H, W, C, seed):
    from slr_sfs_b200 import workloads
    return workloads.scene(H, W, C, "A", seed=seed)


@pytest.mark.parametrize("pipeline", [True, False])
def test_frames_batches_cut_from_one_clip_table(host_pkg, pipeline):
    H, W, C, start, end = 24, 40, 4, 2, 13
    feat, Z, m = _scene(H, W, C, 5)
    js = host_pkg.JointSplat(feat, Z, m)
    js.batch, js.pipeline = 5, pipeline
    out = js.frames(start, end, start, end - start + 1).numpy()
    assert host_pkg.calls.count("slr_clip_table") == 1 and host_pkg.calls.count("slr_clip_bin") == 3
    for t in range(start, end + 1):
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (start, t, end))
        assert rel_err(out[t - start:t - start + 1], want) <= TOL, t
    # a sub-range of the cached table: no new table; another clip: a new one
    del host_pkg.calls[:]
    sub = js.frames(start, end, 6, 3).numpy()
    assert "slr_clip_table" not in host_pkg.calls and np.array_equal(sub, out[4:7])
    js.frames(start, end + 1, 6, 3)
    assert host_pkg.calls.count("slr_clip_table") == 1


def test_per_frame_loop_of_the_reference_integrates_the_clip_once(host_pkg):
    H, W, C, N = 17, 37, 3, 9
    feat, Z, m = _scene(H, W, C, 6)
    js = host_pkg.JointSplat(feat, Z, m, z_mode="v1")
    for t in range(N):                                   # test_v1_4eval_rawsize.py:233-239
        got = js.frame(torch.tensor([[0, t, N - 1]])).numpy()
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, t, N - 1), z_mode="v1")
        assert rel_err(got, want) <= TOL, t
    assert host_pkg.calls.count("slr_clip_table") == 1
    assert host_pkg.calls.count("slr_scene_prep") == 1


def test_prepare_clip_hint_and_two_layer_outputs(host_pkg):
    from slr_sfs_b200 import workloads
    H, W, C, N = 24, 40, 4, 6
    feat, Z, m = _scene(H, W, C, 7)
    a_f, a_bg = workloads.two_layer_extras(H, W, seed=7)
    A = torch.sigmoid(a_f) / torch.clamp(torch.sigmoid(a_f) + a_bg, min=1e-8)
    tail = torch.cat([a_f * A.exp(), A.exp()], 1).contiguous()
    js = host_pkg.JointSplat(feat, Z, m, tail=tail)
    js.batch = 4
    js.prepare_clip(0, N - 1)
    lo, hi = float(np.float32(1.0 / 600.0)), float(np.float32(599.0 / 600.0))
    for t0, n in ((0, 2), (2, 4)):
        gen, aux, mask = js.frames(0, N - 1, t0, n, want_aux=True, want_mask=True, alpha_clamp=(lo, hi))
        for i in range(n):
            w_gen, w_alpha, w_mask = oracle.joint_splat_2layer(feat.numpy(), Z.numpy(), a_f.numpy(), a_bg.numpy(),
                                                               m.numpy(), (0, t0 + i, N - 1), alpha0=True)
            assert rel_err(gen[i:i + 1].numpy(), w_gen) <= TOL
            alpha_fluid = aux[i:i + 1, 0:1] / torch.clamp(aux[i:i + 1, 1:2], min=1e-8)
            assert rel_err(alpha_fluid.numpy(), w_alpha) <= TOL
            assert np.array_equal(mask[i:i + 1].numpy(), w_mask)
    assert host_pkg.calls.count("slr_clip_table") == 1


def test_scatter_variant_through_the_python_layer(host_pkg):
    H, W, C, N = 17, 37, 3, 7
    feat, Z, m = _scene(H, W, C, 8)
    js = host_pkg.JointSplat(feat, Z, m)
    for t in (0, 3, N - 1):
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, t, N - 1))
        assert rel_err(js.frame_scatter((0, t, N - 1)).numpy(), want) <= TOL


def test_scene_and_table_buffers_are_recycled_between_scenes(host_pkg):
    """A new JointSplat per scene (what bench.py and a per-scene loop do) must not allocate: the
    big cross-stream buffers come back from the per-device pool, ordered by events."""
    import gc
    H, W, C, N = 17, 37, 3, 6
    feat, Z, m = _scene(H, W, C, 9)
    seen = []
    for _ in range(3):
        js = host_pkg.JointSplat(feat, Z, m)
        out = js.frames(0, N - 1, 0, N).numpy()
        seen.append((js._scene.data_ptr(), js._table["buf"].data_ptr()))
        pool = js._shared_state()["pool"]
        del js
        gc.collect()
        assert len(pool.free["scene"]) == 1 and len(pool.free["table"]) == 1
    assert seen[0] == seen[1] == seen[2]
    want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, 2, N - 1))
    assert rel_err(out[2:3], want) <= TOL
    # a second clip on the same object: the old table returns to the pool and is taken again
    js = host_pkg.JointSplat(feat, Z, m)
    js.frames(0, N - 1, 0, 2)
    first = js._table["buf"].data_ptr()
    js.frames(0, N, 0, 2)
    assert js._table["buf"].data_ptr() == first and len(js._pooled) == 2


def test_v2_importance_through_the_python_layer(host_pkg, golden_joint):
    """use_softmax_splatter_v2 (Z - maximum_warp_norm_splater(Z, forward flow), animating_softmax_splating.py:849-851)
    through JointSplat(z_mode='v2'): gather and scatter paths against the reference model's golden frames."""
    j = golden_joint
    N = int(j["N"])
    js = host_pkg.JointSplat(torch.from_numpy(j["feat"]), torch.from_numpy(j["Z"]), torch.from_numpy(j["motion"]), z_mode="v2")
    all_frames = js.frames(0, N - 1, 0, N).numpy()
    for t in (0, 3, N - 1):
        want = j[f"baseline/v2/t{t}/gen_fs"]
        assert rel_err(all_frames[t:t + 1], want) <= TOL, t
        assert rel_err(js.frame((0, t, N - 1)).numpy(), want) <= TOL, t
        assert rel_err(js.frame_scatter((0, t, N - 1)).numpy(), want) <= TOL, t
    assert host_pkg.calls.count("slr_clip_table") == 1 and host_pkg.calls.count("slr_maxwarpnorm") >= N
