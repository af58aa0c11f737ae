"""GPU: the joint block (Euler -> forward+backward splat -> normalise) against the
reference models' golden outputs and against the oracle, small and full size."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200
    return slr_sfs_b200


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def paths(js):
    out = [("scatter", js.frame_scatter)]
    if hasattr(js, "frame"):
        out.append(("gather", js.frame))
    return out


def test_baseline_joint_vs_reference_model_golden(pkg, golden_joint):
    j = golden_joint
    N = int(j["N"])
    for z_mode in ("max", "v1", "v2"):
        js = pkg.JointSplat(cu(j["feat"]), cu(j["Z"]), cu(j["motion"]), z_mode=z_mode)
        for t in (0, 3, N - 1):
            want = j[f"baseline/{z_mode}/t{t}/gen_fs"]
            for name, fn in paths(js):
                got = fn((0, t, N - 1)).cpu().numpy()
                assert rel_err(got, want) <= TOL, (z_mode, t, name)
                assert np.all(got[want == 0.0] == 0.0)       # holes stay exactly zero


def test_two_layer_joint_vs_reference_model_golden(pkg, golden_joint):
    j = golden_joint
    N = int(j["N"])
    ao = torch.from_numpy(j["alpha_encoder_out"]).cuda()
    a_bg = torch.sigmoid(ao[:, 0:1])
    a_f = ao[:, 1:2]
    Z = cu(j["Z"])
    for alpha0 in (True, False):
        if alpha0:
            A = torch.sigmoid(a_f) / torch.clamp(torch.sigmoid(a_f) + a_bg, min=1e-8)
            tail = torch.cat([a_f * A.exp(), A.exp()], 1).contiguous()
        else:
            tail = (a_f * (Z - Z.max()).exp()).contiguous()
        js = pkg.JointSplat(cu(j["feat"]), Z, cu(j["motion"]), tail=tail)
        C = js.C
        for t in (0, 3, N - 1):
            tag = f"twolayer/{'alpha0' if alpha0 else 'plain'}/t{t}"
            a = pkg.synthesis.blend_alpha(0, t, N - 1)
            a = min(max(a, float(np.float32(1.0 / 600.0))), float(np.float32(599.0 / 600.0)))   # 2layers...py:952
            acc = js.accumulate_scatter((0, t, N - 1), alpha=a)
            gen = js.normalize(acc).cpu().numpy()
            if alpha0:
                alpha_fluid = (acc[:, C:C + 1] / torch.clamp(acc[:, C + 1:C + 2], min=1e-8)).cpu().numpy()
            else:
                alpha_fluid = (acc[:, C:C + 1] / torch.clamp(acc[:, C + 1:C + 2], min=1e-8)).cpu().numpy()
            assert rel_err(gen, j[f"{tag}/gen_fs"]) <= TOL, tag
            assert rel_err(alpha_fluid, j[f"{tag}/alpha_fluid"]) <= TOL, tag


@pytest.mark.parametrize("motion", ["A", "B", "C"])
def test_joint_vs_oracle_medium(pkg, motion):
    from slr_sfs_b200 import workloads
    H, W, C, N = 96, 128, 16, 20
    feat, Z, m = workloads.scene(H, W, C, motion, seed=3)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    for t in (0, 7, N - 1):
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, t, N - 1))
        for name, fn in paths(js):
            got = fn((0, t, N - 1)).cpu().numpy()
            assert rel_err(got, want) <= TOL, (motion, t, name)


def test_joint_full_size_768x1024(pkg):
    """BASELINE.json configs[1] shape: 768x1024, 64 channels, N=60; two frames against the oracle."""
    from slr_sfs_b200 import workloads
    H, W, C, N = 768, 1024, 64, 60
    feat, Z, m = workloads.scene(H, W, C, "A", seed=0)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    for t in (1, 40):
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, t, N - 1))
        for name, fn in paths(js):
            got = fn((0, t, N - 1)).cpu().numpy()
            assert rel_err(got, want) <= TOL, (t, name)
            assert np.array_equal(got == 0.0, want == 0.0) or np.mean((got == 0.0) != (want == 0.0)) < 1e-6


def test_blend_is_convex_combination_property(pkg):
    """Size-independent property: with constant features the normalised output is that
    constant wherever anything landed (softmax weights cancel), 0 in holes."""
    H, W, C, N = 768, 1024, 8, 60
    from slr_sfs_b200 import workloads
    _, Z, m = workloads.scene(H, W, C, "A", seed=2)
    feat = torch.full((1, C, H, W), 0.75)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    for name, fn in paths(js):
        out = fn((0, 30, N - 1))
        covered = out[:, 0] != 0
        assert covered.float().mean() > 0.5
        vals = out[:, :, covered[0]]
        close = torch.isclose(vals, torch.tensor(0.75, device="cuda"), rtol=1e-5, atol=0)
        # cells whose total weight is below the 1e-8 clamp legitimately come out smaller
        assert close.float().mean() > 0.9999 and bool((vals <= 0.75 * (1 + 1e-5)).all()), name


# ---------------------------------------------------------------------------
# gather pipeline specifics
# ---------------------------------------------------------------------------
def test_two_layer_gather_aux_and_mask(pkg, golden_joint):
    j = golden_joint
    N = int(j["N"])
    ao = torch.from_numpy(j["alpha_encoder_out"]).cuda()
    a_bg, a_f = torch.sigmoid(ao[:, 0:1]), ao[:, 1:2]
    Z = cu(j["Z"])
    for alpha0 in (True, False):
        if alpha0:
            A = torch.sigmoid(a_f) / torch.clamp(torch.sigmoid(a_f) + a_bg, min=1e-8)
            tail = torch.cat([a_f * A.exp(), A.exp()], 1).contiguous()
        else:
            tail = (a_f * (Z - Z.max()).exp()).contiguous()
        js = pkg.JointSplat(cu(j["feat"]), Z, cu(j["motion"]), tail=tail)
        clamp = (float(np.float32(1.0 / 600.0)), float(np.float32(599.0 / 600.0)))     # 2layers...py:952
        gen, aux, mask = js.frames(0, N - 1, 0, N, want_aux=True, want_mask=True, alpha_clamp=clamp)
        for t in (0, 3, N - 1):
            tag = f"twolayer/{'alpha0' if alpha0 else 'plain'}/t{t}"
            alpha_fluid = aux[t:t + 1, 0:1] / torch.clamp(aux[t:t + 1, 1:2], min=1e-8)
            assert rel_err(gen[t:t + 1].cpu().numpy(), j[f"{tag}/gen_fs"]) <= TOL, tag
            assert rel_err(alpha_fluid.cpu().numpy(), j[f"{tag}/alpha_fluid"]) <= TOL, tag
            norm = aux[t:t + 1, -1:]
            assert torch.equal(mask[t:t + 1], (norm > 1e-8).float())


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 7), (4, 8, 32), (5, 9, 33), (7, 31, 65), (64, 40, 100)])
@pytest.mark.parametrize("motion", ["A", "B"])
def test_gather_ragged_shapes(pkg, shape, motion):
    from slr_sfs_b200 import workloads
    C, H, W = shape
    N = 9
    feat, Z, m = workloads.scene(H, W, C, motion, seed=H + W)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    out = js.frames(0, N - 1, 0, N).cpu().numpy()
    for t in (0, 4, N - 1):
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, t, N - 1))
        assert rel_err(out[t:t + 1], want) <= TOL, (shape, motion, t)
        assert np.all(out[t:t + 1][want == 0.0] == 0.0)


def test_gather_sink_and_compression_use_heavy_tiles(pkg):
    """Flows that pile many sources onto few destinations overflow the per-lane lists: the
    tail loop (deep lists) and the heavy-tile path (L2 reductions) must give the same sums."""
    H, W, C, N = 40, 72, 6, 3
    rng = np.random.default_rng(9)
    feat = rng.standard_normal((1, C, H, W)).astype(np.float32)
    Z = rng.standard_normal((1, 1, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    sink = np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys])[None].astype(np.float32)      # everything -> one point in 1 step
    squeeze = np.stack([-(xs - W / 2) * 0.45, -(ys - H / 2) * 0.2])[None].astype(np.float32)  # strong compression
    for m in (sink, squeeze):
        js = pkg.JointSplat(cu(feat), cu(Z), cu(m))
        for t in (1, 2):
            want = oracle.joint_splat_baseline(feat, Z, m, (0, t, N - 1))
            got = js.frame((0, t, N - 1)).cpu().numpy()
            assert rel_err(got, want) <= TOL
            assert np.all(got[want == 0.0] == 0.0)


def test_frames_batching_and_nonzero_start(pkg):
    from slr_sfs_b200 import workloads
    H, W, C = 64, 96, 8
    feat, Z, m = workloads.scene(H, W, C, "A", seed=5)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    start, end = 2, 13
    js.batch = 5
    all_frames = js.frames(start, end, start, end - start + 1)
    js.batch = 1
    for t in range(start, end + 1):
        one = js.frame((start, t, end))
        assert rel_err(all_frames[t - start:t - start + 1].cpu().numpy(), one.cpu().numpy()) <= 1e-5
        want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (start, t, end))
        assert rel_err(one.cpu().numpy(), want) <= TOL, t


def test_gather_matches_scatter_full_size_stress_motion(pkg):
    """768x1024x64 with i.i.d. U(-8,8) motion (worst-case incoherent scatter): the two
    algorithms agree; mass is conserved (sum of norm-weighted output == sum of inputs that land)."""
    from slr_sfs_b200 import workloads
    H, W, C, N = 768, 1024, 64, 60
    feat, Z, m = workloads.scene(H, W, C, "B", seed=1)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    for t in (3, 57):
        a = js.frame_scatter((0, t, N - 1))
        b = js.frame((0, t, N - 1))
        assert rel_err(b.cpu().numpy(), a.cpu().numpy()) <= TOL


def test_static_pixels_are_implicit_self_contributions(pkg):
    """Pixels with exactly zero motion are not binned; the gather adds their self-contribution.
    All-static scene: output == features (weights cancel); half-static scene with sources
    flowing onto static pixels: matches the oracle."""
    H, W, C, N = 40, 70, 8, 12
    rng = np.random.default_rng(21)
    feat = rng.standard_normal((1, C, H, W)).astype(np.float32)
    Z = rng.standard_normal((1, 1, H, W)).astype(np.float32)
    zero = np.zeros((1, 2, H, W), np.float32)
    js = pkg.JointSplat(cu(feat), cu(Z), cu(zero))
    out = js.frames(0, N - 1, 0, N).cpu().numpy()
    for t in (0, 5, N - 1):
        assert rel_err(out[t:t + 1], feat) <= 1e-6
    m = zero.copy()
    m[0, 0, :, : W // 2] = 1.7          # left half streams right, onto the static right half
    m[0, 1, : H // 2, : W // 2] = -0.6
    m[0, 0, 3, 5] = -0.0                # negative zero is static too
    m[0, 1, 3, 5] = 0.0
    js = pkg.JointSplat(cu(feat), cu(Z), cu(m))
    for t in (1, 6, N - 1):
        want = oracle.joint_splat_baseline(feat, Z, m, (0, t, N - 1))
        assert rel_err(js.frame((0, t, N - 1)).cpu().numpy(), want) <= TOL
        assert rel_err(js.frame_scatter((0, t, N - 1)).cpu().numpy(), want) <= TOL


def test_pipelined_and_single_stream_frames_agree(pkg):
    from slr_sfs_b200 import workloads
    H, W, C, N = 72, 100, 8, 14
    feat, Z, m = workloads.scene(H, W, C, "A", seed=8)
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda())
    js.batch = 3
    js.pipeline = True
    a = js.frames(0, N - 1, 0, N)
    js.pipeline = False
    b = js.frames(0, N - 1, 0, N)
    torch.cuda.synchronize()
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5
    want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), m.numpy(), (0, 9, N - 1))
    assert rel_err(a[9:10].cpu().numpy(), want) <= TOL


def test_c_abi_clip_frames_single_call(pkg):
    """slr_clip_frames (plan + expand + gather + heavy in one C call) against the oracle."""
    from slr_sfs_b200 import _lib, workloads
    H, W, C, N = 56, 80, 12, 10
    feat, Z, m = workloads.scene(H, W, C, "A", seed=12)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    lib = _lib.load()
    s = _lib.current_stream(feat.device)
    zmax = torch.empty(1, device="cuda")
    scene = torch.empty(lib.slr_scene_bytes(C, 0, H, W) // 4, device="cuda")
    ws_bytes = lib.slr_clip_workspace_bytes(H, W, 4)
    ws = torch.empty((ws_bytes + 3) // 4, device="cuda")
    out = torch.empty(4, C, H, W, device="cuda")
    mask = torch.empty(4, 1, H, W, device="cuda")
    nnz = torch.empty(4, 1, H, W, device="cuda")
    _lib.call("slr_reduce_max", _lib.ptr(Z), Z.numel(), _lib.ptr(zmax), s)
    _lib.call("slr_scene_prep", _lib.ptr(feat), _lib.ptr(Z), _lib.ptr(zmax), None, 0, _lib.ptr(scene), C, H, W, s)
    _lib.call("slr_clip_frames", _lib.ptr(scene), _lib.ptr(m), C, 0, H, W, 0, N - 1, 3, 4, 0.0, 1.0,
              _lib.ptr(out), None, _lib.ptr(mask), _lib.ptr(nnz), _lib.ptr(ws), ws_bytes, s)
    torch.cuda.synchronize()
    for i, t in enumerate(range(3, 7)):
        want = oracle.joint_splat_baseline(feat.cpu().numpy(), Z.cpu().numpy(), m.cpu().numpy(), (0, t, N - 1))
        got = out[i:i + 1].cpu().numpy()
        assert rel_err(got, want) <= TOL
        # the mask is exactly the decoder's hole test (networks/architectures.py:369) on the norm
        covered = (np.abs(want).sum(1, keepdims=True) != 0)
        assert np.mean(mask[i:i + 1].cpu().numpy().astype(bool) != covered) < 1e-3
        # nnz is exactly the channel sum of the decoder's per-element mask (x != 0) on what was written
        assert torch.equal(nnz[i:i + 1], (out[i:i + 1] != 0).float().sum(1, keepdim=True))
    with pytest.raises(_lib.SlrError):          # workspace too small is an argument error, not a crash
        _lib.call("slr_clip_frames", _lib.ptr(scene), _lib.ptr(m), C, 0, H, W, 0, N - 1, 3, 4, 0.0, 1.0,
                  _lib.ptr(out), None, None, None, _lib.ptr(ws), 1024, s)


@pytest.mark.parametrize("motion", ["A", "B"])
def test_staged_gather_mode_full_size(pkg, motion, monkeypatch):
    """SLR_GATHER_MODE=staged (sources staged in shared memory by TMA bulk copies, csrc/clip_gather.cu:
    stagegather_kernel; the L1 gather takes the tiles that do not fit) against the default L1 gather and
    the oracle at 768x1024x64 -- smooth flow (99 % of the tiles staged) and incoherent flow (mostly fallback)."""
    from slr_sfs_b200 import workloads
    H, W, C, N = 768, 1024, 64, 60
    feat, Z, m = workloads.scene(H, W, C, motion, seed=2)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    monkeypatch.setenv("SLR_GATHER_MODE", "ldg")
    base = pkg.JointSplat(feat, Z, m).frames(0, N - 1, 27, 5)
    monkeypatch.setenv("SLR_GATHER_MODE", "staged")
    staged = pkg.JointSplat(feat, Z, m).frames(0, N - 1, 27, 5)         # 5 frames: the last CTA of a tile has one frame
    torch.cuda.synchronize()
    assert rel_err(staged.cpu().numpy(), base.cpu().numpy()) <= 1e-5
    want = oracle.joint_splat_baseline(feat.cpu().numpy(), Z.cpu().numpy(), m.cpu().numpy(), (0, 29, N - 1))
    assert rel_err(staged[2:3].cpu().numpy(), want) <= TOL


def test_staged_gather_mode_two_layer_and_ragged(pkg, golden_joint, monkeypatch):
    monkeypatch.setenv("SLR_GATHER_MODE", "staged")
    from slr_sfs_b200 import workloads
    C, H, W, N = 21, 45, 101, 7
    feat, Z, m = workloads.scene(H, W, C, "A", seed=9)
    a_f, a_bg = workloads.two_layer_extras(H, W, seed=9)
    A = torch.sigmoid(a_f) / torch.clamp(torch.sigmoid(a_f) + a_bg, min=1e-8)
    tail = torch.cat([a_f * A.exp(), A.exp()], 1).contiguous()
    js = pkg.JointSplat(feat.cuda(), Z.cuda(), m.cuda(), tail=tail.cuda())
    lo, hi = float(np.float32(1.0 / 600.0)), float(np.float32(599.0 / 600.0))
    gen, aux, mask = js.frames(0, N - 1, 0, N, want_aux=True, want_mask=True, alpha_clamp=(lo, hi))
    for t in (0, 3, N - 1):
        w_gen, w_alpha, w_mask = oracle.joint_splat_2layer(feat.numpy(), Z.numpy(), a_f.numpy(), a_bg.numpy(), m.numpy(),
                                                           (0, t, N - 1), alpha0=True)
        assert rel_err(gen[t:t + 1].cpu().numpy(), w_gen) <= TOL
        alpha_fluid = aux[t:t + 1, 0:1] / torch.clamp(aux[t:t + 1, 1:2], min=1e-8)
        assert rel_err(alpha_fluid.cpu().numpy(), w_alpha) <= TOL
        assert np.mean(mask[t:t + 1].cpu().numpy() != w_mask) < 1e-4


def test_warp_flow_twin_vs_reference_model_golden(pkg, golden_warp_flow):
    """synthesis.warp_flow_block against AnimatingSoftmaxSplating.warp_flow (imported unmodified by
    tests/golden/make_golden.py): animating_softmax_splating.py:1064-1138."""
    from slr_sfs_b200.synthesis import warp_flow_block
    g = golden_warp_flow
    N = int(g["N"])
    for t in (0, 3, N - 2):
        got = warp_flow_block(cu(g["img"]), cu(g["flow_f"][t:t + 1]), cu(g["flow_p"][N - 1 - t:N - t]), (0, t, N - 1))
        assert rel_err(got.cpu().numpy(), g[f"t{t}/PredImg"]) <= TOL, t
