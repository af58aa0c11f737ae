"""CPU: the CUDA kernel SOURCES (slr-sfs_b200/csrc/*.cu), compiled for the CPU by tests/emu
and called through the same C ABI, against the oracle and the reference's golden vectors.

This is not a product path (the package only ever loads csrc/libslr_splat.so and has no CPU
mode); it lets the container without a GPU check the logic of the very kernel text the B200
runs -- index arithmetic, list building, heavy-tile fallbacks, barriers -- before GPU time is
spent.  What it cannot see: races (lanes run one after the other), device expf (1 ulp), speed.
Same tolerances as tests/test_gpu_*.py."""
import numpy as np
import pytest

import oracle
from conftest import rel_err

emu = pytest.importorskip("emu", reason="tests/emu")
TOL = 1e-4


def _rng(seed):
    return np.random.default_rng(seed)


def _case(seed, B, C, H, W, amp=3.0):
    r = _rng(seed)
    return (r.standard_normal((B, C, H, W)).astype(np.float32),
            (r.uniform(-amp, amp, (B, 2, H, W))).astype(np.float32))


# ---------------------------------------------------------------------------- operator level
@pytest.mark.parametrize("shape", [(1, 3, 17, 23), (2, 5, 32, 32), (1, 9, 40, 70)])
def test_emu_summation_splat_and_grads(shape):
    B, C, H, W = shape
    inp, flow = _case(1, B, C, H, W)
    out = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, 1, None)
    assert rel_err(out, oracle.softsplat_sum(inp, flow)) <= TOL
    gout = _rng(2).standard_normal(inp.shape).astype(np.float32)
    gin = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_grad_input", emu.p(flow), emu.p(gout), emu.p(gin), B, C, H, W, None)
    assert rel_err(gin, oracle.softsplat_grad_input(flow, gout)) <= TOL
    gflow = np.full_like(flow, np.nan)
    emu.call("slr_softsplat_grad_flow", emu.p(inp), emu.p(flow), emu.p(gout), emu.p(gflow), B, C, H, W, None)
    assert rel_err(gflow, oracle.softsplat_grad_flow(inp, flow, gout)) <= TOL


@pytest.mark.parametrize("shape", [(1, 3, 17, 23), (2, 9, 40, 70), (1, 65, 24, 40)])
def test_emu_summation_splat_through_the_gather(shape):
    """slr_softsplat_sum_fwd_gather: the operator-level splat as scene prep + one-frame flow table + insert + gather of
    the un-normalised sums, against the oracle and against the atomic scatter; static rows, far flows, ragged tiles."""
    B, C, H, W = shape
    inp, flow = _case(4, B, C, H, W, amp=5.0)
    flow[:, :, : H // 4] = 0.0                       # pixels that do not move receive themselves
    flow[:, :, -2:, :5] = 1000.0                     # ... and some that leave the frame
    nb = emu.lib().slr_softsplat_gather_scratch_bytes(C, H, W)
    assert nb > 0
    scratch = emu.aligned(nb)
    scratch[:] = 0x5A
    out = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_sum_fwd_gather", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, emu.p(scratch), nb, None)
    want = oracle.softsplat_sum(inp, flow)
    assert rel_err(out, want) <= TOL
    assert np.all(out[want == 0.0] == 0.0)
    ref = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(ref), B, C, H, W, 1, None)
    assert rel_err(out, ref) <= 1e-5


def test_emu_summation_splat_through_the_gather_convergent_flow():
    """Everything onto a few pixels: lists cut at the list depth, excess pairs by reductions, and (one frame, small
    excess list) the whole-batch scatter fallback -- all without the normalisation."""
    B, C, H, W = 1, 5, 40, 72
    inp, _ = _case(6, B, C, H, W)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    for flow in (np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys])[None].astype(np.float32),
                 np.stack([-(xs - W / 2) * 0.45, -(ys - H / 2) * 0.2])[None].astype(np.float32)):
        nb = emu.lib().slr_softsplat_gather_scratch_bytes(C, H, W)
        scratch = emu.aligned(nb)
        out = np.full_like(inp, np.nan)
        emu.call("slr_softsplat_sum_fwd_gather", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, emu.p(scratch), nb, None)
        assert rel_err(out, oracle.softsplat_sum(inp, flow)) <= TOL


def test_emu_golden_cases_from_the_reference_kernels(golden_softsplat):
    g = golden_softsplat
    for case in sorted({k.split("/")[0] for k in g.files if "/" in k}):
        inp, flow, gout = g[case + "/inp"], g[case + "/flow"], g[case + "/gout"]
        B, C, H, W = inp.shape
        out, gin, gflow = np.empty_like(inp), np.empty_like(inp), np.empty_like(flow)
        emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, 1, None)
        emu.call("slr_softsplat_grad_input", emu.p(flow), emu.p(gout), emu.p(gin), B, C, H, W, None)
        emu.call("slr_softsplat_grad_flow", emu.p(inp), emu.p(flow), emu.p(gout), emu.p(gflow), B, C, H, W, None)
        assert rel_err(out, g[case + "/sum"]) <= TOL, case
        assert rel_err(gin, g[case + "/gin"]) <= TOL, case
        assert rel_err(gflow, g[case + "/gflow"]) <= TOL, case
        z = g[case + "/z"]
        scratch, mw = np.empty_like(z), np.empty_like(z)
        emu.call("slr_maxwarpnorm", emu.p(z), emu.p(flow), emu.p(scratch), emu.p(mw), *z.shape, None)
        assert np.array_equal(mw, g[case + "/maxwarpnorm"]), case


def test_emu_max_warp_norm_bit_exact():
    inp, flow = _case(3, 2, 1, 24, 40, amp=5.0)
    scratch, out = np.empty_like(inp), np.empty_like(inp)
    emu.call("slr_maxwarpnorm", emu.p(inp), emu.p(flow), emu.p(scratch), emu.p(out), 2, 1, 24, 40, None)
    assert np.array_equal(out, oracle.max_warp_norm(inp, flow))


@pytest.mark.parametrize("T", [0, 1, 7, 60])
def test_emu_euler_bit_exact(T):
    H, W = 33, 50
    motion = (_rng(4).uniform(-2.5, 2.5, (1, 2, H, W))).astype(np.float32)
    motion[:, :, 10:20, 5:30] = 0.0
    for sign in (1.0, -1.0):
        disp = np.empty((1, 2, H, W), dtype=np.float32)
        vis = np.empty((1, 1, H, W), dtype=np.float32)
        emu.call("slr_euler", emu.p(motion), sign, T, emu.p(disp), emu.p(vis), H, W, None)
        want_d, want_v = oracle.euler(np.float32(sign) * motion, T)
        assert np.array_equal(disp, want_d) and np.array_equal(vis, want_v)


def test_emu_reduce_max_mixed_signs():
    x = _rng(5).standard_normal(5000).astype(np.float32) - 3.0
    out = np.zeros(1, dtype=np.float32)
    emu.call("slr_reduce_max", emu.p(x), x.size, emu.p(out), None)
    assert out[0] == x.max()
    x = -np.abs(x) - 1.0
    emu.call("slr_reduce_max", emu.p(x), x.size, emu.p(out), None)
    assert out[0] == x.max()


# ---------------------------------------------------------------------------- joint block
def _scene(H, W, C, kind, seed):
    from slr_sfs_b200 import workloads
    feat, Z, m = workloads.scene(H, W, C, kind, seed=seed)
    return feat.numpy(), Z.numpy(), m.numpy()


def test_emu_scatter_joint_vs_oracle():
    H, W, C, N, t = 24, 40, 6, 9, 4
    feat, Z, motion = _scene(H, W, C, "A", 1)
    zsub = np.array([Z.max()], dtype=np.float32)
    disp = np.empty((2, 2, H, W), dtype=np.float32)
    emu.call("slr_euler", emu.p(motion), 1.0, t, emu.p(disp[0]), None, H, W, None)
    emu.call("slr_euler", emu.p(motion), -1.0, N - 1 - t + 1, emu.p(disp[1]), None, H, W, None)
    acc = np.full((1, C + 1, H, W), np.nan, dtype=np.float32)
    alpha = float(oracle.blend_alpha(t, 0, N - 1))
    emu.call("slr_joint_scatter", emu.p(feat), emu.p(Z), emu.p(zsub), None, 0, emu.p(disp[0]), emu.p(disp[1]),
             alpha, emu.p(acc), C, H, W, None)
    out = np.empty((1, C, H, W), dtype=np.float32)
    emu.call("slr_normalize", emu.p(acc), emu.p(out), None, C, C, C + 1, 1e-8, H, W, None)
    assert rel_err(out, oracle.joint_splat_baseline(feat, Z, motion, (0, t, N - 1))) <= TOL


@pytest.mark.parametrize("kind,shape", [("A", (40, 64)), ("B", (24, 40)), ("C", (17, 37)), ("A", (9, 33))])
def test_emu_gather_clip_vs_oracle(kind, shape):
    H, W = shape
    C, N = 6, 8
    feat, Z, motion = _scene(H, W, C, kind, 2)
    sc = emu.Scene(feat, Z, motion)
    got = sc.frames(0, N - 1, 0, N)
    assert not np.isnan(got).any()
    for t in range(N):
        want = oracle.joint_splat_baseline(feat, Z, motion, (0, t, N - 1))
        assert rel_err(got[t:t + 1], want) <= TOL, (kind, t)
        assert np.all(got[t:t + 1][want == 0.0] == 0.0)


@pytest.mark.parametrize("shape", ["1x4", "2x2"])
def test_emu_gather_cta_shapes_agree(shape, monkeypatch):
    """rowgather_kernel's CTA shapes (frames x row pairs per CTA, SLR_GATHER_SHAPE) only regroup
    the same warps: identical results, including a ragged frame count and ragged image edges."""
    H, W, C, N = 21, 45, 5, 7
    feat, Z, motion = _scene(H, W, C, "A", 6)
    sc = emu.Scene(feat, Z, motion)
    monkeypatch.delenv("SLR_GATHER_SHAPE", raising=False)
    base = sc.frames(0, N - 1, 0, N)
    monkeypatch.setenv("SLR_GATHER_SHAPE", shape)
    got = sc.frames(0, N - 1, 0, N)
    assert np.array_equal(got, base)
    want = oracle.joint_splat_baseline(feat, Z, motion, (0, 3, N - 1))
    assert rel_err(got[3:4], want) <= TOL


def test_emu_gather_nonzero_start_split_calls_and_v1():
    H, W, C = 24, 40, 5
    feat, Z, motion = _scene(H, W, C, "A", 3)
    sc = emu.Scene(feat, Z, motion, z_mode="v1")
    got = sc.frames(2, 11, 5, 3, split=True)
    for i, t in enumerate((5, 6, 7)):
        want = oracle.joint_splat_baseline(feat, Z, motion, (2, t, 11), z_mode="v1")
        assert rel_err(got[i:i + 1], want) <= TOL, t


def test_emu_clip_table_batches_equal_per_batch_plans():
    """slr_clip_table once for the clip + slr_clip_bin per batch == slr_clip_plan per batch
    (same chains, continued instead of restarted)."""
    H, W, C = 24, 72, 4
    feat, Z, motion = _scene(H, W, C, "A", 7)
    sc = emu.Scene(feat, Z, motion)
    start, end = 1, 14
    tab = sc.table(start, end, 3, 11)                       # frames 3..13 of the clip [1, 14]
    for t0, n in ((3, 4), (7, 5), (12, 2), (5, 1)):
        a = sc.frames(start, end, t0, n, table=tab)
        b = sc.frames(start, end, t0, n, split=True)
        assert np.array_equal(a, b), (t0, n)
    want = oracle.joint_splat_baseline(feat, Z, motion, (start, 9, end))
    assert rel_err(sc.frames(start, end, 9, 1, table=tab), want) <= TOL
    # the last frame index the reference loop ever asks for (t = end) and a whole-clip table
    tab = sc.table(start, end, start, end - start + 1)
    want = oracle.joint_splat_baseline(feat, Z, motion, (start, end, end))
    assert rel_err(sc.frames(start, end, end, 1, table=tab), want) <= TOL


def test_emu_two_layer_aux_and_mask():
    from slr_sfs_b200 import workloads
    H, W, C, N = 24, 40, 4, 6
    feat, Z, motion = _scene(H, W, C, "A", 4)
    a_f, a_bg = [t.numpy() for t in workloads.two_layer_extras(H, W, seed=4)]
    s = 1.0 / (1.0 + np.exp(-a_f))
    A = (s / np.maximum(s + a_bg, 1e-8)).astype(np.float32)
    tail = np.concatenate([a_f * np.exp(A), np.exp(A)], 1).astype(np.float32)
    sc = emu.Scene(feat, Z, motion, tail=tail)
    lo, hi = float(np.float32(1.0 / 600.0)), float(np.float32(599.0 / 600.0))
    gen, aux, mask = sc.frames(0, N - 1, 0, N, alpha_clamp=(lo, hi), want_aux=True, want_mask=True)
    for t in (0, 2, N - 1):
        w_gen, w_alpha, w_mask = oracle.joint_splat_2layer(feat, Z, a_f, a_bg, motion, (0, t, N - 1), alpha0=True)
        assert rel_err(gen[t:t + 1], w_gen) <= TOL
        alpha_fluid = aux[t:t + 1, 0:1] / np.maximum(aux[t:t + 1, 1:2], np.float32(1e-8))
        assert rel_err(alpha_fluid, w_alpha) <= TOL
        assert np.array_equal(mask[t:t + 1], w_mask)


def _sink_scene(H, W, C, seed):
    feat, Z, _ = _scene(H, W, C, "A", seed)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    sink = np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys])[None].astype(np.float32)     # everything -> one point in 1 step
    squeeze = np.stack([-(xs - W / 2) * 0.45, -(ys - H / 2) * 0.2])[None].astype(np.float32)  # strong compression
    return feat, Z, sink, squeeze


def test_emu_convergent_flows_take_the_heavy_paths():
    """Flows that pile many sources onto few destinations: deep lists (tail loop of the gather),
    lists cut at the list depth with the excess pairs added by reductions (flag 1)."""
    H, W, C, N = 40, 72, 6, 3
    feat, Z, sink, squeeze = _sink_scene(H, W, C, 9)
    seen_flagged = 0
    for m in (sink, squeeze):
        sc = emu.Scene(feat, Z, m)
        got = sc.frames(0, N - 1, 1, 2)
        seen_flagged += sc.stats["flagged"]
        for i, t in enumerate((1, 2)):
            want = oracle.joint_splat_baseline(feat, Z, m, (0, t, N - 1))
            assert rel_err(got[i:i + 1], want) <= TOL
            assert np.all(got[i:i + 1][want == 0.0] == 0.0)
    assert seen_flagged > 0


@pytest.mark.parametrize("mode", ["ldg", "bins"])
def test_emu_excess_list_overflow_falls_back_to_reductions(mode, monkeypatch):
    """A one-step sink puts 4*P pairs onto four pixels; with a single frame in the batch the excess
    list (2*P*n entries) overflows.  Bin pipeline: the tile goes the flag-2 way (zeroed, every pair of its
    bin added by reductions, divided at the end).  Direct index: pairs were dropped, so the whole batch is
    redone by scatter + divide (overflow_*_kernel); also with the optional planes and a 2-layer tail."""
    monkeypatch.setenv("SLR_GATHER_MODE", mode)
    H, W, C, N = 40, 72, 6, 3
    feat, Z, sink, _ = _sink_scene(H, W, C, 9)
    sc = emu.Scene(feat, Z, sink)
    got, aux, mask, nnz = sc.frames(0, N - 1, 1, 1, want_aux=True, want_mask=True, want_nnz=True)
    assert sc.stats["excess"] > sc.stats["excess_cap"]
    assert (sc.stats["full"] > 0) == (mode == "bins")
    want = oracle.joint_splat_baseline(feat, Z, sink, (0, 1, N - 1))
    assert rel_err(got, want) <= TOL
    assert np.all(got[want == 0.0] == 0.0)
    assert np.array_equal(nnz[0, 0], (got[0] != 0).sum(0).astype(np.float32))
    assert np.array_equal(mask[0, 0] != 0, aux[0, -1] > 1e-8)


@pytest.mark.parametrize("kind", ["A", "B", "sink", "squeeze"])
def test_emu_direct_index_equals_bin_pipeline(kind, monkeypatch):
    """insert_kernel (every source writes its cells straight into the destination lanes' lists) against
    bin_fill + expand_kernel: the same (source, weight) pairs in the same canonical slots -- only which of two
    competing sources gets the canonical slot and which the overflow slot may differ, i.e. the summation order."""
    from slr_sfs_b200 import workloads
    H, W, C, N = 40, 72, 6, 7
    feat, Z, motion = _scene(H, W, C, "A" if kind in ("sink", "squeeze") else kind, 31)
    if kind in ("sink", "squeeze"):
        _, _, sink, squeeze = _sink_scene(H, W, C, 31)
        motion = sink if kind == "sink" else squeeze
    tail = np.abs(feat[:, :2]) + 0.5
    res = {}
    for mode in ("ldg", "bins"):
        monkeypatch.setenv("SLR_GATHER_MODE", mode)
        sc = emu.Scene(feat, Z, motion, tail=tail)
        res[mode] = sc.frames(0, N - 1, 1, 5, want_aux=True, want_mask=True, want_nnz=True)
        res[mode + "_stats"] = sc.stats
    for a, b in zip(res["ldg"], res["bins"]):
        assert rel_err(a, b) <= 1e-5
    # an overflow cell of the direct index carries both rows' weights of its source (expand_kernel spills one entry
    # per pair): its lists are never deeper
    assert res["ldg_stats"]["flagged"] <= res["bins_stats"]["flagged"]
    for i in (0, 4):
        want = oracle.joint_splat_baseline(feat, Z, motion, (0, 1 + i, N - 1))
        assert rel_err(res["ldg"][0][i:i + 1], want) <= TOL
        assert np.all(res["ldg"][0][i:i + 1][want == 0.0] == 0.0)


def test_emu_static_pixels_and_negative_zero():
    H, W, C, N = 24, 40, 4, 8
    r = _rng(21)
    feat = r.standard_normal((1, C, H, W)).astype(np.float32)
    Z = r.standard_normal((1, 1, H, W)).astype(np.float32)
    zero = np.zeros((1, 2, H, W), np.float32)
    out = emu.Scene(feat, Z, zero).frames(0, N - 1, 0, N)
    for t in (0, 5, N - 1):
        assert rel_err(out[t:t + 1], feat) <= 1e-6
    m = zero.copy()
    m[0, 0, :, : W // 2] = 1.7
    m[0, 1, : H // 2, : W // 2] = -0.6
    m[0, 0, 3, 5] = -0.0
    got = emu.Scene(feat, Z, m).frames(0, N - 1, 0, N)
    for t in (1, 6, N - 1):
        assert rel_err(got[t:t + 1], oracle.joint_splat_baseline(feat, Z, m, (0, t, N - 1))) <= TOL


def test_emu_clip_table_bin_stats_case_shared_with_the_gpu_suite():
    import clip_abi_cases
    clip_abi_cases.clip_table_bin_stats(clip_abi_cases.EmuBackend())


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_emu_euler_grad_motion_vs_oracle(sign):
    H, W, T = 19, 27, 9
    r = _rng(31)
    motion = r.uniform(-2.0, 2.0, (1, 2, H, W)).astype(np.float32)
    g = r.standard_normal((1, 2, H, W)).astype(np.float32)
    got = np.full_like(motion, np.nan)
    emu.call("slr_euler_grad_motion", emu.p(motion), sign, T, emu.p(g), emu.p(got), H, W, None)
    want = np.float32(sign) * oracle.euler_grad_motion(np.float32(sign) * motion, T, g)
    assert rel_err(got, want) <= 1e-6
    emu.call("slr_euler_grad_motion", emu.p(motion), sign, 0, emu.p(g), emu.p(got), H, W, None)
    assert not got.any()


# ---------------------------------------------------------------------------
# stagegather_kernel (sources staged in shared memory by bulk copies) vs rowgather_kernel (through L1), both on the
# lists expand_kernel builds from the bins (SLR_GATHER_MODE=bins: the direct index fills the same slots in another order)
# ---------------------------------------------------------------------------
def _both_modes(monkeypatch, make):
    monkeypatch.setenv("SLR_GATHER_MODE", "bins")
    sc = make()
    ldg = sc.frames(0, sc.N - 1, 0, sc.N, want_aux=True, want_mask=True)
    assert sc.stats["fallback"] == 0
    monkeypatch.setenv("SLR_GATHER_MODE", "staged")
    sc = make()
    staged = sc.frames(0, sc.N - 1, 0, sc.N, want_aux=True, want_mask=True)
    return ldg, staged, sc.stats


@pytest.mark.parametrize("C", [3, 16, 21, 40])
def test_emu_staged_gather_matches_l1_gather_smooth_flow(monkeypatch, C):
    """Regular flow: every tile is staged (no fallback), and the staged gather adds the same
    products in the same order as the L1 gather -- bit-identical outputs, for channel counts
    that are not multiples of the 16-channel chunk too."""
    H, W, N = 40, 100, 5
    feat, Z, motion = _scene(H, W, C, "A", 17)
    tail = np.abs(feat[:, :2]) + 0.5

    def make():
        sc = emu.Scene(feat, Z, motion, tail=tail)
        sc.N = N
        return sc
    ldg, staged, stats = _both_modes(monkeypatch, make)
    assert stats["fallback"] == 0
    for a, b in zip(ldg, staged):
        assert np.array_equal(a, b)
    want = oracle.joint_splat_baseline(feat, Z, motion, (0, 3, N - 1))
    assert rel_err(staged[0][3:4], want) <= TOL


def test_emu_staged_gather_odd_frame_count_and_ragged_edges(monkeypatch):
    H, W, C, N = 37, 67, 5, 3          # 3 frames: the second CTA of a tile has one frame only; ragged tiles
    feat, Z, motion = _scene(H, W, C, "A", 18)

    def make():
        sc = emu.Scene(feat, Z, motion)
        sc.N = N
        return sc
    ldg, staged, stats = _both_modes(monkeypatch, make)
    for a, b in zip(ldg, staged):
        assert np.array_equal(a, b)
    for t in range(N):
        assert rel_err(staged[0][t:t + 1], oracle.joint_splat_baseline(feat, Z, motion, (0, t, N - 1))) <= TOL


def test_emu_staged_gather_incoherent_flow_falls_back(monkeypatch):
    """i.i.d. flow of +-8 px over several steps scatters a tile's sources over more rows / pixels than
    the staging area holds: those tiles are flagged and gathered through L1; results unchanged."""
    H, W, C, N = 192, 96, 4, 14
    feat, Z, motion = _scene(H, W, C, "B", 19)

    def make():
        sc = emu.Scene(feat, Z, motion)
        sc.N = N
        return sc
    monkeypatch.setenv("SLR_GATHER_MODE", "bins")
    sc = make()
    ldg = sc.frames(0, N - 1, 5, 4)
    monkeypatch.setenv("SLR_GATHER_MODE", "staged")
    sc = make()
    staged = sc.frames(0, N - 1, 5, 4)
    assert sc.stats["fallback"] > 0
    assert np.array_equal(ldg, staged)


def test_emu_staged_gather_deep_lists_and_heavy_tiles(monkeypatch):
    """Convergent flows: list slots beyond the 16 in registers are converted in place and read per
    channel group; flagged tiles leave raw sums to the heavy kernels -- through the staged path too."""
    H, W, C, N = 40, 72, 6, 3
    feat, Z, sink, squeeze = _sink_scene(H, W, C, 9)
    for m in (sink, squeeze):
        monkeypatch.setenv("SLR_GATHER_MODE", "bins")
        ldg = emu.Scene(feat, Z, m).frames(0, N - 1, 1, 2)
        monkeypatch.setenv("SLR_GATHER_MODE", "staged")
        sc = emu.Scene(feat, Z, m)
        staged = sc.frames(0, N - 1, 1, 2)
        assert sc.stats["fallback"] < sc.stats["tiles"]
        assert rel_err(staged, ldg) <= 1e-6           # heavy tiles add their excess pairs with atomics: same on the emulator
        for i, t in enumerate((1, 2)):
            assert rel_err(staged[i:i + 1], oracle.joint_splat_baseline(feat, Z, m, (0, t, N - 1))) <= TOL


def test_emu_staged_gather_single_stage_and_capacity_fallback(monkeypatch):
    """With a small staging area (compile-time variant) the same scene exercises all three cases:
    tiles that fit twice (double buffered), tiles that fit once (single stage: the copies of the next
    chunk wait for the gather of the current one) and tiles that do not fit (L1 fallback)."""
    H, W, C, N = 40, 100, 37, 6
    feat, Z, motion = _scene(H, W, C, "A", 23)
    monkeypatch.setenv("SLR_GATHER_MODE", "bins")
    ldg = emu.Scene(feat, Z, motion).frames(0, N - 1, 0, N)
    monkeypatch.setenv("SLR_GATHER_MODE", "staged")
    seen = []
    for nbytes in (64 * 1024, 40 * 1024, 24 * 1024):
        with emu.variant("stage%d" % nbytes, ["-DSLR_STAGE_BYTES=%d" % nbytes]):
            sc = emu.Scene(feat, Z, motion)
            got = sc.frames(0, N - 1, 0, N)
            seen.append(sc.stats["fallback"])
        assert np.array_equal(got, ldg), nbytes
    assert seen[0] <= seen[1] <= seen[2] and seen[2] > 0 and seen[0] < sc.stats["tiles"]


def test_emu_warp_flow_twin_vs_reference_model_golden(golden_warp_flow):
    """slr_joint_scatter_weights + slr_normalize as synthesis.warp_flow_block drives them, against the
    reference's warp_flow (RGB, Z = 1, separate direction weights, alpha without the + 1)."""
    g = golden_warp_flow
    N = int(g["N"])
    img = emu.f32(g["img"])
    _, C, H, W = img.shape
    ones = np.ones((1, 1, H, W), np.float32)
    zmax = np.ones(1, np.float32)
    for t in (0, 3, N - 2):
        fwd, bwd = emu.f32(g["flow_f"][t]), emu.f32(g["flow_p"][N - 1 - t])
        alpha = np.float32(1.0) - np.float32(t) / np.float32(N - 1)
        acc = np.full((1, C + 1, H, W), np.nan, np.float32)
        out = np.full((1, C, H, W), np.nan, np.float32)
        emu.call("slr_joint_scatter_weights", emu.p(img), emu.p(ones), emu.p(zmax), None, 0, emu.p(fwd), emu.p(bwd),
                 float(alpha), float(np.exp(np.float32(1.0)) * (np.float32(1.0) - alpha)), emu.p(acc), C, H, W, None)
        emu.call("slr_normalize", emu.p(acc), emu.p(out), None, C, C, C + 1, 1e-8, H, W, None)
        assert rel_err(out, g[f"t{t}/PredImg"]) <= TOL, t
