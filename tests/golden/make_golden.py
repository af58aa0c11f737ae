"""Generate tests/golden/*.npz from the REFERENCE itself.  Run once, in the build
container (needs /root/reference); the vectors are committed, this script is the
record of how they were made.

    python tests/golden/make_golden.py

Three sources, all the reference's own code:

1. ``softsplat_ref.npz`` -- the reference's CUDA kernel text (models/softsplat.py:12-326)
   compiled for the CPU by oracle/build.py and run single-threaded: summation splat,
   grad-input, grad-flow, max-warp-norm on seeded inputs incl. edge cases.
2. ``euler_ref.npz`` -- ``euler_integration`` imported unmodified from
   models/projection/euler_integration_manipulator.py and executed on the CPU
   (the hard-coded ``device='cuda'`` of its tensor factories is dropped by a
   context manager; no arithmetic is touched).
3. ``joint_ref.npz`` -- ``AnimatingSoftmaxSplating.forward_flow`` and
   ``AnimatingSoftmaxSplatingJoint.forward_flow`` imported unmodified and called on
   a stand-in ``self`` whose ``softsplater`` runs source (1), whose
   ``projector`` / ``net_alpha_decoder`` record their input (= the joint block's
   output, animating_softmax_splating.py:975 / 2layers...py:1048,1052) and whose
   ``net_alpha_encoder`` returns a fixed tensor.  ``Tensor.cuda()`` is an identity.
"""
import argparse
import contextlib
import os
import sys
import types
import warnings
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"

import oracle  # noqa: E402


# ----------------------------------------------------------------------------
def flow_cases(rng, B, H, W):
    """Flows that exercise the reference's bounds tests, floor and collisions."""
    cases = {}
    cases["zero"] = np.zeros((B, 2, H, W), np.float32)
    cases["int_shift"] = np.tile(np.array([2.0, -3.0], np.float32).reshape(1, 2, 1, 1), (B, 1, H, W))
    cases["half"] = np.full((B, 2, H, W), 0.5, np.float32)
    cases["neg_frac"] = np.tile(np.array([-0.25, -1.75], np.float32).reshape(1, 2, 1, 1), (B, 1, H, W))
    cases["uniform4"] = rng.uniform(-4, 4, (B, 2, H, W)).astype(np.float32)
    cases["uniform_big"] = rng.uniform(-1.5 * W, 1.5 * W, (B, 2, H, W)).astype(np.float32)
    # every pixel lands on the last row / column exactly (only in-bounds corners written)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    edge = np.stack([(W - 1) - xs, (H - 1) - ys], 0)[None].repeat(B, 0)
    cases["onto_last_cell"] = edge.astype(np.float32)
    # many-to-one: everything collapses onto one fractional point
    sink = np.stack([(W / 2 + 0.3) - xs, (H / 3 + 0.6) - ys], 0)[None].repeat(B, 0)
    cases["sink"] = sink.astype(np.float32)
    # Euler sentinel: invalid pixels carry max(H,W)+1 in both channels
    sent = rng.uniform(-2, 2, (B, 2, H, W)).astype(np.float32)
    m = rng.uniform(size=(B, 1, H, W)) < 0.3
    sent = np.where(m, np.float32(max(H, W) + 1), sent).astype(np.float32)
    cases["sentinel"] = sent
    return cases


def make_softsplat(rng):
    out = {}
    B, C, H, W = 2, 3, 11, 14
    for name, flow in flow_cases(rng, B, H, W).items():
        inp = rng.standard_normal((B, C, H, W)).astype(np.float32)
        gout = rng.standard_normal((B, C, H, W)).astype(np.float32)
        out[f"{name}/inp"] = inp
        out[f"{name}/flow"] = flow
        out[f"{name}/gout"] = gout
        out[f"{name}/sum"] = oracle.ref_softsplat_sum(inp, flow)
        out[f"{name}/gin"] = oracle.ref_softsplat_grad_input(inp, flow, gout)
        out[f"{name}/gflow"] = oracle.ref_softsplat_grad_flow(inp, flow, gout)
        z = rng.standard_normal((B, 1, H, W)).astype(np.float32)
        out[f"{name}/z"] = z
        out[f"{name}/maxwarpnorm"] = oracle.ref_max_warp_norm(z, flow)
    # config 1 of BASELINE.json in miniature: softmax mode through FunctionSoftsplat's
    # own python (restated; the kernel is the reference's)
    return out


# ----------------------------------------------------------------------------
@contextlib.contextmanager
def cpu_as_cuda():
    """Let reference code that hard-codes device='cuda' / .cuda() run on the CPU."""
    def strip(fn):
        def wrapped(*a, **kw):
            kw.pop("device", None)
            return fn(*a, **kw)
        return wrapped

    names = ["linspace", "zeros", "ones"]
    saved = {n: getattr(torch, n) for n in names}
    saved_cuda = torch.Tensor.cuda
    try:
        for n in names:
            setattr(torch, n, strip(saved[n]))
        torch.Tensor.cuda = lambda self, *a, **kw: self
        yield
    finally:
        for n in names:
            setattr(torch, n, saved[n])
        torch.Tensor.cuda = saved_cuda


def import_reference():
    warnings.simplefilter("ignore")
    sys.path.insert(0, REFERENCE)
    for m in ["cupy", "cv2", "av", "lz4framed", "lpips", "tensorboardX", "matplotlib", "matplotlib.pyplot"]:
        try:
            __import__(m)
        except Exception:
            sys.modules[m] = mock.MagicMock()
    if isinstance(sys.modules["cupy"], mock.MagicMock):
        sys.modules["cupy"].memoize = lambda **kw: (lambda fn: fn)
    import importlib
    eul = importlib.import_module("models.projection.euler_integration_manipulator")
    base = importlib.import_module("models.animating_softmax_splating")
    two = importlib.import_module("models.animating_softmax_splating_2layers_alpha_seperate")
    return eul, base, two


def motion_fields(rng, H, W):
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    fields = {}
    fields["const"] = np.stack([np.full((H, W), 1.25, np.float32), np.full((H, W), -0.5, np.float32)])[None]
    fields["half_steps"] = np.stack([np.full((H, W), 0.5, np.float32), np.full((H, W), 0.5, np.float32)])[None]
    smooth = np.stack([1.7 * np.sin(xs / 5.0) * np.cos(ys / 7.0) + 0.4,
                       1.3 * np.cos(xs / 6.0 + 1.0) * np.sin(ys / 4.0)])[None]
    fields["smooth"] = smooth.astype(np.float32)
    fields["random"] = rng.uniform(-3, 3, (1, 2, H, W)).astype(np.float32)
    # leaves the frame then points back inside: invalidity must stay sticky
    swirl = np.stack([np.where(xs > W / 2, 2.5, -2.5), np.where(ys > H / 2, -1.5, 1.5)])[None]
    fields["re_enter"] = swirl.astype(np.float32)
    # lands exactly on W-1 / H-1 (valid: the test is strict)
    fields["exact_edge"] = np.stack([np.full((H, W), 1.0, np.float32), np.zeros((H, W), np.float32)])[None]
    return fields


def make_euler(rng, eul):
    out = {}
    for (H, W) in [(9, 13), (16, 16)]:
        for name, m in motion_fields(rng, H, W).items():
            for T in [0, 1, 2, 5, 12]:
                with cpu_as_cuda(), torch.no_grad():
                    d, v = eul.euler_integration(torch.from_numpy(m), T)
                key = f"{name}_{H}x{W}/T{T}"
                out[f"{name}_{H}x{W}/motion"] = m
                out[f"{key}/disp"] = d.numpy().astype(np.float32)
                out[f"{key}/vis"] = v.numpy().astype(np.float32)
    # 1-element tensor as destination_frame (how the models call it)
    m = motion_fields(rng, 9, 13)["smooth"]
    with cpu_as_cuda(), torch.no_grad():
        d, v = eul.euler_integration(torch.from_numpy(m), torch.tensor([4]))
    out["tensorT/motion"] = m
    out["tensorT/T4/disp"] = d.numpy()
    out["tensorT/T4/vis"] = v.numpy()
    return out


def make_euler_grad(rng, eul):
    """torch autograd through the reference's euler_integration (unmodified): the gradient the
    motion regressor receives with --train_motion (animating_softmax_splating.py:514-580)."""
    out = {}
    for (H, W) in [(9, 13), (16, 16)]:
        for name, m in motion_fields(rng, H, W).items():
            for T in [0, 1, 3, 7]:
                g = rng.standard_normal((1, 2, H, W)).astype(np.float32)
                mt = torch.from_numpy(m.copy()).requires_grad_(True)
                with cpu_as_cuda():
                    d, _ = eul.euler_integration(mt, T)
                    if d.requires_grad:          # T = 0 returns constants (:33-34)
                        (d * torch.from_numpy(g)).sum().backward()
                key = f"{name}_{H}x{W}/T{T}"
                out[f"{name}_{H}x{W}/motion"] = m
                out[f"{key}/gdisp"] = g
                out[f"{key}/gmotion"] = (mt.grad if mt.grad is not None else torch.zeros_like(mt)).numpy().astype(np.float32)
    return out


# ----------------------------------------------------------------------------
class RefSplat(torch.nn.Module):
    """ModuleSoftsplat('summation') stand-in backed by the reference kernels on CPU."""
    def forward(self, tenInput, tenFlow, tenMetric):
        o = oracle.ref_softsplat_sum(tenInput.detach().numpy(), tenFlow.detach().numpy())
        return torch.from_numpy(o)


class RefMaxWarpNorm(torch.nn.Module):
    def forward(self, tenInput, tenFlow):
        o = oracle.ref_max_warp_norm(tenInput.detach().numpy(), tenFlow.detach().numpy())
        return torch.from_numpy(o)


class Recorder(torch.nn.Module):
    def __init__(self, out_ch):
        super().__init__()
        self.out_ch = out_ch
        self.seen = None

    def forward(self, x, *rest):
        self.seen = x.detach().clone()
        return torch.zeros(x.shape[0], self.out_ch, x.shape[2], x.shape[3])


def make_joint(rng, base, two):
    out = {}
    W, C, N = 20, 6, 8
    feat = rng.standard_normal((1, C, W, W)).astype(np.float32)
    Z = (2.0 * rng.standard_normal((1, 1, W, W))).astype(np.float32)
    motion = motion_fields(rng, W, W)["smooth"]
    img = rng.standard_normal((1, 3, W, W)).astype(np.float32)
    a_out = rng.standard_normal((1, 2, W, W)).astype(np.float32)  # net_alpha_encoder output
    bg = rng.standard_normal((1, 3, W, W)).astype(np.float32)
    out.update(feat=feat, Z=Z, motion=motion, alpha_encoder_out=a_out, N=np.int64(N))

    for z_mode, flags in [("max", {}), ("v1", {"use_softmax_splatter_v1": True}),
                          ("v2", {"use_softmax_splatter_v2": True})]:
        for t in [0, 3, N - 1]:
            # argparse.Namespace supports the reference's `"flag" in self.opt` idiom; the parser
            # always defines no_clamp_Z (options/train_options.py:583), so the clamp never runs
            opt = argparse.Namespace(W=W, refine_model_type="resnet_256W8UpDown64", no_clamp_Z=False, **flags)
            me = types.SimpleNamespace(opt=opt, softsplater=RefSplat(), projector=Recorder(3),
                                       maximum_warp_norm_splater=RefMaxWarpNorm())
            batch = {"features": [(torch.from_numpy(feat), torch.from_numpy(Z))],
                     "images": [torch.from_numpy(img)],
                     "motions": [torch.from_numpy(motion)],
                     "index": torch.tensor([[0, t, N - 1]])}
            with cpu_as_cuda(), torch.no_grad():
                base.AnimatingSoftmaxSplating.forward_flow(me, batch)
            out[f"baseline/{z_mode}/t{t}/gen_fs"] = me.projector.seen.numpy()

    for alpha0 in [True, False]:
        for t in [0, 3, N - 1]:
            opt = argparse.Namespace(W=W, ngf=C, use_alpha0_as_blending_weight=alpha0)
            me = types.SimpleNamespace(opt=opt, softsplater=RefSplat(), projector=Recorder(3),
                                       net_alpha_decoder=Recorder(1),
                                       net_alpha_encoder=lambda x: torch.from_numpy(a_out))
            batch = {"features": [(torch.from_numpy(feat), torch.from_numpy(Z))],
                     "images": [torch.from_numpy(img)], "BGImg": [torch.from_numpy(bg)],
                     "motions": [torch.from_numpy(motion)],
                     "index": torch.tensor([[0, t, N - 1]])}
            with cpu_as_cuda(), torch.no_grad():
                two.AnimatingSoftmaxSplatingJoint.forward_flow(me, batch)
            dec_in = me.net_alpha_decoder.seen.numpy()      # cat([gen_fs, alpha_fluid]) :1052
            tag = f"twolayer/{'alpha0' if alpha0 else 'plain'}/t{t}"
            out[f"{tag}/gen_fs"] = me.projector.seen.numpy()
            out[f"{tag}/alpha_fluid"] = dec_in[:, -1:]
    return out


def make_warp_flow(rng, base, eul):
    """AnimatingSoftmaxSplating.warp_flow (animating_softmax_splating.py:983-1140) imported unmodified:
    RGB image, Z = 1, precomputed flow lists (here: the reference's own euler_integration per step)."""
    out = {}
    W, N = 20, 8
    img = rng.standard_normal((1, 3, W, W)).astype(np.float32)
    motion = motion_fields(rng, W, W)["smooth"]
    with cpu_as_cuda(), torch.no_grad():
        flow_f = torch.cat([eul.euler_integration(torch.from_numpy(motion), t)[0] for t in range(N + 1)], 0)
        flow_p = torch.cat([eul.euler_integration(torch.from_numpy(-motion), t)[0] for t in range(N + 1)], 0)
    out.update(img=img, flow_f=flow_f.numpy(), flow_p=flow_p.numpy(), N=np.int64(N))
    for t in [0, 3, N - 2]:
        opt = argparse.Namespace(W=W)
        me = types.SimpleNamespace(opt=opt, softsplater=RefSplat())
        batch = {"images": [torch.from_numpy(img)], "motions": [flow_f, flow_p], "index": torch.tensor([[0, t, N - 1]])}
        with cpu_as_cuda(), torch.no_grad():
            pred = base.AnimatingSoftmaxSplating.warp_flow(me, batch)
        out[f"t{t}/PredImg"] = pred["PredImg"].numpy()
    return out


def main():
    assert oracle.ref_available(), "build oracle/_ref first (python oracle/build.py)"
    rng = np.random.default_rng(20261017)
    eul, base, two = import_reference()
    only = set(sys.argv[1:])       # e.g. `make_golden.py euler_grad_ref`: regenerate just that file
    makers = [("softsplat_ref", lambda: make_softsplat(rng)), ("euler_ref", lambda: make_euler(rng, eul)),
              ("joint_ref", lambda: make_joint(rng, base, two)),
              ("euler_grad_ref", lambda: make_euler_grad(np.random.default_rng(20261018), eul)),
              ("warp_flow_ref", lambda: make_warp_flow(np.random.default_rng(20261019), base, eul))]
    for name, make in makers:
        if only and name not in only:
            continue
        data = make()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(path, len(data), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
