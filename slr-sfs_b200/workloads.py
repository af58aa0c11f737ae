"""Synthetic scenes for benchmarks and parity tests (SURVEY.md section 8d / BASELINE.md section 4).

Seeded, CPU-generated fp32 tensors shaped like what the reference's encoder hands
to the joint block: features [1,C,H,W], importance Z [1,1,H,W], Eulerian motion
[1,2,H,W] in pixels per frame.
"""
import math

import torch


def motion_field(kind, H, W, seed=0):
    """'A': smooth low-frequency field, |m| <= 3 px/frame, about half of the pixels static
    (mirrors the fluid / static-background split of real scenes);
    'B': i.i.d. U(-8, 8) per pixel (worst-case incoherent scatter);
    'C': constant (1.25, -0.5) (known answer)."""
    g = torch.Generator().manual_seed(1000 + seed)
    ys = torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W)
    xs = torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W)
    if kind == "C":
        return torch.stack([torch.full((H, W), 1.25), torch.full((H, W), -0.5)])[None].contiguous()
    if kind == "B":
        return (torch.rand(1, 2, H, W, generator=g) * 16.0 - 8.0).contiguous()
    assert kind == "A"
    mx = torch.zeros(H, W)
    my = torch.zeros(H, W)
    for _ in range(4):
        fx, fy = (torch.rand(2, generator=g) * 2.5 + 0.5).tolist()
        px, py = (torch.rand(2, generator=g) * 2 * math.pi).tolist()
        ax, ay = (torch.rand(2, generator=g) * 2 - 1).tolist()
        wave = torch.sin(2 * math.pi * fx * xs / W + px) * torch.cos(2 * math.pi * fy * ys / H + py)
        mx += ax * wave
        my += ay * torch.cos(2 * math.pi * fx * xs / W + py) * torch.sin(2 * math.pi * fy * ys / H + px)
    mag = torch.sqrt(mx * mx + my * my).max().clamp(min=1e-6)
    mx, my = mx * (3.0 / mag), my * (3.0 / mag)
    # smooth blob mask: roughly half of the frame is static (motion exactly 0)
    blob = torch.sin(2 * math.pi * xs / W * 1.5 + 0.7) * torch.sin(2 * math.pi * ys / H * 1.0 + 0.3)
    soft = torch.clamp(blob * 4.0, 0.0, 1.0)
    return torch.stack([mx * soft, my * soft])[None].contiguous()


def scene(H, W, C=64, motion="A", seed=0):
    """features, Z, motion for one synthetic scene (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(1, C, H, W, generator=g)
    Z = torch.randn(1, 1, H, W, generator=g)
    return feat, Z, motion_field(motion, H, W, seed)


def two_layer_extras(H, W, seed=0):
    """Raw fluid-alpha logits a_f and background alpha (after sigmoid) a_bg of the
    2-layer model's alpha encoder (2layers...py:938-948)."""
    g = torch.Generator().manual_seed(5000 + seed)
    return torch.randn(1, 1, H, W, generator=g), torch.rand(1, 1, H, W, generator=g)
