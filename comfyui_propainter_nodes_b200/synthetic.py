"""Seeded synthetic clips for parity tests and the bench (no datasets exist offline).

A low-frequency random texture translated by (+1.5, +0.75) px/frame plus 5 % per-pixel noise, and a
static rectangular mask covering rows H/3..2H/3, cols W/3..2W/3.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def synthetic_clip(T: int, H: int, W: int, seed: int = 1234) -> torch.Tensor:
    """IMAGE tensor [T,H,W,3] float32 in [0,1] (ComfyUI layout)."""
    g = torch.Generator().manual_seed(seed)
    pad = max(160, int(1.5 * T) + 8)  # room for the translation (160 keeps the <=100-frame clips unchanged)
    base = torch.rand(1, 3, (H + pad) // 8 + 2, (W + pad) // 8 + 2, generator=g)
    big = F.interpolate(base, scale_factor=8, mode="bicubic", align_corners=False).clamp(0, 1)[0]
    frames = []
    for t in range(T):
        dx, dy = 1.5 * t, 0.75 * t
        x0, y0 = int(np.floor(dx)), int(np.floor(dy))
        fx, fy = dx - x0, dy - y0
        # bilinear sub-pixel crop
        c = lambda yy, xx: big[:, yy:yy + H, xx:xx + W]
        fr = ((1 - fy) * (1 - fx) * c(y0, x0) + (1 - fy) * fx * c(y0, x0 + 1)
              + fy * (1 - fx) * c(y0 + 1, x0) + fy * fx * c(y0 + 1, x0 + 1))
        fr = fr + 0.05 * (torch.rand(3, H, W, generator=g) - 0.5)
        frames.append(fr.clamp(0, 1).permute(1, 2, 0))
    return torch.stack(frames).float()


def synthetic_mask(T: int, H: int, W: int) -> torch.Tensor:
    """MASK tensor [T,H,W] float32, 1.0 inside the hole."""
    m = torch.zeros(T, H, W)
    m[:, H // 3: 2 * H // 3, W // 3: 2 * W // 3] = 1.0
    return m
