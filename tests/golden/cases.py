"""Seeded inputs of the golden cases (shared by make_golden.py, the oracle tests and the GPU tests)."""
import numpy as np
import torch

from comfyui_propainter_nodes_b200.synthetic import synthetic_clip, synthetic_mask

RAFT_ITERS = 4


def _clip(T, H, W, seed):
    img = synthetic_clip(T, H, W, seed)  # [T,H,W,3] 0..1
    return (img.permute(0, 3, 1, 2) * 2 - 1)[None].contiguous()  # [1,T,3,H,W] in [-1,1]


def _mask(T, H, W):
    return synthetic_mask(T, H, W)[None, :, None].contiguous()  # [1,T,1,H,W]


def _flows(T, H, W, seed, amp=3.0):
    g = torch.Generator().manual_seed(seed)
    lo = torch.randn(2, T, 2, H // 8 + 1, W // 8 + 1, generator=g)
    up = torch.nn.functional.interpolate(lo.view(-1, 2, H // 8 + 1, W // 8 + 1), size=(H, W), mode="bicubic",
                                         align_corners=True).view(2, 1, T, 2, H, W)
    base = torch.tensor([1.5, 0.75]).view(1, 1, 1, 2, 1, 1)
    f = base + amp * 0.3 * up
    return f[0].contiguous(), (-f[1]).contiguous()


def raft_case():
    # H/8 >= 16 is required: the 4th pyramid level must be >= 2x2 or the reference's own
    # normalisation 2*y/(H-1) divides by zero (RAFT/utils/utils.py:69-70)
    return _clip(3, 128, 160, 11)


def rfc_case():
    T, H, W = 5, 64, 96
    ff, fb = _flows(T - 1, H, W, 21)
    return (ff, fb), _mask(T, H, W)


def imgprop_case():
    T, H, W = 6, 48, 64
    ff, fb = _flows(T - 1, H, W, 31, amp=2.0)
    return _clip(T, H, W, 32), _mask(T, H, W), (ff, fb)


def window_case():
    # 5 local + 2 reference frames at 96x144 -> tokens 8x12, windows 2x2 (padded 10x18), pooled 2x4
    l_t, n_ref, H, W = 5, 2, 96, 144
    t = l_t + n_ref
    ff, fb = _flows(l_t - 1, H, W, 41, amp=2.0)
    m = _mask(t, H, W)
    g = torch.Generator().manual_seed(42)
    mu = (m * (torch.rand(1, t, 1, H, W, generator=g) > 0.5)).float()
    fr = _clip(t, H, W, 43) * (1 - mu)
    return dict(frames=fr, flows=(ff, fb), masks_in=m, masks_upd=mu, l_t=l_t)


def e2e_case():
    T, H, W = 8, 128, 160
    return dict(T=T, H=H, W=W, image=synthetic_clip(T, H, W, 51), mask=synthetic_mask(T, H, W),
                ref_stride=3, neighbor_length=4, subvideo_length=80, raft_iter=3)


# ---- round 2: cases on BASELINE.json's configs and on the branches the small cases above never take -----------

NODE_DEFAULTS = dict(mask_dilates=5, flow_mask_dilates=8, ref_stride=10, neighbor_length=10, subvideo_length=80)


def c1_case():
    """BASELINE config[0]: 16 frames 320x180 -> processed at 320x176 (PIL bicubic resize), raft_iter=5, fp32 reference,
    through the Inpaint node."""
    T, H, W = 16, 180, 320
    kw = dict(NODE_DEFAULTS, width=320, height=180, raft_iter=5, fp16="disable")
    return dict(image=synthetic_clip(T, H, W, 61), mask=synthetic_mask(T, H, W), kwargs=kw)


RAFT20_ITERS = (1, 5, 10, 15, 20)
RAFT20_GAINS = {"damped": 0.15, "undamped": 1.3}


def raft20_case():
    """3 frames (2 pairs) at BASELINE's 640x360 for the raft_iter=20 error-growth test."""
    return _clip(3, 360, 640, 71)


def chunked_case():
    """T > subvideo_length: drives the halo branches of complete_flow / image_propagation and the ref_num schedule
    (reference propainter_inference.py:115-139, 172-209, 49-57)."""
    T, H, W = 26, 128, 128
    return dict(T=T, H=H, W=W, image=synthetic_clip(T, H, W, 81), mask=synthetic_mask(T, H, W),
                ref_stride=3, neighbor_length=4, subvideo_length=12, raft_iter=2)


def outpaint_case():
    """Outpaint node: 8 frames 160x128, width_scale 1.2 -> canvas 192x128; token grid 11x16 is padded to 15x18 (both
    axes), so this is also the padded-grid case."""
    T, H, W = 8, 128, 160
    kw = dict(width=160, height=128, width_scale=1.2, height_scale=1.0, mask_dilates=5, flow_mask_dilates=8,
              ref_stride=3, neighbor_length=4, subvideo_length=80, raft_iter=3, fp16="disable")
    return dict(image=synthetic_clip(T, H, W, 91), kwargs=kw)
