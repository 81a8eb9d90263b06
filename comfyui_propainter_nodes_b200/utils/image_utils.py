"""Host-side pre/post-processing with the reference's exact quantisation semantics.

Mirrors the public names of the reference's utils/image_utils.py (ImageConfig, ImageOutpaintConfig,
convert_image_to_frames, read_masks, prepare_frames_and_masks, extrapolation,
prepare_frames_and_masks_for_outpaint, handle_output) so callers can switch packages unchanged.
These functions define the tensors the CUDA path receives, so they follow the reference's integer
semantics step by step (reference: utils/image_utils.py):

* IMAGE float -> uint8 by ``*255``, clip, **truncate** (:106-116)
* resize to (w - w%8, h - h%8) with PIL's default (bicubic) filter, frames and masks alike (:22-27, :98-103)
* masks: 8-bit, resized, then "any non-zero" grown by N iterations of a cross-shaped binary dilation
  (scipy.ndimage.binary_dilation default structure), or thresholded at 0.1 when N == 0 (:142-175)
* tensors [1,T,C,H,W]; frames scaled to [-1,1], masks in {0,1} (:178-197)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.ndimage
import torch
from PIL import Image


def _floor8(v: int) -> int:
    return v - v % 8


@dataclass
class ImageConfig:
    width: int
    height: int
    mask_dilates: int
    flow_mask_dilates: int
    input_size: tuple
    video_length: int
    process_size: tuple = field(init=False)

    def __post_init__(self) -> None:
        self.process_size = (_floor8(self.width), _floor8(self.height))


@dataclass
class ImageOutpaintConfig(ImageConfig):
    width_scale: float = 1.0
    height_scale: float = 1.0
    outpaint_size: tuple = field(init=False)

    def __post_init__(self) -> None:
        super().__post_init__()
        self.outpaint_size = (_floor8(int(self.width_scale * self.width)),
                              _floor8(int(self.height_scale * self.height)))


def convert_image_to_frames(images: torch.Tensor) -> list:
    """IMAGE [T,H,W,3] float 0..1 -> list of PIL RGB frames (uint8 by truncation)."""
    arr = (images.detach().cpu().numpy() * 255).clip(0, 255).astype(np.uint8)
    return [Image.fromarray(a) for a in arr]


def convert_mask_to_frames(masks: torch.Tensor) -> list:
    out = []
    for m in masks:
        m = m.detach().cpu()
        if m.dtype == torch.float32:
            m = (m * 255).clamp(0, 255).byte()
        out.append(Image.fromarray(m.numpy(), mode="L"))
    return out


def resize_images(images: list, config: ImageConfig) -> list:
    if tuple(config.process_size) == tuple(config.input_size):
        return images
    return [im.resize(config.process_size) for im in images]


def _grow(mask_u8: np.ndarray, iterations: int) -> np.ndarray:
    if iterations > 0:
        return scipy.ndimage.binary_dilation(mask_u8, iterations=iterations).astype(np.uint8)
    return (mask_u8 > 0.1).astype(np.uint8)


def read_masks(masks: torch.Tensor, config: ImageConfig):
    """-> (flow_masks, masks_dilated) as lists of PIL 'L' images with values {0,255}."""
    flow, dil = [], []
    for im in resize_images(convert_mask_to_frames(masks), config):
        a = np.array(im.convert("L"))
        flow.append(Image.fromarray(_grow(a, config.flow_mask_dilates) * 255))
        dil.append(Image.fromarray(_grow(a, config.mask_dilates) * 255))
    if len(flow) == 1:
        flow, dil = flow * config.video_length, dil * config.video_length
    return flow, dil


def _stack_to_tensor(images: list) -> torch.Tensor:
    """list of PIL (RGB or L) -> float [T,C,H,W] in 0..1."""
    a = np.stack([np.asarray(im) if im.mode == "RGB" else np.asarray(im.convert("L"))[..., None] for im in images])
    return torch.from_numpy(np.ascontiguousarray(a.transpose(0, 3, 1, 2))).float().div(255)


def _tensorise(frames, flow_masks, masks_dilated, device):
    originals = [np.array(f) for f in frames]
    ft = (_stack_to_tensor(frames).unsqueeze(0) * 2 - 1).to(device)
    fm = _stack_to_tensor(flow_masks).unsqueeze(0).to(device)
    md = _stack_to_tensor(masks_dilated).unsqueeze(0).to(device)
    return ft, fm, md, originals


def prepare_frames_and_masks(frames: list, mask: torch.Tensor, config: ImageConfig, device):
    frames = resize_images(frames, config)
    flow_masks, masks_dilated = read_masks(mask, config)
    return _tensorise(frames, flow_masks, masks_dilated, device)


def extrapolation(frames: list, config: ImageOutpaintConfig):
    """Outpainting canvas: frames centred on a zero canvas, side-band masks (reference :200-252).

    The flow mask keeps a 4-px inset into the known region on sides whose band is wider than 10 px."""
    frames = resize_images(frames, config)
    rw, rh = frames[0].size
    pw, ph = config.outpaint_size
    x0, y0 = int((pw - rw) / 2), int((ph - rh) / 2)
    canvas = []
    for f in frames:
        c = np.zeros((ph, pw, 3), dtype=np.uint8)
        c[y0:y0 + rh, x0:x0 + rw] = f
        canvas.append(Image.fromarray(c))
    ih, iw = (4 if y0 > 10 else 0), (4 if x0 > 10 else 0)
    band = np.ones((ph, pw), dtype=np.uint8)
    band[y0 + ih:y0 + rh - ih, x0 + iw:x0 + rw - iw] = 0
    flow_mask = Image.fromarray(band * 255)
    band[y0:y0 + rh, x0:x0 + rw] = 0
    dil_mask = Image.fromarray(band * 255)
    n = config.video_length
    return canvas, [flow_mask] * n, [dil_mask] * n


def prepare_frames_and_masks_for_outpaint(frames, flow_masks, masks_dilated, device):
    return _tensorise(frames, flow_masks, masks_dilated, device)


def outpaint_tensors(images: torch.Tensor, config: ImageOutpaintConfig, device):
    """Tensor version of convert_image_to_frames + extrapolation + prepare_frames_and_masks_for_outpaint for the
    no-resize case (process_size == input_size): the same integer semantics (uint8 truncation, zero canvas, side-band
    masks with the 4-px inset) with a handful of tensor ops on `device` instead of per-frame PIL/numpy work.
    -> frames [1,T,3,H',W'] in [-1,1], flow_masks / masks_dilated [1,T,1,H',W'] in {0,1}, originals uint8 [T,H',W',3]."""
    assert tuple(config.process_size) == tuple(config.input_size)
    T, rh, rw = images.shape[0], images.shape[1], images.shape[2]
    pw, ph = config.outpaint_size
    x0, y0 = int((pw - rw) / 2), int((ph - rh) / 2)
    q = (images.to(device=device, dtype=torch.float32) * 255).clamp_(0, 255).to(torch.uint8)   # truncation
    canvas = torch.zeros(T, ph, pw, 3, dtype=torch.uint8, device=device)
    canvas[:, y0:y0 + rh, x0:x0 + rw] = q
    ih, iw = (4 if y0 > 10 else 0), (4 if x0 > 10 else 0)
    band = torch.ones(ph, pw, dtype=torch.float32, device=device)
    band[y0 + ih:y0 + rh - ih, x0 + iw:x0 + rw - iw] = 0
    flow = band.clone()
    band[y0:y0 + rh, x0:x0 + rw] = 0
    ft = canvas.permute(0, 3, 1, 2).to(torch.float32).div_(255).unsqueeze(0) * 2 - 1
    fm = flow.view(1, 1, 1, ph, pw).expand(1, T, 1, ph, pw).contiguous()
    md = band.view(1, 1, 1, ph, pw).expand(1, T, 1, ph, pw).contiguous()
    return ft, fm, md, canvas


def handle_output(composed_frames, flow_masks: torch.Tensor, masks_dilated: torch.Tensor):
    """uint8 HWC frames -> IMAGE [T,H,W,3] float32 CPU; masks squeezed to [T,H,W] (reference :276-290)."""
    if isinstance(composed_frames, torch.Tensor):
        imgs = composed_frames.to(torch.float32).div(255.0).cpu()
    else:
        imgs = torch.stack([torch.from_numpy(f.astype(np.float32) / 255.0) for f in composed_frames])
    return imgs, flow_masks.squeeze(), masks_dilated.squeeze()
