"""TEST INFRASTRUCTURE ONLY (see oracle/propainter_oracle.py): numpy restatement of Pillow's 8-bit bicubic resampler,
the algorithm behind the reference's ``Image.resize(size)`` calls (utils/image_utils.py:98-103 for the frames,
:142-150 for the mask images).  Pillow is a third-party dependency of the reference (unpinned; 12.2.0 in this image);
the restatement follows its published C source (src/libImaging/Resample.c: ``precompute_coeffs``,
``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``) and is pinned against the installed
Pillow by tests/test_host_logic.py.  The device kernel ``resize_axis_u8`` (csrc/kernels_pre.cu) implements the same
arithmetic and is compared with Pillow itself in the GPU tests.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int):
    """-> (kk int64 [out_size, ksize], bounds int64 [out_size, 2] = (first input index, count)); box = whole axis."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int64)
    bounds = np.zeros((out_size, 2), np.int64)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def _resample_axis(a: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    kk, b = coefficients(a.shape[0], out_size)
    out = np.zeros((out_size,) + a.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, xmax = b[xx]
        ss = np.full(a.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(xmax):
            ss += a[xmin + x] * kk[xx, x]
        out[xx] = np.clip(ss >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_u8(a: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """uint8 [H,W] or [H,W,C] -> [out_h,out_w(,C)]: horizontal pass first, then vertical, uint8 in between."""
    H, W = a.shape[:2]
    if out_w != W:
        a = _resample_axis(a, out_w, 1)
    if out_h != H:
        a = _resample_axis(a, out_h, 0)
    return a
