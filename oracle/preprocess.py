"""CPU oracle of the zero-shot evaluators' image preprocessing.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates, in numpy integer / float arithmetic,
what the reference does to every image before the vision encoder:

  1. ``collate_fn`` (exp/cxr_pt/inference/dataset.py:31-51): ``cv2.normalize(np.array(item), None, 0, 255,
     NORM_MINMAX, dtype=CV_8U)`` -- min-max stretch to uint8.  The arithmetic is OpenCV's, a third-party
     dependency the reference does not vendor (``opencv-python==4.9.0.80``, requirements.txt:144):
     ``scale = 255 * (1 / (max - min))`` in double (0 when max - min <= DBL_EPSILON), ``shift = -min *
     scale``; then ``convertTo(CV_8U, scale, shift)`` = ``saturate_cast<uchar>(rint(fma((float)src,
     (float)scale, (float)shift)))`` -- a single-rounding float FMA, round-half-to-even.
  2. the image processor (exp/cxr_pt/model/processing.py:85-101: a ``BlipImageProcessor`` resized to
     518, ``transformers==4.39.3``): convert to RGB, PIL bicubic resize of the uint8 image, ``* 1/255``
     (in float64, cast to float32), ``(x - mean) / std`` in float32, channels first.
     The resize is Pillow's (``pillow==10.2.0``, src/libImaging/Resample.c): separable, horizontal pass
     first, antialiased (filter support scaled by the down-sampling factor), 8-bit fixed point --
     coefficients rounded to 22 fractional bits, accumulator started at 2^21, result ``>> 22`` clipped
     to [0, 255], the intermediate image stored as uint8.

Pinned (tests/test_oracle_preprocess.py) against cv2, Pillow and transformers' PIL-backed
``BlipImageProcessor`` executed in the build container -- bit-exact -- and frozen into
tests/golden/preprocess_golden.npz by tests/golden/make_preprocess_golden.py.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

__all__ = ["minmax_to_uint8", "bicubic_coeffs", "resize_bicubic_u8", "normalize_lut", "preprocess_image",
           "OPENAI_CLIP_MEAN", "OPENAI_CLIP_STD"]

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # BlipImageProcessor defaults
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2                                # Resample.c


def minmax_to_uint8(x: np.ndarray) -> np.ndarray:
    """cv2.normalize(x, None, 0, 255, NORM_MINMAX, dtype=CV_8U) -- dataset.py:37-41."""
    smin, smax = float(x.min()), float(x.max())
    scale = 255.0 * (1.0 / (smax - smin) if (smax - smin) > np.finfo(np.float64).eps else 0.0)
    shift = 0.0 - smin * scale
    a, b = np.float32(scale), np.float32(shift)
    # fused multiply-add in float: the product is exact in double (24 x 24 bit mantissas), one rounding
    v = (x.astype(np.float32).astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc of Resample.c for the full-image box:
    returns (xmin int32 [out], count int32 [out], weights int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        xmin[xx], cnt[xx] = lo, n
        for x, v in enumerate(w):
            kk[xx, x] = int(v * (1 << PRECISION_BITS) + (-0.5 if v < 0 else 0.5))     # C truncation toward zero
    return xmin, cnt, kk


def _resample_axis(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8bpc pass along the LAST axis."""
    xmin, cnt, kk = bicubic_coeffs(img.shape[-1], out_size)
    out = np.empty(img.shape[:-1] + (out_size,), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n = int(cnt[xx])
        acc = (src[..., xmin[xx]:xmin[xx] + n] * kk[xx, :n].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """PIL ``Image.resize((W, H), BICUBIC)`` of a uint8 image (H, W) or (H, W, C)."""
    h_out, w_out = out_hw
    x = img if img.ndim == 2 else np.moveaxis(img, -1, 0)          # (..., H, W)
    if x.shape[-1] != w_out:
        x = _resample_axis(x, w_out)                                # horizontal pass first
    if x.shape[-2] != h_out:
        x = np.swapaxes(_resample_axis(np.swapaxes(x, -1, -2), h_out), -1, -2)
    return x if img.ndim == 2 else np.moveaxis(x, 0, -1)


def normalize_lut(mean: Sequence[float] = OPENAI_CLIP_MEAN, std: Sequence[float] = OPENAI_CLIP_STD,
                  rescale_factor: float = 1 / 255) -> np.ndarray:
    """(channels, 256) float32: what rescale + normalize map each uint8 level to (image_transforms.py)."""
    lv = (np.arange(256, dtype=np.uint8).astype(np.float64) * rescale_factor).astype(np.float32)
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    return ((lv[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def preprocess_image(raw: np.ndarray, size: Tuple[int, int] = (518, 518), mean=OPENAI_CLIP_MEAN,
                     std=OPENAI_CLIP_STD, rescale_factor: float = 1 / 255) -> np.ndarray:
    """One image (H, W) or (H, W, 3) of any cv2-supported dtype -> pixel_values (3, size_h, size_w) float32."""
    u8 = minmax_to_uint8(raw)
    r = resize_bicubic_u8(u8, size)
    lut = normalize_lut(mean, std, rescale_factor)
    if r.ndim == 2:
        return np.stack([lut[c][r] for c in range(3)])           # convert_to_rgb replicates the plane
    return np.stack([lut[c][r[..., c]] for c in range(3)])
