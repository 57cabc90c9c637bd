"""GPU image preprocessing of the zero-shot evaluators.

``collate_fn(batch, image_processor)`` keeps the reference's name, arguments and result
(exp/cxr_pt/inference/dataset.py:31-51: a float tensor ``(B, 3, H, W)`` of ``pixel_values``) but the
work -- min-max stretch to uint8 (``cv2.normalize``), RGB conversion, the image processor's bicubic
resize, rescale and normalisation (exp/cxr_pt/model/processing.py:85-101) -- runs as four CUDA
launches per group of same-sized images (``rz_preprocess_images``), bit-identical to the host chain.
The raw pixels cross PCIe once, in their source type (1-2 bytes per pixel instead of the 12 bytes per
output pixel of the float tensor the reference's DataLoader workers produce).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import RzError

_BICUBIC = 3          # PIL.Image.BICUBIC / PILImageResampling.BICUBIC


def processor_params(image_processor) -> Dict:
    """size / mean / std / rescale of a Blip-style HF image processor (the branch RadZero uses,
    processing.py:85-101).  Anything the kernel does not restate raises instead of approximating."""
    from .inference import processor_kind
    if processor_kind(image_processor) != "blip":
        raise NotImplementedError("GPU preprocessing implements the BlipImageProcessor branch (resize to a "
                                  "fixed square, no crop); got %s" % type(image_processor).__name__)
    g = lambda k, d=None: getattr(image_processor, k, d)
    if int(g("resample", _BICUBIC)) != _BICUBIC:
        raise NotImplementedError("only bicubic resampling (the BlipImageProcessor default) is implemented")
    for flag in ("do_resize", "do_rescale", "do_normalize", "do_convert_rgb"):
        if g(flag, True) is False:
            raise NotImplementedError(f"{flag}=False is not implemented on the GPU path")
    size = g("size")
    size = size if isinstance(size, dict) else dict(size)
    return dict(out_hw=(int(size["height"]), int(size["width"])), mean=list(g("image_mean")),
                std=list(g("image_std")), rescale_factor=float(g("rescale_factor", 1 / 255)))


def _as_array(item) -> np.ndarray:
    a = np.array(item)                                   # PIL image or array, as dataset.py:38 does
    if a.dtype == np.float64:
        a = a.astype(np.float32)                         # cv2 converts to float before the stretch
    if a.dtype == np.int64:
        a = a.astype(np.int32)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.ndim == 3 and a.shape[-1] == 1:
        a = a[..., 0]
    if a.ndim == 3 and a.shape[-1] == 4:
        raise NotImplementedError("RGBA sources are not handled on the GPU path")
    if a.ndim not in (2, 3):
        raise RzError(f"image must be (H, W) or (H, W, 3), got {a.shape}")
    return np.ascontiguousarray(a)


@torch.no_grad()
def preprocess_arrays(arrays: Sequence[np.ndarray], device, *, out_hw: Tuple[int, int], mean, std,
                      rescale_factor: float = 1 / 255, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Raw host images (any mix of sizes / dtypes) -> pixel_values (B, 3, h, w) on ``device``.
    Images of the same shape and dtype share one upload and one set of launches."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RzError("radzero_b200 preprocessing runs on CUDA only (there is no CPU fallback)")
    out = torch.empty((len(arrays), 3, out_hw[0], out_hw[1]), dtype=out_dtype, device=device)
    groups: Dict[tuple, List[int]] = {}
    for i, a in enumerate(arrays):
        groups.setdefault((a.shape, a.dtype.str), []).append(i)
    for (shape, _), idx in groups.items():
        host = torch.from_numpy(np.stack([arrays[i] for i in idx]))
        raw = host.pin_memory().to(device, non_blocking=True)
        pv = ops.preprocess_images(raw, out_hw, mean=mean, std=std, rescale_factor=rescale_factor,
                                   out_dtype=out_dtype)
        if len(groups) == 1:
            return pv
        out[torch.as_tensor(idx, device=device)] = pv
    return out


def collate_fn(batch, image_processor, device="cuda", out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Drop-in for exp/cxr_pt/inference/dataset.py:31-51; returns the ``pixel_values`` tensor on ``device``."""
    kw = processor_params(image_processor)
    return preprocess_arrays([_as_array(item) for item in batch], device, out_dtype=out_dtype, **kw)
