"""Consumers of the patch-grid similarity map, mirroring the reference's helpers.

``interpolate_similarity_scores`` / ``get_grounding_point`` keep the reference's names,
argument meaning and return shapes (exp/cxr_pt/inference/segmentation_utils.py:36-122,
exp/cxr_pt/inference/grounding_utils.py:166-261) but run as ONE fused CUDA launch for any
number of maps instead of one ``F.interpolate`` per (image, prompt) plus a host copy.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib, ops

_KIND_BY_CLASSNAME = {
    "AspectRatioBlipImageProcessor": "aspect_blip",
    "BlipImageProcessor": "blip",
    "BitImageProcessor": "bit",
    "M3AEImageProcessor": "m3ae",
}


def processor_kind(image_processor) -> str:
    """Map an image-processor object (or a kind string) to the reference's isinstance branch.

    The reference tests AspectRatioBlipImageProcessor before its base BlipImageProcessor
    (segmentation_utils.py:41, :62); walking the MRO most-derived-first reproduces that.
    """
    if isinstance(image_processor, str):
        if image_processor not in _KIND_BY_CLASSNAME.values():
            raise NotImplementedError(f"Image processor {image_processor} is not supported")
        return image_processor
    for cls in type(image_processor).__mro__:
        if cls.__name__ in _KIND_BY_CLASSNAME:
            return _KIND_BY_CLASSNAME[cls.__name__]
    raise NotImplementedError(f"Image processor {type(image_processor)} is not supported")


def interpolate_params(origin_size: Tuple[int, int], kind: str) -> Dict:
    """Resize size / paste offset / fill that turn each processor branch into one kernel call."""
    height, width = int(origin_size[0]), int(origin_size[1])
    if kind == "blip":            # segmentation_utils.py:62-70
        return dict(interp_hw=(height, width), offset=(0, 0), fill=-999.0)
    if kind == "aspect_blip":     # :41-60  resize to a square of the long side, crop the centre
        side = max(height, width)
        return dict(interp_hw=(side, side), offset=(-((side - height) // 2), -((side - width) // 2)),
                    fill=-999.0)
    if kind == "bit":             # :72-91  resize to the short side, paste centred, -999 elsewhere
        side = min(height, width)
        return dict(interp_hw=(side, side), offset=((height - side) // 2, (width - side) // 2),
                    fill=-999.0)
    if kind == "m3ae":            # :92-121 224/256 centre crop inside the padded square
        side = max(height, width)
        crop = int(side * 224 / 256)
        off = (side - crop) // 2
        return dict(interp_hw=(crop, crop),
                    offset=(off - (side - height) // 2, off - (side - width) // 2), fill=-999.0)
    raise NotImplementedError(kind)


def interpolate_similarity_scores(similarity_scores: torch.Tensor, origin_size, image_processor,
                                  mode: str = "raw", threshold: float = 0.5) -> torch.Tensor:
    """Drop-in for the reference function: ``similarity_scores`` (P*P,) -> (1, H, W).

    Extension: a 2-d ``(M, P*P)`` input upsamples M maps in one launch and returns
    ``(M, H, W)``; ``mode`` in {"raw", "sigmoid", "mask"} fuses the consumer
    (torch.sigmoid, segmentation_utils.py:225; ``> t``, :258).
    """
    kind = processor_kind(image_processor)
    kw = interpolate_params(origin_size, kind)
    m = {"raw": _lib.RZ_UP_RAW, "sigmoid": _lib.RZ_UP_SIGMOID, "mask": _lib.RZ_UP_MASK}[mode]
    s = similarity_scores
    single = s.dim() == 1
    s2 = s.reshape(1, -1) if single else s.reshape(-1, s.shape[-1])
    out = ops.upsample_maps(s2.float(), (int(origin_size[0]), int(origin_size[1])), mode=m,
                            threshold=threshold, **kw)
    return out  # (1, H, W) for a single map, like the reference's squeeze(1)


def get_grounding_point(similarity_score: torch.Tensor, image_size, image_processor):
    """Drop-in for grounding_utils.get_grounding_point: (x, y) of the global maximum.

    A 2-d input returns an int64 tensor (M, 2) without leaving the device.
    """
    kind = processor_kind(image_processor)
    kw = interpolate_params(image_size, kind)
    s = similarity_score
    single = s.dim() == 1
    s2 = s.reshape(1, -1) if single else s.reshape(-1, s.shape[-1])
    pts = ops.upsample_maps(s2.float(), (int(image_size[0]), int(image_size[1])),
                            mode=_lib.RZ_UP_ARGMAX, **kw)
    if single:
        x, y = pts[0].tolist()
        return (x, y)
    return pts
