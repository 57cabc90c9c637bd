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
    "BlipImageProcessorPil": "blip",      # transformers >= 5: the PIL-backed class of the same processor
    "BlipImageProcessorFast": "blip",
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
    ``(M, H, W)``; ``mode`` in {"raw", "sigmoid", "mask", "mask_bits"} fuses the consumer ("mask_bits": one bit per pixel,
    ``(M, H, ceil(W / 32))`` int32 -- what a 128 x 1024 open-vocabulary sweep can afford to keep)
    (torch.sigmoid, segmentation_utils.py:225; ``> t``, :258).
    """
    kind = processor_kind(image_processor)
    kw = interpolate_params(origin_size, kind)
    m = {"raw": _lib.RZ_UP_RAW, "sigmoid": _lib.RZ_UP_SIGMOID, "mask": _lib.RZ_UP_MASK,
         "mask_bits": _lib.RZ_UP_MASK_BITS}[mode]
    s = similarity_scores
    out = ops.upsample_maps(s.float(), (int(origin_size[0]), int(origin_size[1])), mode=m,
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
    pts = ops.upsample_maps(s.float(), (int(image_size[0]), int(image_size[1])),
                            mode=_lib.RZ_UP_ARGMAX, **kw)
    if single:
        x, y = pts[0].tolist()
        return (x, y)
    return pts


# ----------------------------------------------------------------------------- fused consumers
def threshold_logits(thresholds, device) -> torch.Tensor:
    """Probability thresholds -> score-domain thresholds: sigmoid(v) > t  <=>  v > logit(t).
    t <= 0 maps to -ln(FLT_MAX) (below it fp32 sigmoid is exactly 0, so `prob > 0` is false, as for
    the reference's -999 fill), t >= 1 to +inf (never exceeded)."""
    import math
    out = []
    for t in thresholds:
        t = float(t)
        out.append(-88.72283905206835 if t <= 0.0 else (math.inf if t >= 1.0 else math.log(t / (1.0 - t))))
    return torch.tensor(out, dtype=torch.float32, device=device)


@torch.no_grad()
def dice_sweep_stats(similarity_scores: torch.Tensor, masks: torch.Tensor, origin_size, image_processor,
                     thresholds=None):
    """Sufficient statistics of the reference's Dice threshold sweep and specificity
    (exp/cxr_pt/inference/segmentation_utils.py:255-261, 136-158) for a batch of maps, computed on
    the GPU from the patch-grid scores: the pixel maps are never written, copied or re-thresholded.

    similarity_scores (..., P*P); masks uint8 (M, H, W).  Returns a dict of tensors indexed
    [map, threshold]: ``pred`` = |prob > t|, ``inter`` = |(prob > t) & mask|, plus ``gt`` = |mask|
    per map, ``max_prob`` per map and ``thresholds``.  Dice_t = 2 inter / (pred + gt); a negative
    image is a true negative at t iff max_prob <= t."""
    import numpy as np
    thresholds = np.arange(0, 1.01, 0.01) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    kind = processor_kind(image_processor)
    kw = interpolate_params(origin_size, kind)
    dev = similarity_scores.device
    thr = threshold_logits(thresholds, dev)
    ha, hg, mx = ops.map_threshold_stats(similarity_scores, (int(origin_size[0]), int(origin_size[1])), thr,
                                         gt_masks=masks, **kw)
    # bin k = pixels above exactly k thresholds  ->  above threshold j  <=>  k > j
    pred = ha.flip(1).cumsum(1).flip(1)[:, 1:]
    inter = hg.flip(1).cumsum(1).flip(1)[:, 1:]
    return {"pred": pred, "inter": inter, "gt": hg.sum(1), "max_prob": torch.sigmoid(mx),
            "thresholds": torch.as_tensor(thresholds)}


def best_dice_and_specificity(pos_stats, neg_stats=None, aggregate: str = "samplewise"):
    """The reference's selection loop (segmentation_utils.py:255-270) on the statistics above.

    The reference scores each threshold with ``torchmetrics.segmentation.DiceScore(num_classes=1)``
    (``torchmetrics==1.6.1``, requirements.txt:242; defaults ``average="micro"``, ``include_background=
    True``): Dice is computed PER IMAGE -- ``2 |P & G| / (|P| + |G|)``, 1.0 when the denominator is 0 --
    and then nan-averaged over the images.  That is ``aggregate="samplewise"`` (default).
    ``aggregate="pooled"`` sums the confusion counts over the batch before the ratio (a different
    metric, kept as an explicit option).  Returns the first threshold maximising the Dice of the
    positive maps (``if cur_dice > best_dice``) and the image-level specificity of the negative maps
    at that threshold."""
    pred, inter = pos_stats["pred"].double(), pos_stats["inter"].double()          # (maps, thresholds)
    gt = pos_stats["gt"].double()
    if aggregate == "samplewise":
        den = pred + gt[:, None]
        per_map = torch.where(den > 0, 2.0 * inter / den.clamp_min(1.0), torch.ones_like(den))
        dice = per_map.nanmean(dim=0)
    elif aggregate == "pooled":
        dice = 2.0 * inter.sum(0) / (pred.sum(0) + gt.sum()).clamp_min(1.0)
    else:
        raise ValueError(f"aggregate must be 'samplewise' or 'pooled', got {aggregate!r}")
    # `best_dice = 0.0; if cur_dice > best_dice`: the first strict maximum (nothing selected if all are 0)
    j = int(dice.argmax())
    t = float(pos_stats["thresholds"][j])
    out = {"dice": float(dice[j]), "best_threshold": t}
    if neg_stats is not None:
        out["specificity"] = float((neg_stats["pred"][:, j] == 0).double().mean())
    return out


# ----------------------------------------------------------------------------- surface 1
def _load_image(image):
    """Accept a path, a PIL image or an (H, W[, C]) uint8 array/tensor; return (PIL RGB, (H, W))."""
    from PIL import Image
    import numpy as np
    if isinstance(image, (str, bytes)) or hasattr(image, "__fspath__"):
        image = Image.open(image)
    elif torch.is_tensor(image):
        image = Image.fromarray(image.detach().cpu().numpy())
    elif isinstance(image, np.ndarray):
        image = Image.fromarray(image)
    width, height = image.size
    return image, (height, width)


@torch.no_grad()
def extract_similarity_map(image, text, model, image_processor, tokenizer):
    """Drop-in for the reference's six ``extract_similarity_map`` copies
    (exp/cxr_pt/inference/visualization/attention_map_base.py:12-42): pixel-level map at the
    cos / tau scale, shape (H, W) of the ORIGINAL image.  ``text`` may be a list of prompts,
    in which case the result is (N, H, W) from one launch."""
    import numpy as np
    pil, image_size = _load_image(image)
    pix = image_processor(pil)
    pixel_values = torch.as_tensor(np.array(pix["pixel_values"]), dtype=torch.float32).to(model.device)
    tokenized = tokenizer(text, padding=True, truncation=True, return_tensors="pt").to(model.device)
    out = model.compute_logits(pixel_values, [tokenized])
    scores = out["similarity_scores"]                     # (1, N, P*P)
    maps = interpolate_similarity_scores(scores, image_size, image_processor)   # strided views are fine
    return maps.squeeze(0) if maps.shape[0] == 1 else maps


@torch.no_grad()
def model_inference(image, text, tokenizer, image_processor, model):
    """``utils.model_inference(image, text, tokenizer, image_processor, model)`` of the
    reference's usage snippet (README.md:66, 104-111) -> ``(similarity_prob, similarity_map)``.

    The hub-side body is not in the reference checkout; this follows
    ``extract_similarity_map`` (same preprocessing, ``model.compute_logits``, bilinear
    upsample to the original image size) and returns similarity_prob = sigmoid(logits) and
    the pixel-level similarity_map at the cos / tau scale, both computed by the fused CUDA
    path: the similarity + pooling kernel, then ONE upsample launch for all prompts.
    """
    import numpy as np
    pil, image_size = _load_image(image)
    pix = image_processor(pil)
    pixel_values = torch.as_tensor(np.array(pix["pixel_values"]), dtype=torch.float32).to(model.device)
    tokenized = tokenizer(text, padding=True, truncation=True, return_tensors="pt").to(model.device)
    out = model.compute_logits(pixel_values, [tokenized])
    similarity_prob = torch.sigmoid(out["logits"])        # (1, N)
    scores = out["similarity_scores"]
    similarity_map = interpolate_similarity_scores(scores, image_size, image_processor)
    if similarity_map.shape[0] == 1:                      # single prompt: scalar prob, (H, W) map
        return similarity_prob.reshape(()), similarity_map.squeeze(0)
    return similarity_prob.squeeze(0), similarity_map


class SyntheticTokenizer:
    """Stand-in tokenizer for offline runs (no vocabulary files on disk): deterministic ids
    from a hash of the words, HF ``__call__`` convention returning a dict with ``.to()``."""

    def __init__(self, vocab_size: int = 30527, max_length: int = 32):
        self.vocab_size, self.max_length = vocab_size, max_length

    class _Enc(dict):
        def to(self, device):
            return SyntheticTokenizer._Enc({k: v.to(device) for k, v in self.items()})

    def __call__(self, text, padding=True, truncation=True, return_tensors="pt", **kw):
        import zlib
        texts = [text] if isinstance(text, str) else list(text)
        rows = []
        for t in texts:
            ids = [0] + [4 + zlib.crc32(w.lower().encode()) % (self.vocab_size - 4) for w in t.split()] + [2]
            rows.append(ids[: self.max_length])
        m = max(len(r) for r in rows)
        ids = torch.ones((len(rows), m), dtype=torch.long)
        mask = torch.zeros((len(rows), m), dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return SyntheticTokenizer._Enc(input_ids=ids, attention_mask=mask)


# ----------------------------------------------------------------------------- result formats
# the on-disk formats of the zero-shot evaluation, so that a whole-dataset
# run through the CUDA path leaves the same files as the reference's evaluators.
def process_class_prompts(text_prompt: Dict, tokenizer, model) -> Dict:
    """inference/utils.py:40-66: class ``i`` -> its first prompt (``text_prompt[str(i)][0]``) and the
    negated prompt (``"There is" -> "There is no"``), both tokenised as one padded batch on the
    model's device."""
    pos = [text_prompt[str(i)][0] for i in range(len(text_prompt))]
    neg = [p.replace("There is", "There is no") for p in pos]
    tok = lambda texts: tokenizer(texts, padding=True, truncation=True, return_tensors="pt").to(model.device)
    return {"encoded_key_phrases": tok(pos), "encoded_negative_phrases": tok(neg)}


@torch.no_grad()
def calculate_similarities(pixel_batches, text_batch: Dict, model):
    """inference/utils.py:69-107 without the DataLoader: ``pixel_batches`` is any iterable of
    ``pixel_values`` batches; returns the (images, classes) float32 numpy matrix of ``logits`` the
    reference concatenates (``compute_logits(...)["logits"]`` per batch, :94-103)."""
    rows = []
    for image in pixel_batches:
        image = image.to(model.device)
        rows.append(model.compute_logits(pixel_values=image,
                                         encoded_key_phrases=[text_batch["encoded_key_phrases"]],
                                         encoded_negative_phrases=[text_batch.get("encoded_negative_phrases")])["logits"])
    return torch.cat(rows, dim=0).float().detach().cpu().numpy()


def save_similarities_csv(similarities, save_root_dir: str, sel_dataset: str) -> str:
    """``pd.DataFrame(similarities).to_csv(<save_root_dir>/<dataset>.csv, index=False)``
    (inference/utils.py:213-215): header ``0,1,...``, one row per image, pandas' float formatting."""
    import os
    import pandas as pd
    path = os.path.join(save_root_dir, sel_dataset) + ".csv"
    pd.DataFrame(similarities).to_csv(path, index=False)
    return path


def save_result_json(result: Dict, save_root_dir: str) -> str:
    """``save_json(result, <save_root_dir>/result.json)`` (inference/inference.py:62, 109, 166;
    common/utils.py:123-125: utf-8, indent 2)."""
    import json
    import os
    os.makedirs(save_root_dir, exist_ok=True)
    path = os.path.join(save_root_dir, "result.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(result, f, indent=2)
    return path
