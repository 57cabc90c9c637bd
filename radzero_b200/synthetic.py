"""Synthetic chest-X-ray-shaped inputs for the VL-CABS path.

There are no datasets or checkpoints offline, so every test and benchmark uses inputs
of the reference's shapes drawn here: vision tokens ``(B, 1370, 768)`` = N(0,1) noise
plus a low-rank shared component (so cosines span roughly [-0.3, 0.7] instead of the
|cos| < 0.16 of iid noise, which would hide precision bugs), sentence embeddings
``(N, 768)`` as noisy copies of the shared directions, LayerNorm gamma ~ U(0.5, 1.5),
beta ~ U(-0.2, 0.2) (not the 1/0 init, to exercise LN), and log_tau = log(0.07)
(exp/cxr_pt/configs/radzero.yaml:43).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

L_TOKENS = 1370   # CLS + 37*37, DINOv2-base patch 14 at 518x518 (radzero.yaml:19)
GRID = 37
HIDDEN = 768      # radzero.yaml:31,40
LOG_TAU_INIT = math.log(0.07)

CHEXPERT_FINDINGS = [  # external/CARZero/inference.py:306-319 (prompt = "There is {finding}")
    "atelectasis", "cardiomegaly", "consolidation", "edema", "enlarged cardiomediastinum",
    "fracture", "lung lesion", "lung opacity", "no finding", "pleural effusion",
    "pleural other", "pneumonia", "pneumothorax", "support devices",
]


def make_inputs(batch: int, n_text: int, *, tokens_per_image: int = L_TOKENS,
                hidden: int = HIDDEN, seed: int = 42, device="cpu",
                dtype=torch.float32, rank_shared: int = 8):
    """Return (vision_tokens (B,L,D), text (N,D), gamma (D,), beta (D,), log_tau (1,))."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    f32 = dict(device=device, dtype=torch.float32, generator=gen)
    dirs = torch.randn(rank_shared, hidden, **f32)
    tokens = torch.randn(batch, tokens_per_image, hidden, **f32)
    coef = 0.5 * torch.randn(batch, tokens_per_image, rank_shared, **f32)
    tokens.add_(coef @ dirs)
    pick = torch.arange(n_text, device=device) % rank_shared
    text = dirs[pick] + 0.3 * torch.randn(n_text, hidden, **f32)
    gamma = 0.5 + torch.rand(hidden, **f32)
    beta = -0.2 + 0.4 * torch.rand(hidden, **f32)
    log_tau = torch.full((1,), LOG_TAU_INIT, device=device, dtype=torch.float32)
    return (tokens.to(dtype), text.to(dtype), gamma.to(dtype), beta.to(dtype), log_tau)


def sentence_counts(batch: int, *, seed: int = 42, lo: int = 3, hi: int = 9,
                    fixed: int | None = None) -> List[int]:
    """Sentences per image: n_i ~ U{lo..hi} (mean 6) or a fixed count (C4 variants)."""
    if fixed is not None:
        return [int(fixed)] * batch
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed + 1000003)
    return torch.randint(lo, hi + 1, (batch,), generator=gen).tolist()


def group_map_from_counts(counts: Sequence[int], first_image: int = 0, device="cpu") -> torch.Tensor:
    """group_map[j] = global image index of sentence j (losses.py:148-151), contiguous."""
    idx = torch.arange(len(counts), device=device) + first_image
    return torch.repeat_interleave(idx, torch.as_tensor(list(counts), device=device))


def synthetic_cxr(height: int = 1024, width: int = 1024, seed: int = 42) -> torch.Tensor:
    """One synthetic CXR: uint8 grayscale, smooth gradient + noise (config C1)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    yy = torch.linspace(0, 1, height).unsqueeze(1)
    xx = torch.linspace(0, 1, width).unsqueeze(0)
    img = 0.6 * torch.sin(3.0 * yy) * torch.cos(2.0 * xx) + 0.3 * yy + 0.1 * torch.randn(
        height, width, generator=gen)
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255.0).round().to(torch.uint8)


def checksum(t: torch.Tensor) -> Tuple[float, float]:
    """(sum, abs-sum) in float64 -- used by the golden fixtures to detect RNG drift."""
    d = t.detach().double()
    return float(d.sum().item()), float(d.abs().sum().item())


def align_layer_weights(seed: int = 42, layers: int = 2, hidden: int = HIDDEN, mlp_ratio: int = 4):
    """Seeded weights of the AlignTransformer's Dinov2 layers, keyed like ``Dinov2Layer.state_dict()``.

    Linear weights N(0, 0.02) (the HF ``_init_weights`` the reference trains from,
    exp/cxr_pt/model/common_layers.py:13-28), but query / key at std 0.08 so that the attention is
    not a uniform average, small non-zero biases, LayerNorm gamma ~ U(0.5, 1.5), beta ~ U(-0.2, 0.2)
    and LayerScale lambda ~ U(0.5, 1.5) (not the 1 / 0 inits, to exercise every term).
    """
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed + 7919)
    n = lambda *s, std=0.02: std * torch.randn(*s, generator=gen)
    u = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(*s, generator=gen)
    out = []
    for _ in range(layers):
        w = {}
        for nm in ("norm1", "norm2"):
            w[f"{nm}.weight"] = u(0.5, 1.5, hidden)
            w[f"{nm}.bias"] = u(-0.2, 0.2, hidden)
        for nm, std in (("query", 0.08), ("key", 0.08), ("value", 0.02)):
            w[f"attention.attention.{nm}.weight"] = n(hidden, hidden, std=std)
            w[f"attention.attention.{nm}.bias"] = n(hidden, std=0.05)
        w["attention.output.dense.weight"] = n(hidden, hidden)
        w["attention.output.dense.bias"] = n(hidden, std=0.05)
        w["layer_scale1.lambda1"] = u(0.5, 1.5, hidden)
        w["mlp.fc1.weight"] = n(mlp_ratio * hidden, hidden)
        w["mlp.fc1.bias"] = n(mlp_ratio * hidden, std=0.05)
        w["mlp.fc2.weight"] = n(hidden, mlp_ratio * hidden)
        w["mlp.fc2.bias"] = n(hidden, std=0.05)
        w["layer_scale2.lambda1"] = u(0.5, 1.5, hidden)
        out.append(w)
    return out


def build_align_encoder(seed: int = 42, layers: int = 2, device="cpu"):
    """A transformers ``Dinov2Encoder`` (the reference's AlignTransformer body,
    align_transformers.py:27-28) loaded with :func:`align_layer_weights`."""
    from transformers import Dinov2Config
    from transformers.models.dinov2.modeling_dinov2 import Dinov2Encoder
    cfg = Dinov2Config(hidden_size=HIDDEN, num_hidden_layers=layers, num_attention_heads=12)
    enc = Dinov2Encoder(cfg)
    for layer, w in zip(enc.layer, align_layer_weights(seed, layers)):
        missing = layer.load_state_dict(w, strict=True)
        del missing
    return enc.to(device).eval()
