"""CPU oracle of the AlignTransformer forward.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates, in plain torch tensor algebra,
the forward of the reference's ``AlignTransformer`` (exp/cxr_pt/model/align_transformers.py:37-45):
``Dinov2Encoder(config)(vision_tokens)["last_hidden_state"]`` followed by the optional
``nn.LayerNorm``.  The encoder is third-party code, not vendored in the reference:
``transformers==4.39.0`` (reference requirements.txt), ``models/dinov2/modeling_dinov2.py`` --
``Dinov2Layer.forward`` (pre-norm block with LayerScale), ``Dinov2SelfAttention`` (12 heads of 64,
scores scaled by 1/sqrt(64), softmax, no mask, dropout 0 in eval), ``Dinov2MLP`` (fc1, erf-GELU, fc2).
Pinned by ``tests/test_oracle_align.py`` against the installed transformers' ``Dinov2Encoder``
(the same published algorithm; 5.5 here) and by the golden vectors in
``tests/golden/align_golden.npz`` generated from that module by ``tests/golden/make_align_golden.py``.

Weights are per-layer dicts keyed like ``Dinov2Layer.state_dict()``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

LAYER_NORM_EPS = 1e-6   # Dinov2Config.layer_norm_eps
HEADS = 12              # Dinov2Config.num_attention_heads default, kept by AlignTransformerConfig


def _ln(x, g, b, eps):
    # nn.LayerNorm: biased variance over the last dim
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def self_attention(h: torch.Tensor, w: Dict[str, torch.Tensor], heads: int = HEADS) -> torch.Tensor:
    """Dinov2SelfAttention.forward + eager_attention_forward: (B, L, D) -> (B, L, D)."""
    B, L, D = h.shape
    hd = D // heads
    p = "attention.attention."
    q = F.linear(h, w[p + "query.weight"], w[p + "query.bias"]).view(B, L, heads, hd).transpose(1, 2)
    k = F.linear(h, w[p + "key.weight"], w[p + "key.bias"]).view(B, L, heads, hd).transpose(1, 2)
    v = F.linear(h, w[p + "value.weight"], w[p + "value.bias"]).view(B, L, heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    a = torch.softmax(s, dim=-1)
    return torch.matmul(a, v).transpose(1, 2).reshape(B, L, D)


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def dinov2_layer(x: torch.Tensor, w: Dict[str, torch.Tensor], heads: int = HEADS,
                 eps: float = LAYER_NORM_EPS) -> torch.Tensor:
    """Dinov2Layer.forward: x + ls1*attn(norm1 x), then + ls2*mlp(norm2 .)."""
    w = {k: v.to(x.dtype) for k, v in w.items()}
    a = self_attention(_ln(x, w["norm1.weight"], w["norm1.bias"], eps), w, heads)
    a = F.linear(a, w["attention.output.dense.weight"], w["attention.output.dense.bias"])
    x = x + a * w["layer_scale1.lambda1"]
    m = _ln(x, w["norm2.weight"], w["norm2.bias"], eps)
    m = F.linear(gelu_erf(F.linear(m, w["mlp.fc1.weight"], w["mlp.fc1.bias"])),
                 w["mlp.fc2.weight"], w["mlp.fc2.bias"])
    return x + m * w["layer_scale2.lambda1"]


def align_transformer(vision_tokens: torch.Tensor, layers: Sequence[Dict[str, torch.Tensor]],
                      final_ln: Optional[Sequence[torch.Tensor]] = None, heads: int = HEADS) -> torch.Tensor:
    """AlignTransformer.forward (align_transformers.py:37-45)."""
    x = vision_tokens
    for w in layers:
        x = dinov2_layer(x, w, heads)
    if final_ln is not None:
        x = _ln(x, final_ln[0].to(x.dtype), final_ln[1].to(x.dtype), 1e-5)
    return x
