"""Mirror of the reference model surface for the VL-CABS path.

``CxrAlignModel`` keeps the reference's method names, argument meaning and output keys
(exp/cxr_pt/model/modeling.py: ``forward_vision_model`` :96-123, ``forward_text_model``
:125-211, ``forward`` :213-276, ``compute_logits`` :278-356) and adds the hub model's
``__call__``-style entry returning ``(similarity_prob, similarity_map)`` (README.md:104-111).
The encoders themselves are OUT OF SCOPE (stock HF modules):
they are passed in; everything from ``vision_tokens`` / ``text_features_wo_l2_norm`` onwards
runs on the hand-written CUDA path through ``RadZeroLoss``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .losses import RadZeroLoss

RADZERO_LOSS_CFG = dict(  # exp/cxr_pt/configs/radzero.yaml:37-46
    hidden_dim=768, use_vision_cls_token=True, attn_temperature=None, loss_temperature=0.07,
    text_features_l2_norm=False, mpnce_row_sum=False, mpnce_col_sum=False, sim_op="cos",
    use_layer_norm=True)


def mean_pooling(last_hidden_state: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Masked mean over tokens (exp/cxr_pt/model/text_encoders.py:32-41)."""
    m = attention_mask.unsqueeze(-1).to(last_hidden_state.dtype)
    return (last_hidden_state * m).sum(1) / m.sum(1).clamp(min=1e-9)


class CxrAlignModel(nn.Module):
    def __init__(self, vision_model: nn.Module, align_transformer: nn.Module, text_model: nn.Module,
                 loss_cfg: Optional[Dict] = None, compute_logits_type: str = "radzero"):
        super().__init__()
        self.vision_model = vision_model
        self.align_transformer = align_transformer
        self.text_model = text_model
        self.loss_fns = nn.ModuleDict({"RadZeroLoss": RadZeroLoss(**(loss_cfg or RADZERO_LOSS_CFG))})
        self.loss_ratio = {"RadZeroLoss": 1.0}
        self.compute_logits_type = compute_logits_type

    @property
    def device(self):
        return next(self.parameters()).device

    # ------------------------------------------------------------------ encoders (stock HF)
    def forward_vision_model(self, pixel_values):
        out = self.vision_model(pixel_values)
        vision_tokens = out["last_hidden_state"] if not torch.is_tensor(out) else out
        at = self.align_transformer(vision_tokens)
        vision_tokens = at["last_hidden_state"] if not torch.is_tensor(at) else at
        cls_token = vision_tokens[:, 0]
        patch_tokens = vision_tokens[:, 1:]
        image_features = F.normalize(torch.cat([cls_token, patch_tokens.mean(dim=1)], dim=1), p=2, dim=1)
        return {"vision_tokens": vision_tokens, "image_cls_token": cls_token,
                "image_patch_tokens": patch_tokens, "image_features": image_features}

    def _text_hidden(self, encoded_input):
        out = self.text_model(input_ids=encoded_input["input_ids"],
                              attention_mask=encoded_input["attention_mask"])
        return out["last_hidden_state"] if not torch.is_tensor(out) else out

    def forward_text_model(self, encoded_input):
        hidden = self._text_hidden(encoded_input)
        mask = encoded_input["attention_mask"]
        if hidden.is_cuda and hidden.shape[-1] == 768 and not (torch.is_grad_enabled() and hidden.requires_grad):
            # one fused launch for the whole padded batch (rz_text_pool); autograd keeps the torch formula
            from . import ops
            feats, _ = ops.text_pool(hidden, mask, want_q16=False)
            feats = feats.to(hidden.dtype)
        else:
            feats = mean_pooling(hidden, mask)
        return {"text_features_wo_l2_norm": feats, "text_features": F.normalize(feats, p=2, dim=1)}

    # ------------------------------------------------------------------ training forward
    def forward(self, pixel_values, encoded_findings=None, encoded_key_phrases=None,
                encoded_negative_phrases=None, encoded_random_key_phrases=None, return_loss=True,
                **kwargs):
        outputs = {}
        outputs.update(self.forward_vision_model(pixel_values))
        if return_loss:
            loss = 0
            losses = {}
            for loss_type, loss_fn in self.loss_fns.items():
                lo = loss_fn(encoded_key_phrases, outputs["vision_tokens"], self.forward_text_model)
                rl = dict(lo["losses"])
                losses["radzero_loss"] = rl.pop("loss")
                losses.update(rl)
                loss = loss + losses["radzero_loss"] * self.loss_ratio[loss_type]
            losses["loss"] = loss
            outputs["losses"] = losses
        return outputs

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def compute_logits(self, pixel_values, encoded_key_phrases, **kwargs):
        """radzero branch of modeling.py:278-328.

        The N prompts of ``encoded_key_phrases[0]`` are encoded in ONE text-model call (the
        reference makes N calls of batch 1, :290-298; rows are independent so the features
        are the same) and the similarity runs as one fused kernel.
        """
        if self.compute_logits_type != "radzero":
            raise NotImplementedError(self.compute_logits_type)
        vision = self.forward_vision_model(pixel_values)
        loss_fn: RadZeroLoss = self.loss_fns["RadZeroLoss"]
        enc = encoded_key_phrases[0]
        hidden = self._text_hidden(enc)
        q16 = None
        if hidden.is_cuda and hidden.shape[-1] == loss_fn.hidden_dim and loss_fn.sim_op in ("cos", "dot"):
            # mean pooling + the loss's LayerNorm + L2 of ALL prompts in one launch
            from . import ops
            g, b = loss_fn._ln()
            text, q16 = ops.text_pool(hidden, enc["attention_mask"], g, b, l2=loss_fn.sim_op == "cos")
        else:
            text = mean_pooling(hidden, enc["attention_mask"])
            if text.shape[-1] == 2 * loss_fn.hidden_dim:
                text = text[:, loss_fn.hidden_dim:]
        logits, scores, z = loss_fn.similarity(text, vision["vision_tokens"], want_scores=True, q16=q16)
        scores_with_cls = None  # the reference also returns the pre-drop tensor; not materialised here
        return {"logits": logits, "similarity_scores": scores, "t2i_logits": z,
                "t2i_attn_weights": scores_with_cls}

    @torch.no_grad()
    def similarity(self, pixel_values, encoded_text):
        """The hub model's forward: ``(similarity_prob (B, N), similarity_map (B, N, P, P))``.

        README.md:104-111 shows the pair being returned; the hub-side code is not in the
        reference checkout: similarity_prob = sigmoid(logits) as the
        reference's own consumers do (segmentation_utils.py:225), the map is the patch-grid
        scores at the cos / tau scale.
        """
        out = self.compute_logits(pixel_values, [encoded_text])
        s = out["similarity_scores"]
        p = int(round(s.shape[-1] ** 0.5))
        return torch.sigmoid(out["logits"]), s.view(s.shape[0], s.shape[1], p, p)


def build_random_init_model(device="cuda", dtype=torch.float32, seed: int = 42,
                            vision_layers: int = 12, text_layers: int = 12) -> CxrAlignModel:
    """The RadZero architecture with random-init weights, built offline from HF classes.

    DINOv2-base @518 (patch 14 -> 37x37 + CLS), AlignTransformer = 2 DINOv2 layers
    (exp/cxr_pt/model/align_transformers.py:23-45), MPNet-base text encoder
    (exp/cxr_pt/configs/radzero.yaml:16-35).  No checkpoints exist offline.
    """
    from transformers import Dinov2Config, Dinov2Model, MPNetConfig, MPNetModel
    from transformers.models.dinov2.modeling_dinov2 import Dinov2Encoder
    torch.manual_seed(seed)
    vcfg = Dinov2Config(hidden_size=768, num_hidden_layers=vision_layers, num_attention_heads=12,
                        image_size=518, patch_size=14)
    vision = Dinov2Model(vcfg)
    acfg = Dinov2Config(hidden_size=768, num_hidden_layers=2, num_attention_heads=12)
    from .align import AlignTransformer
    align = AlignTransformer(Dinov2Encoder(acfg))     # align_transformers.py:23-45, use_layer_norm=False
    tcfg = MPNetConfig(num_hidden_layers=text_layers)
    text = MPNetModel(tcfg)
    model = CxrAlignModel(vision, align, text)
    return model.to(device=device, dtype=dtype).eval()
