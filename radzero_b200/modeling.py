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
from transformers import AutoConfig, AutoModel, PretrainedConfig, PreTrainedModel

from .losses import RadZeroLoss

RADZERO_LOSS_CFG = dict(  # exp/cxr_pt/configs/radzero.yaml:37-46
    hidden_dim=768, use_vision_cls_token=True, attn_temperature=None, loss_temperature=0.07,
    text_features_l2_norm=False, mpnce_row_sum=False, mpnce_col_sum=False, sim_op="cos",
    use_layer_norm=True)


def mean_pooling(last_hidden_state: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Masked mean over tokens (exp/cxr_pt/model/text_encoders.py:32-41)."""
    m = attention_mask.unsqueeze(-1).to(last_hidden_state.dtype)
    return (last_hidden_state * m).sum(1) / m.sum(1).clamp(min=1e-9)


# exp/cxr_pt/configs/radzero.yaml:16-35 restated as explicit sub-configs.  The reference resolves the
# vision / text sub-configs from the hub at construction time (configuration.py:25-27: network); here every
# field needed to BUILD the architecture offline is spelled out.
VISION_CONFIG = dict(model_type="dinov2", hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                     image_size=518, patch_size=14, img_size=518)
TEXT_CONFIG = dict(model_type="mpnet", use_text_projection=False, use_cls_token=False, num_hidden_layers=12)
ALIGN_TRANSFORMER_CONFIG = dict(model_type="align_transformer", hidden_size=768, num_hidden_layers=2,
                                num_attention_heads=12, projector_config=None, use_layer_norm=False)


class CxrAlignConfig(PretrainedConfig):
    """Mirror of exp/cxr_pt/model/configuration.py:108-131 (``CxrAlignConfig``: vision_config, text_config,
    align_transformer_config + ``loss`` / ``compute_logits_type`` kwargs), serialisable to config.json so that
    ``save_pretrained`` / ``AutoModel.from_pretrained`` round-trip.  Sub-configs are plain dicts."""

    model_type = "radzero"

    def __init__(self, vision_config=None, text_config=None, align_transformer_config=None, loss=None,
                 compute_logits_type="radzero", **kwargs):
        super().__init__(**kwargs)
        self.vision_config = dict(VISION_CONFIG, **(vision_config or {}))
        self.text_config = dict(TEXT_CONFIG, **(text_config or {}))
        self.align_transformer_config = dict(ALIGN_TRANSFORMER_CONFIG, **(align_transformer_config or {}))
        self.loss = loss or {"apply": ["RadZeroLoss"], "ratio": [1.0], "RadZeroLoss": dict(RADZERO_LOSS_CFG)}
        self.compute_logits_type = compute_logits_type


def _build_vision(cfg: Dict) -> nn.Module:
    from transformers import Dinov2Config, Dinov2Model
    if cfg.get("model_type", "dinov2") != "dinov2":
        raise NotImplementedError(cfg.get("model_type"))          # modeling.py:107-108
    keys = ("hidden_size", "num_hidden_layers", "num_attention_heads", "image_size", "patch_size")
    return Dinov2Model(Dinov2Config(**{k: cfg[k] for k in keys if k in cfg}))


def _build_align(cfg: Dict) -> nn.Module:
    from transformers import Dinov2Config
    from transformers.models.dinov2.modeling_dinov2 import Dinov2Encoder
    from .align import AlignTransformer
    if cfg.get("model_type", "align_transformer") != "align_transformer":
        raise NotImplementedError(cfg.get("model_type"))          # align_transformers.py:9-21
    enc = Dinov2Encoder(Dinov2Config(hidden_size=cfg["hidden_size"], num_hidden_layers=cfg["num_hidden_layers"],
                                     num_attention_heads=cfg.get("num_attention_heads", 12)))
    return AlignTransformer(enc, nn.LayerNorm(cfg["hidden_size"]) if cfg.get("use_layer_norm", False) else None)


def _build_text(cfg: Dict) -> nn.Module:
    from transformers import MPNetConfig, MPNetModel
    if cfg.get("model_type", "mpnet") != "mpnet":
        raise NotImplementedError(cfg.get("model_type"))          # modeling.py:205-206
    return MPNetModel(MPNetConfig(num_hidden_layers=cfg.get("num_hidden_layers", 12)))


class CxrAlignModel(PreTrainedModel):
    """``CxrAlignModel(PreTrainedModel)`` with ``config_class = CxrAlignConfig`` (modeling.py:23-25), registered
    with ``AutoConfig`` / ``AutoModel`` below so that ``AutoModel.from_pretrained(path)`` (README.md:77-82)
    resolves to this class once ``radzero_b200`` is imported.  State-dict keys follow the reference:
    ``vision_model.*``, ``text_model.*``, ``align_transformer.transformer_layers.*``,
    ``loss_fns.RadZeroLoss.{loss_temperature, layer_norm.weight, layer_norm.bias}``."""

    config_class = CxrAlignConfig
    base_model_prefix = "cxr_align"
    main_input_name = "pixel_values"
    supports_gradient_checkpointing = False
    _no_split_modules = []

    def __init__(self, config: Optional[CxrAlignConfig] = None, vision_model: Optional[nn.Module] = None,
                 align_transformer: Optional[nn.Module] = None, text_model: Optional[nn.Module] = None,
                 loss_cfg: Optional[Dict] = None, compute_logits_type: Optional[str] = None):
        if isinstance(config, nn.Module):      # round-1 call style: CxrAlignModel(vision, align, text[, loss_cfg])
            config, vision_model, align_transformer, text_model, loss_cfg = (
                None, config, vision_model, align_transformer, text_model if isinstance(text_model, dict) else loss_cfg)
        config = config or CxrAlignConfig()
        super().__init__(config)
        given = vision_model is not None or align_transformer is not None or text_model is not None
        self.vision_model = vision_model if vision_model is not None else _build_vision(config.vision_config)
        self.text_model = text_model if text_model is not None else _build_text(config.text_config)
        self.hidden_size = int(config.align_transformer_config["hidden_size"])
        self.text_projector = None               # use_text_projection: False (radzero.yaml:23)
        self.align_transformer = (align_transformer if align_transformer is not None
                                  else _build_align(config.align_transformer_config))
        self.loss_ratio = {}
        self.loss_fns = nn.ModuleDict()
        lc = config.loss
        for loss_type, ratio in zip(lc["apply"], lc["ratio"]):
            if loss_type != "RadZeroLoss":
                raise NotImplementedError(f"{loss_type}: only the VL-CABS loss is on this path (DESIGN.md section 8)")
            self.loss_fns[loss_type] = RadZeroLoss(**(loss_cfg or lc.get(loss_type) or RADZERO_LOSS_CFG))
            self.loss_ratio[loss_type] = ratio
        self.compute_logits_type = compute_logits_type or config.compute_logits_type
        if not given:
            self.post_init()

    def _init_weights(self, module):
        """exp/cxr_pt/model/common_layers.py:13-28.  Parameters already filled by ``from_pretrained`` carry
        transformers' ``_is_hf_initialized`` mark and are left alone."""
        fresh = lambda p: p is not None and not getattr(p, "_is_hf_initialized", False)
        with torch.no_grad():
            if isinstance(module, (nn.Conv2d, nn.Embedding, nn.Linear)):
                if fresh(module.weight):
                    module.weight.normal_(mean=0.0, std=0.02)
                if fresh(getattr(module, "bias", None)):
                    module.bias.zero_()
            elif isinstance(module, nn.LayerNorm):
                if fresh(module.bias):
                    module.bias.zero_()
                if fresh(module.weight):
                    module.weight.fill_(1.0)

    # ------------------------------------------------------------------ encoders (stock HF)
    def forward_vision_model(self, pixel_values, handoff_f16: bool = False):
        """modeling.py:96-123.  ``handoff_f16`` (used by ``compute_logits``, whose only consumer of the tokens
        is the similarity kernel): the AlignTransformer emits its last layer in fp16."""
        out = self.vision_model(pixel_values)
        vision_tokens = out["last_hidden_state"] if not torch.is_tensor(out) else out
        if handoff_f16 and hasattr(self.align_transformer, "_forward_kernels") and vision_tokens.is_cuda:
            at = self.align_transformer(vision_tokens, handoff_f16=True)
        else:
            at = self.align_transformer(vision_tokens)
        vision_tokens = at["last_hidden_state"] if not torch.is_tensor(at) else at
        if handoff_f16:          # compute_logits reads nothing but the tokens (image_features are unused there)
            return {"vision_tokens": vision_tokens}
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
                input_ids=None, attention_mask=None, **kwargs):
        """modeling.py:213-276 (training forward).  Called hub-style -- ``model(pixel_values=...,
        input_ids=..., attention_mask=...)`` with no key phrases -- it is the released model's inference
        forward and returns ``(similarity_prob, similarity_map)`` (README.md:104-111)."""
        if input_ids is not None and encoded_key_phrases is None:
            if attention_mask is None:
                attention_mask = torch.ones_like(input_ids)
            return self.similarity(pixel_values, {"input_ids": input_ids, "attention_mask": attention_mask})
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
        vision = self.forward_vision_model(pixel_values, handoff_f16=bool(kwargs.get("handoff_f16", True)))
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
        # the scores are produced WITH the CLS column, exactly the tensor the reference passes through as
        # ``t2i_attn_weights`` (modeling.py:300-308); ``similarity_scores`` is its ``[:, :, 1:]`` view
        # (modeling.py:316-317 slices the same way), so the passthrough costs nothing
        keep_cls = bool(kwargs.get("return_attn_weights", True)) and loss_fn.use_vision_cls_token
        logits, scores, z = loss_fn.similarity(text, vision["vision_tokens"], want_scores=True, q16=q16,
                                               drop_cls=not keep_cls)
        attn = None
        if keep_cls:
            attn = [scores]
            scores = scores[:, :, 1:]
        elif not loss_fn.use_vision_cls_token:
            attn = [scores]
        return {"logits": logits, "similarity_scores": scores, "t2i_logits": z, "t2i_attn_weights": attn}

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
    torch.manual_seed(seed)
    cfg = CxrAlignConfig(vision_config={"num_hidden_layers": vision_layers},
                         text_config={"num_hidden_layers": text_layers})
    model = CxrAlignModel(cfg)
    return model.to(device=device, dtype=dtype).eval()


AutoConfig.register(CxrAlignConfig.model_type, CxrAlignConfig)
AutoModel.register(CxrAlignConfig, CxrAlignModel)
