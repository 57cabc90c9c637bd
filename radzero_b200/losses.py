"""Host-side mirror of the reference's ``exp/cxr_pt/model/losses.py`` for the VL-CABS path.

Same class / function names, constructor arguments, call signatures, output keys, shapes
(including the ``.squeeze()`` / ``.T`` quirks) and state-dict keys as the reference
(``RadZeroLoss`` losses.py:33-184, ``SimilarityLogit`` :187-240,
``multi_positive_nce_loss`` :243-344), but every tensor operation on the path is one of the
hand-written sm_100a kernels behind the C ABI (``include/rz_b200.h``).  There is no CPU or
eager-PyTorch fallback: CPU tensors raise ``RzError``.

What differs by design:
  * the (B, N, L) probability tensor and the (B, N, 768) expanded query are never formed;
  * under DDP the reference all-gathers the vision tokens of every rank and computes the full
    (N_total x B_global) problem redundantly (losses.py:87-88, 156-161); here each rank keeps
    its own images (columns of the logit matrix), all-gathers only the fp16 normalised
    sentence embeddings, and all-reduces the row sums -- the loss and gradients are the same
    up to summation order.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from ._lib import RzError

__all__ = ["RadZeroLoss", "SimilarityLogit", "multi_positive_nce_loss", "pad_and_gather"]


def _squeeze_quirk(z_nb: torch.Tensor) -> torch.Tensor:
    """``matmul(...).squeeze()`` then ``.T`` on the (B, N) logits (losses.py:229-233).

    z_nb is (N, B).  The reference squeezes (B, N, 1, 1): B == 1 or N == 1 collapse to 1-d,
    both to 0-d; ``.T`` of those is the identity.
    """
    n, b = z_nb.shape
    if n == 1 and b == 1:
        return z_nb.reshape(())
    if b == 1:
        return z_nb.reshape(n)
    if n == 1:
        return z_nb.reshape(b)
    return z_nb


# ------------------------------------------------------------------------------------ a3
class SimilarityLogit(nn.Module):
    """Drop-in for losses.py:187-240 running on the fused tcgen05 kernel.

    ``queries`` (N, D) and ``local_tokens`` (B, L, D) are the (already LayerNorm-ed) inputs
    the reference passes; L2 normalisation (sim_op "cos"), the similarity GEMM, the softmax
    pooling and the pooled logit happen inside ``rz_prep_rows`` + ``rz_sim_fwd``.
    Returns ``(logits (N, B) with the reference's squeeze quirk, [scores (B, N, L)] | None)``.
    """

    def __init__(self, sim_op: str = "dot", **kwargs):
        super().__init__()
        self.sim_op = sim_op

    def forward(self, queries: torch.Tensor, local_tokens: torch.Tensor,
                need_attn_weights: bool = False, repeat: bool = True, **kwargs):
        if not repeat:
            assert queries.dim() == 3
        if self.sim_op == "cos":
            temperature = kwargs.get("temperature")
            assert temperature is not None
            scale = 1.0 / float(temperature)
            l2 = True
        elif self.sim_op == "dot":
            scale = 1.0 / math.sqrt(local_tokens.size(-1))
            l2 = False
        else:
            raise NotImplementedError
        if not repeat:
            # per-image queries (B, N, D), losses.py:204-206: no caller in the reference (radzero.yaml always
            # shares the prompts), so one launch per image of the shared-prompt kernel; inference only
            if torch.is_grad_enabled() and (queries.requires_grad or local_tokens.requires_grad):
                raise NotImplementedError("repeat=False (per-image queries) has no backward on the VL-CABS path")
            if queries.shape[0] != local_tokens.shape[0]:
                raise RzError("repeat=False needs one (N, D) query set per image")
            per = [_similarity_forward(queries[b], local_tokens[b:b + 1], None, None, scale, l2,
                                       need_attn_weights, drop_cls=False) for b in range(queries.shape[0])]
            z = torch.cat([zb for zb, _ in per], dim=1)
            scores = torch.cat([sb for _, sb in per], dim=0) if need_attn_weights else None
        elif torch.is_grad_enabled() and (queries.requires_grad or local_tokens.requires_grad):
            from .training import similarity_logit_autograd
            z, scores = similarity_logit_autograd(queries, local_tokens, scale, l2, need_attn_weights)
        else:
            z, scores = _similarity_forward(queries, local_tokens, None, None, scale, l2,
                                            need_attn_weights, drop_cls=False)
        return _squeeze_quirk(z), ([scores] if need_attn_weights else None)


def _similarity_forward(text: torch.Tensor, tokens: torch.Tensor, gamma, beta, scale: float,
                        l2: bool, want_scores: bool, drop_cls: bool, q16: Optional[torch.Tensor] = None, **zkw):
    """The inference forward.  N <= 16 prompts: prep(text) + ONE kernel that reads the raw
    tokens (LayerNorm + L2 fused into the GEMM's loader warps).  Larger prompt sets:
    prep of both operands to fp16, then the TMA-fed fused forward.  Returns (Z, scores)."""
    if tokens.dim() != 3:
        raise RzError("vision tokens must be (B, L, 768)")
    B, L, _ = tokens.shape
    if (q16 is None and l2 and ops.USE_FUSED_PREP and text.dim() == 2 and text.shape[0] <= ops.FUSED_PREP_MAX_TEXT
            and text.dtype == torch.float32):
        # few prompts, cosine: their LayerNorm + L2 also happen inside the one kernel (no prep launch)
        out = ops.sim_fwd_tokens(tokens, gamma, beta, None, scale, l2=True, text_raw=text, want_scores=want_scores,
                                 drop_cls=drop_cls, **zkw)
        return out["z"], out["scores"]
    if q16 is None:      # else: rows already normalised by ops.text_pool (fused with the mean pooling)
        q16, _, _ = ops.prep_rows(text, gamma, beta, l2=l2)
    qin = None
    if not l2:
        qin = 1.0 / q16.float().norm(dim=-1).clamp_min(1e-12)   # F.normalize(query), losses.py:226
    if ops.USE_FUSED_PREP and q16.shape[0] <= ops.FUSED_PREP_MAX_TEXT:
        out = ops.sim_fwd_tokens(tokens, gamma, beta, q16, scale, l2=l2, want_scores=want_scores,
                                 drop_cls=drop_cls, q_inv_norm=qin, **zkw)
    else:
        Lp = ops.padded_tokens(L)
        k16, _, _ = ops.prep_rows(tokens, gamma, beta, rows_per_group=L, rows_per_group_padded=Lp, l2=l2)
        out = ops.sim_fwd(k16.view(B, Lp, ops.HIDDEN), q16, L, scale, want_scores=want_scores,
                          drop_cls=drop_cls, q_inv_norm=qin, **zkw)
    return out["z"], out["scores"]


# ------------------------------------------------------------------------------------ a5
class _MpNce(torch.autograd.Function):
    """Loss value + closed-form dL/dZ and dL/dtemperature from the two MP-NCE kernels."""

    @staticmethod
    def forward(ctx, logits, group_map, temperature, eps, row_sum, col_sum):
        z = logits.detach()
        z = z if (z.dtype == torch.float32 and z.stride(-1) == 1) else z.float().contiguous()
        n, b = z.shape
        on_device = isinstance(temperature, torch.Tensor) and temperature.is_cuda
        if on_device:
            # tau stays on the device (the kernels read log tau through a pointer): no host sync
            tau = temperature.detach().float().reshape(1)
            kw = dict(log_tau=tau.log())
            inv_tau = 1.0
        else:
            tau = float(temperature)
            kw = {}
            inv_tau = 1.0 / tau
        rs, ps, cn, cp = ops.mpnce_partials(z, group_map, 0, inv_tau, eps=eps, col_sum=col_sum, b_global=b, **kw)
        terms, dz = ops.mpnce_finish(z, group_map, 0, b, inv_tau, rs, ps, cn, cp, eps=eps,
                                     row_sum=row_sum, col_sum=col_sum, want_dz=True, **kw)
        loss = terms[3].clone()
        ctx.save_for_backward(dz, terms, tau if on_device else None)
        ctx.tau = None if on_device else tau
        ctx.temp_is_tensor = isinstance(temperature, torch.Tensor)
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        dz, terms, tau_t = ctx.saved_tensors
        gz = (dz * g).to(ctx.in_dtype)
        gt = None
        if ctx.temp_is_tensor:
            tau = tau_t if tau_t is not None else ctx.tau
            gt = (-(terms[2] / tau) * g).reshape(())
        return gz, None, gt, None, None, None


def multi_positive_nce_loss(logits: torch.Tensor, group_map: torch.Tensor, temperature=1.0,
                            eps: float = 1e-8, row_sum: bool = False, col_sum: bool = False):
    """Drop-in for losses.py:243-293 (+ get_row_loss :296-320, get_col_loss :323-344).

    ``logits`` (N_total, B_global) fp32 on CUDA, ``group_map`` (N_total,) int64.  Two fused
    kernels produce the loss and (for autograd) dL/dlogits and dL/dtemperature.
    """
    if logits.dim() != 2:
        raise RzError("logits must be (N_total, B_global)")
    t = temperature
    if isinstance(t, torch.Tensor) and t.numel() == 1 and t.dim() > 0:
        out = _MpNce.apply(logits, group_map, t.reshape(()), eps, row_sum, col_sum)
    else:
        out = _MpNce.apply(logits, group_map, t, eps, row_sum, col_sum)
    return out


def _index_tensor(values, device) -> torch.Tensor:
    """int64 index tensor on ``device``.  CUDA: staged in pinned memory and copied asynchronously (a
    pageable host-to-device copy would synchronise the stream, i.e. drain the GPU queue)."""
    t = values.to(torch.int64) if torch.is_tensor(values) else torch.tensor(values, dtype=torch.int64)
    if torch.device(device).type != "cuda":
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


# ------------------------------------------------------------------------------------ a8
def pad_and_gather(tensor: torch.Tensor, group=None) -> torch.Tensor:
    """Ragged all-gather along dim 0 (losses.py:386-409), without autograd.

    Sizes are exchanged first, rows padded to the maximum, gathered in ONE collective into a
    single buffer and trimmed.  Keeps the input dtype (the reference pads into fp32).
    """
    world = dist.get_world_size(group)
    n_local = torch.tensor([tensor.shape[0]], device=tensor.device, dtype=torch.int64)
    sizes = torch.empty(world, device=tensor.device, dtype=torch.int64)
    dist.all_gather_into_tensor(sizes, n_local, group=group)
    sizes_l = sizes.tolist()
    n_max = max(sizes_l)
    padded = tensor.new_zeros((n_max,) + tuple(tensor.shape[1:]))
    padded[: tensor.shape[0]] = tensor
    out = tensor.new_empty((world * n_max,) + tuple(tensor.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view((world, n_max) + tuple(tensor.shape[1:]))
    return torch.cat([out[r, : sizes_l[r]] for r in range(world)], dim=0)


# ------------------------------------------------------------------------------------ a1, a2, a4
class RadZeroLoss(nn.Module):
    """Drop-in for losses.py:33-184.

    Parameters (state-dict keys preserved): ``layer_norm.weight``, ``layer_norm.bias``,
    ``loss_temperature`` (log tau) and, when given, ``attn_temperature``.
    """

    def __init__(self, hidden_dim=768, use_vision_cls_token=True, attn_temperature=None,
                 loss_temperature=0.07, text_features_l2_norm=False, mpnce_row_sum=False,
                 mpnce_col_sum=False, sim_op="dot", use_layer_norm=True, **kwargs):
        super().__init__()
        if hidden_dim != ops.HIDDEN:
            raise RzError(f"the B200 path is built for hidden_dim={ops.HIDDEN} (radzero.yaml:40)")
        self.hidden_dim = hidden_dim
        self.layer_norm = nn.LayerNorm(hidden_dim) if use_layer_norm else None
        self.use_vision_cls_token = use_vision_cls_token
        self.loss_temperature = nn.Parameter(torch.FloatTensor([np.log(loss_temperature)]))
        if attn_temperature is not None:
            self.attn_temperature = nn.Parameter(torch.FloatTensor([np.log(attn_temperature)]))
        else:
            self.attn_temperature = None
        self.text_features_l2_norm = text_features_l2_norm
        self.sim_op = sim_op
        self.similarity_logit = SimilarityLogit(sim_op)
        self.mpnce_row_sum = mpnce_row_sum
        self.mpnce_col_sum = mpnce_col_sum
        # the attribute CxrAlignModel.compute_logits reads (modeling.py:320) but the
        # reference's __init__ never sets; False = released behaviour
        self.compute_i2t_loss = False
        # one text-model call for all sentences of the local batch instead of one per image
        self.batch_text_calls = bool(kwargs.get("batch_text_calls", True))
        self.text_pad_token_id = int(kwargs.get("text_pad_token_id", 1))      # MPNet / RoBERTa <pad>

    # -- helpers -------------------------------------------------------------------------
    def _ln(self):
        if self.layer_norm is None:
            return None, None
        return self.layer_norm.weight, self.layer_norm.bias

    def _attn_log_tau(self):
        return self.attn_temperature if self.attn_temperature is not None else self.loss_temperature

    def _scale(self) -> float:
        if self.sim_op == "cos":
            return 1.0 / float(self._attn_log_tau().detach().exp())
        if self.sim_op == "dot":
            return 1.0 / math.sqrt(self.hidden_dim)
        raise NotImplementedError

    @staticmethod
    def _merge_encodings(key_phrases, pad_token_id: int = 1):
        """The per-image tokenised sentence batches (dataset.py:172-181: ``input_ids (n_i, T_i)``,
        ``attention_mask (n_i, T_i)``) as ONE padded batch, or None when they are not plain tensors.
        Padding positions carry ``attention_mask`` 0, so they take no part in attention or pooling."""
        try:
            ids = [kp["input_ids"] for kp in key_phrases]
            ams = [kp["attention_mask"] for kp in key_phrases]
        except (KeyError, TypeError, IndexError):
            return None
        if not ids:
            return None
        for kp, a, b in zip(key_phrases, ids, ams):
            if not (torch.is_tensor(a) and torch.is_tensor(b) and a.dim() == 2 and b.dim() == 2):
                return None
            if len(kp) != 2:
                return None                           # token_type_ids etc.: keep the reference's call pattern
        counts = [int(t.shape[0]) for t in ids]
        t_max = max(t.shape[1] for t in ids)
        if any(t.shape[1] != t_max for t in ids):
            pad = torch.nn.functional.pad
            ids = [t if t.shape[1] == t_max else pad(t, (0, t_max - t.shape[1]), value=pad_token_id) for t in ids]
            ams = [t if t.shape[1] == t_max else pad(t, (0, t_max - t.shape[1]), value=0) for t in ams]
        return {"input_ids": torch.cat(ids, dim=0), "attention_mask": torch.cat(ams, dim=0)}, counts

    def collect_text_features(self, key_phrases, forward_text_model, rank: int = 0,
                              want_group_map: bool = True):
        """Raw (pre-LayerNorm) sentence embeddings + group_map, losses.py:126-153.

        The reference calls the text model once per image (B_local sequential calls, :135-151).  With
        ``batch_text_calls`` (default) the per-image token batches are merged into one padded batch and
        the text model runs ONCE; rows are independent, so the features
        are the same.  Inputs that are not plain ``input_ids`` / ``attention_mask`` tensors fall back
        to the reference's call pattern.  ``want_group_map=False`` (inference: nothing reads it) skips
        building the index tensor, which is a synchronising host-to-device copy."""
        b_local = len(key_phrases)
        merged = self._merge_encodings(key_phrases, self.text_pad_token_id) if self.batch_text_calls else None
        if merged is not None:
            enc, counts = merged
            f = forward_text_model(enc)
            feat = f["text_features"] if self.text_features_l2_norm else f["text_features_wo_l2_norm"]
            if feat.shape[-1] == 2 * self.hidden_dim:
                feat = feat[:, self.hidden_dim:]
            if not want_group_map:
                return feat, None
            if feat.is_cuda and max(counts, default=0) <= 65535:
                # written on the device from launch parameters: no copy-engine traffic on the compute stream
                return feat, ops.group_map_from_counts(counts, rank * b_local, feat.device)
            group = torch.from_numpy(np.repeat(np.arange(rank * b_local, (rank + 1) * b_local, dtype=np.int64),
                                               np.asarray(counts, dtype=np.int64)))
            return feat, _index_tensor(group, feat.device)
        feats: List[torch.Tensor] = []
        group: List[int] = []
        for i, kp in enumerate(key_phrases):
            f = forward_text_model(kp)
            feat = f["text_features"] if self.text_features_l2_norm else f["text_features_wo_l2_norm"]
            if feat.shape[-1] == 2 * self.hidden_dim:
                feat = feat[:, self.hidden_dim:]
            feats.append(feat)
            group.extend([i + rank * b_local] * feat.size(0))
        text = torch.cat(feats, dim=0)
        if not want_group_map:
            return text, None
        counts = [int(f.size(0)) for f in feats]
        if text.is_cuda and max(counts, default=0) <= 65535:
            # written on the device from launch parameters: no copy-engine traffic on the compute stream
            return text, ops.group_map_from_counts(counts, rank * b_local, text.device)
        return text, _index_tensor(group, text.device)

    def compute_text_features(self, key_phrases, forward_text_model, ddp_gather=True):
        """losses.py:126-166 (kept for API parity; forward() uses the fused path instead)."""
        rank = dist.get_rank() if (ddp_gather and dist.is_initialized()) else 0
        text, group_map = self.collect_text_features(key_phrases, forward_text_model, rank)
        if ddp_gather and dist.is_initialized():
            text = pad_and_gather(text)
            group_map = pad_and_gather(group_map).long()
        if self.layer_norm is not None:
            g, b = self._ln()
            _, text, _ = ops.prep_rows(text, g.detach(), b.detach(), want_f16=False, want_f32=True, l2=False)
        return text, group_map

    # -- forward -------------------------------------------------------------------------
    def forward(self, key_phrases, vision_tokens, forward_text_model, ddp_gather=True,
                need_attn_weights=False, compute_loss=True, **kwargs):
        outputs: Dict = {}
        distributed = bool(ddp_gather and dist.is_initialized() and dist.get_world_size() > 1)
        rank = dist.get_rank() if distributed else 0
        text, group_map = self.collect_text_features(key_phrases, forward_text_model, rank,
                                                     want_group_map=bool(compute_loss or distributed))
        tokens = vision_tokens if self.use_vision_cls_token else vision_tokens[:, 1:]
        gamma, beta = self._ln()
        grad_on = torch.is_grad_enabled()
        if grad_on and not compute_loss and (text.requires_grad or vision_tokens.requires_grad):
            # the reference would hand back differentiable logits here; this build's inference kernels
            # keep no autograd state, so refuse instead of silently detaching
            raise RzError("RadZeroLoss.forward(compute_loss=False) on inputs that require grad: wrap the "
                          "call in torch.no_grad(), or use compute_loss=True for the training node")
        wants_grad = grad_on and compute_loss and (
            text.requires_grad or vision_tokens.requires_grad or self.loss_temperature.requires_grad
            or (gamma is not None and gamma.requires_grad))
        if wants_grad or distributed:
            from .training import contrastive_step
            res = contrastive_step(self, text, group_map, tokens, distributed=distributed,
                                   need_attn_weights=need_attn_weights, compute_loss=compute_loss,
                                   gather_logits=bool(kwargs.get("gather_logits", False)))
            outputs["t2i_logits"] = _squeeze_quirk(res["z"])
            outputs["t2i_attn_weights"] = [res["scores"]] if need_attn_weights else None
            if compute_loss:
                outputs["losses"] = {"t2i_loss": res["loss"], "loss": res["loss"]}
            return outputs
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        # the attention temperature is read from the parameter ON THE DEVICE (no host synchronisation)
        l2 = self.sim_op == "cos"
        zkw = {"log_tau_scale": self._attn_log_tau()} if l2 else {}
        scale = 1.0 if l2 else 1.0 / math.sqrt(self.hidden_dim)
        z, scores = _similarity_forward(text.detach(), tokens.detach(), g, b, scale, l2, need_attn_weights,
                                        drop_cls=False, **zkw)
        outputs["t2i_logits"] = _squeeze_quirk(z)
        outputs["t2i_attn_weights"] = [scores] if need_attn_weights else None
        if compute_loss:
            loss = multi_positive_nce_loss(z, group_map, temperature=self.loss_temperature.detach().exp(),
                                           row_sum=self.mpnce_row_sum, col_sum=self.mpnce_col_sum)
            outputs["losses"] = {"t2i_loss": loss, "loss": loss}
        return outputs

    def compute_t2i_logits(self, text_features, vision_attn_tokens, need_attn_weights, repeat=True):
        """losses.py:168-184: inputs already LayerNorm-ed."""
        return self.similarity_logit(text_features, vision_attn_tokens, need_attn_weights,
                                     repeat=repeat, temperature=self._attn_log_tau().exp())

    # -- inference fast path ---------------------------------------------------------------
    @torch.no_grad()
    def similarity_prob(self, text_features: torch.Tensor, vision_tokens: torch.Tensor) -> torch.Tensor:
        """similarity_prob (B, N) = sigmoid(Z^T / tau) only (zero-shot classification).  The
        temperatures are read from the parameters ON THE DEVICE and the ``/ tau`` + sigmoid run
        in the kernel epilogue: no host synchronisation; for N <= 16 fp32 prompts ONE fused kernel (the
        prompts' LayerNorm + L2 run in its prologue) plus its merge launch."""
        tokens = vision_tokens if self.use_vision_cls_token else vision_tokens[:, 1:]
        gamma, beta = self._ln()
        l2 = self.sim_op == "cos"
        zkw = dict(z_sigmoid=True, z_image_major=True, log_tau_z=self.loss_temperature)
        if l2:
            zkw["log_tau_scale"] = self._attn_log_tau()
        scale = 1.0 if l2 else 1.0 / math.sqrt(self.hidden_dim)
        z, _ = _similarity_forward(text_features, tokens, gamma, beta, scale, l2, False, False, **zkw)
        return z

    @torch.no_grad()
    def similarity(self, text_features: torch.Tensor, vision_tokens: torch.Tensor, *,
                   want_scores: bool = True, drop_cls: Optional[bool] = None,
                   q16: Optional[torch.Tensor] = None):
        """Everything ``compute_logits`` needs in one fused pass.

        Returns ``(logits (B, N) = Z^T / tau, similarity_scores (B, N, L - drop) | None,
        t2i_logits (N, B))`` -- modeling.py:300-328 without the intermediate (B, N, L+1)
        tensor, the CLS-dropping copy or the stack/mean over a one-element list.  ``q16``: the
        prompts' rows already LayerNorm-ed + normalised by ``ops.text_pool`` (skips their prep launch).
        """
        tokens = vision_tokens if self.use_vision_cls_token else vision_tokens[:, 1:]
        if drop_cls is None:
            drop_cls = bool(self.use_vision_cls_token)
        gamma, beta = self._ln()
        l2 = self.sim_op == "cos"
        zkw = {"log_tau_scale": self._attn_log_tau()} if l2 else {}
        scale = 1.0 if l2 else 1.0 / math.sqrt(self.hidden_dim)
        z, scores = _similarity_forward(text_features, tokens, gamma, beta, scale, l2, want_scores,
                                        drop_cls, q16=q16, **zkw)
        logits = z.T / self.loss_temperature.exp()
        return logits, scores, z
