"""CPU oracle: a restatement of RadZero's VL-CABS similarity path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Written from the formulas, in
plain torch tensor algebra so that it runs in fp32 (the reference's inference dtype) or
fp64 (tight checks), on CPU -- or on a CUDA device when a GPU test wants a full-size
checker.  Every function cites the reference lines it restates
(paths relative to the reference checkout).

Notation: B images, N sentences/prompts, L = 1370 tokens (CLS + 37*37),
D = 768, tau = exp(log_tau).

    q  = L2(LN(T))            T (N, D) sentence embeddings
    k  = L2(LN(X))            X (B, L, D) vision tokens
    S  = q k^T / tau          (B, N, L)      "t2i_attn_weights" (pre-softmax!)
    P  = softmax_L(S)
    o  = P k                  (B, N, D)
    Z  = <q, o / |o|>         (N, B)         "t2i_logits"
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

__all__ = [
    "LN_EPS",
    "L2_EPS",
    "NCE_EPS",
    "masked_mean_pool",
    "layer_norm_rows",
    "l2_normalize_rows",
    "similarity_logit",
    "multi_positive_nce_loss",
    "build_group_map",
    "radzero_forward",
    "compute_logits_glue",
    "similarity_prob_from_logits",
    "bilinear_upsample",
    "interpolate_similarity_scores",
    "grounding_point",
    "zero_shot_labels",
    "dice_sweep_stats",
    "dice_score_samplewise",
    "compute_specificity",
    "contrastive_step_reference",
]

LN_EPS = 1e-5  # nn.LayerNorm default, exp/cxr_pt/model/losses.py:51
L2_EPS = 1e-12  # F.normalize default, losses.py:212-213
NCE_EPS = 1e-8  # multi_positive_nce_loss default, losses.py:247


# --------------------------------------------------------------------------- a9 / T0
def masked_mean_pool(token_embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Sentence embedding of the MPNet branch -- modeling.py:147-156 (``use_cls_token: False``,
    radzero.yaml:26): sum of the token embeddings weighted by the float attention mask over
    ``clamp(mask.sum, min=1e-9)``.  (n, T, D), (n, T) -> (n, D) = ``text_features_wo_l2_norm``."""
    m = attention_mask.unsqueeze(-1).expand(token_embeddings.size()).float().to(token_embeddings.dtype)
    return torch.sum(token_embeddings * m, 1) / torch.clamp(m.sum(1), min=1e-9)


# --------------------------------------------------------------------------- a1 / K1
def layer_norm_rows(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                    eps: float = LN_EPS) -> torch.Tensor:
    """Row LayerNorm with biased variance -- losses.py:90-91 (tokens), :163-164 (text).

    The reference applies the SAME nn.LayerNorm(768) module to both (losses.py:51), so
    gamma/beta are shared.
    """
    # the same ATen op nn.LayerNorm calls (biased variance); also what makes the oracle a fair
    # CPU baseline -- a hand-rolled mean/var version is ~20x slower on the host
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), gamma.to(x.dtype), beta.to(x.dtype), eps)


# --------------------------------------------------------------------------- K2
def l2_normalize_rows(x: torch.Tensor, eps: float = L2_EPS) -> torch.Tensor:
    """x / max(|x|_2, eps) -- F.normalize(p=2, dim=-1), losses.py:212-213, 226-227."""
    return torch.nn.functional.normalize(x, p=2, dim=-1, eps=eps)


# --------------------------------------------------------------------------- a3
def similarity_logit(queries: torch.Tensor, local_tokens: torch.Tensor, *,
                     temperature: Optional[torch.Tensor | float] = None,
                     sim_op: str = "cos", need_scores: bool = False,
                     squeeze_quirk: bool = True
                     ) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]]]:
    """SimilarityLogit.forward, losses.py:192-240.

    queries (N, D) [already LayerNorm-ed] -- or (B, N, D), one set per image: the reference's
    ``repeat=False`` (losses.py:204-206) --, local_tokens (B, L, D) [already LayerNorm-ed].
    Returns (Z, [S] or None) with Z of shape (N, B).  With ``squeeze_quirk`` the
    reference's ``.squeeze()`` then ``.T`` (losses.py:229-233) is reproduced: B == 1 or
    N == 1 collapses Z to 1-d, B == N == 1 to 0-d.
    """
    B = local_tokens.shape[0]
    if sim_op == "cos":
        if temperature is None:
            raise AssertionError("cos similarity needs a temperature")  # losses.py:210
        q = l2_normalize_rows(queries)           # losses.py:212 (on the expanded copy)
        k = l2_normalize_rows(local_tokens)      # losses.py:213
        denom = temperature
    elif sim_op == "dot":
        q = queries
        k = local_tokens
        denom = math.sqrt(local_tokens.shape[-1])  # losses.py:215
    else:
        raise NotImplementedError(sim_op)          # losses.py:216-217

    per_image = q.dim() == 3                                     # repeat=False, losses.py:204-206
    scores = torch.einsum("bnd,bld->bnl" if per_image else "nd,bld->bnl", q, k) / denom   # losses.py:219-221
    probs = torch.softmax(scores, dim=-1)                        # losses.py:222
    pooled = torch.einsum("bnl,bld->bnd", probs, k)              # losses.py:224
    qn = l2_normalize_rows(q)                                    # losses.py:226
    on = l2_normalize_rows(pooled)                               # losses.py:227
    z_bn = ((qn if per_image else qn.unsqueeze(0).expand(B, -1, -1)) * on).sum(dim=-1)    # losses.py:229-231
    if squeeze_quirk:
        z = z_bn.squeeze().T if z_bn.squeeze().dim() == 2 else z_bn.squeeze()
    else:
        z = z_bn.transpose(0, 1)
    return z, ([scores] if need_scores else None)


# --------------------------------------------------------------------------- a5
def multi_positive_nce_loss(logits: torch.Tensor, group_map: torch.Tensor,
                            temperature: torch.Tensor | float = 1.0, eps: float = NCE_EPS,
                            row_sum: bool = False, col_sum: bool = False) -> torch.Tensor:
    """multi_positive_nce_loss + get_row_loss + get_col_loss, losses.py:243-344.

    logits (N_total, B_global); group_map (N_total,) int64 = source image of each row.
    Raw exp (no log-sum-exp shift) and eps both inside the ratio and inside the log,
    exactly as the reference does.
    """
    n_total, b_global = logits.shape
    E = torch.exp(logits / temperature)                          # losses.py:265
    rows = torch.arange(n_total, device=logits.device)
    pos = E[rows, group_map]                                     # losses.py:267-269

    if row_sum:                                                  # losses.py:303-315
        rs = torch.zeros(b_global, dtype=E.dtype, device=E.device)
        ps = torch.zeros(b_global, dtype=E.dtype, device=E.device)
        rs.index_add_(0, group_map, E.sum(dim=1))
        ps.index_add_(0, group_map, pos)
        p_row = ps / (rs + eps)
    else:                                                        # losses.py:316-318
        p_row = pos / (E.sum(dim=1) + eps)
    row_loss = -torch.log(p_row + eps)                           # losses.py:320

    onehot = torch.zeros_like(E)
    onehot[rows, group_map] = 1.0
    if col_sum:                                                  # losses.py:331-336
        p_col = (E * onehot).sum(dim=0) / (E.sum(dim=0) + eps)
    else:                                                        # losses.py:337-342
        cneg = (E * (1.0 - onehot)).sum(dim=0)
        p_col = pos / (pos + cneg[group_map] + eps)
    col_loss = -torch.log(p_col + eps)                           # losses.py:344
    return (row_loss.mean() + col_loss.mean()) / 2               # losses.py:291


# --------------------------------------------------------------------------- a2
def build_group_map(counts: Sequence[int], rank: int = 0, device=None) -> torch.Tensor:
    """group index per sentence: image i of this rank -> i + rank * B_local.

    compute_text_features, losses.py:131-151 (``global_index = i + local_rank * B_local``).
    """
    b_local = len(counts)
    out = []
    for i, c in enumerate(counts):
        out.extend([i + rank * b_local] * int(c))
    return torch.tensor(out, dtype=torch.long, device=device)


# --------------------------------------------------------------------------- a4
def radzero_forward(text_features_list: Sequence[torch.Tensor], vision_tokens: torch.Tensor,
                    gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                    log_tau: torch.Tensor, *, attn_log_tau: Optional[torch.Tensor] = None,
                    use_vision_cls_token: bool = True, sim_op: str = "cos",
                    need_attn_weights: bool = False, compute_loss: bool = True,
                    row_sum: bool = False, col_sum: bool = False,
                    hidden_dim: int = 768, squeeze_quirk: bool = True) -> Dict:
    """RadZeroLoss.forward on one process (ddp_gather=False), losses.py:71-166.

    ``text_features_list[i]`` is what ``forward_text_model`` returned under
    ``text_features_wo_l2_norm`` for image / prompt i, shape (n_i, D) (or (n_i, 2D), in
    which case the last D columns are used, losses.py:144-145).
    """
    feats = []
    for f in text_features_list:
        if f.shape[-1] == 2 * hidden_dim:
            f = f[:, hidden_dim:]
        feats.append(f)
    text = torch.cat(feats, dim=0)                               # losses.py:153
    group_map = build_group_map([f.shape[0] for f in feats], device=text.device)
    if gamma is not None:
        text = layer_norm_rows(text, gamma, beta)                # losses.py:163-164
        vision_tokens = layer_norm_rows(vision_tokens, gamma, beta)  # losses.py:90-91
    attn_tokens = vision_tokens if use_vision_cls_token else vision_tokens[:, 1:]
    tau_attn = torch.exp(attn_log_tau if attn_log_tau is not None else log_tau)
    z, scores = similarity_logit(text, attn_tokens, temperature=tau_attn, sim_op=sim_op,
                                 need_scores=need_attn_weights, squeeze_quirk=squeeze_quirk)
    out = {"t2i_logits": z, "t2i_attn_weights": scores, "group_map": group_map}
    if compute_loss:
        loss = multi_positive_nce_loss(z, group_map, temperature=torch.exp(log_tau),
                                       row_sum=row_sum, col_sum=col_sum)
        out["losses"] = {"t2i_loss": loss, "loss": loss}         # losses.py:119-123
    return out


# --------------------------------------------------------------------------- a6
def compute_logits_glue(t2i_logits: torch.Tensor, scores: torch.Tensor, log_tau: torch.Tensor,
                        use_vision_cls_token: bool = True) -> Dict:
    """radzero branch of CxrAlignModel.compute_logits, modeling.py:309-328.

    ``similarity_scores`` = mean over the 1-element list, CLS column dropped;
    ``logits`` = Z^T / tau.  (The broken ``compute_i2t_loss`` read at modeling.py:320 is
    taken as False -- SURVEY.md section 0.4.)
    """
    sim = torch.stack([scores]).mean(dim=0)                      # modeling.py:311-313
    if use_vision_cls_token:
        sim = sim[:, :, 1:]                                      # modeling.py:316-317
    logits = t2i_logits.T / torch.exp(log_tau)                   # modeling.py:324-328
    return {"logits": logits, "similarity_scores": sim}


def similarity_prob_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """similarity_prob = sigmoid(logits) (README.md:104-106; segmentation_utils.py:225).

    INFERRED: the hub-side ``model_inference`` body is not in the reference checkout.
    """
    return torch.sigmoid(logits)


# --------------------------------------------------------------------------- a7 / K8
def bilinear_upsample(grid: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """F.interpolate(mode="bilinear", align_corners=False) on the trailing 2 dims.

    Own index arithmetic (ATen upsample_bilinear2d semantics):
        src = max(0, (dst + 0.5) * in/out - 0.5), i0 = floor(src), i1 = min(i0+1, in-1).
    segmentation_utils.py:64-69, grounding_utils.py:194-199.
    """
    in_h, in_w = grid.shape[-2], grid.shape[-1]
    dt, dev = grid.dtype, grid.device

    def axis(n_in: int, n_out: int):
        scale = torch.tensor(n_in / n_out, dtype=dt, device=dev)
        dst = torch.arange(n_out, dtype=dt, device=dev)
        src = torch.clamp((dst + 0.5) * scale - 0.5, min=0.0)
        i0 = src.floor().to(torch.long).clamp(max=n_in - 1)
        i1 = torch.clamp(i0 + 1, max=n_in - 1)
        w1 = src - i0.to(dt)
        return i0, i1, 1.0 - w1, w1

    y0, y1, wy0, wy1 = axis(in_h, out_h)
    x0, x1, wx0, wx1 = axis(in_w, out_w)
    top = grid[..., y0, :]
    bot = grid[..., y1, :]
    t = top[..., x0] * wx0 + top[..., x1] * wx1
    b = bot[..., x0] * wx0 + bot[..., x1] * wx1
    return t * wy0.unsqueeze(-1) + b * wy1.unsqueeze(-1)


def interpolate_similarity_scores(similarity_scores: torch.Tensor, origin_size: Tuple[int, int],
                                  processor_kind: str = "blip", fill: float = -999.0
                                  ) -> torch.Tensor:
    """interpolate_similarity_scores, segmentation_utils.py:36-122, returns (1, H, W).

    ``processor_kind`` stands for the reference's isinstance dispatch:
      "blip"        BlipImageProcessor              :62-70   plain resize to (H, W)
      "aspect_blip" AspectRatioBlipImageProcessor   :41-60   resize to max(H,W)^2, crop centre
      "bit"         BitImageProcessor               :72-91   resize to min(H,W)^2, paste, -999 fill
      "m3ae"        M3AEImageProcessor              :92-121  224/256 centre crop inside a pad
    """
    height, width = origin_size
    p = int(similarity_scores.shape[-1] ** 0.5)
    grid = similarity_scores.reshape(p, p)
    if processor_kind == "blip":
        out = bilinear_upsample(grid, height, width)
    elif processor_kind == "aspect_blip":
        side = max(height, width)
        full = bilinear_upsample(grid, side, side)
        left, top = (side - width) // 2, (side - height) // 2
        out = full[top:top + height, left:left + width]
    elif processor_kind == "bit":
        side = min(height, width)
        small = bilinear_upsample(grid, side, side)
        left, top = (width - side) // 2, (height - side) // 2
        out = torch.full((height, width), fill, dtype=grid.dtype, device=grid.device)
        out[top:top + side, left:left + side] = small
    elif processor_kind == "m3ae":
        side = max(height, width)
        crop = int(side * 224 / 256)
        small = bilinear_upsample(grid, crop, crop)
        canvas = torch.full((side, side), fill, dtype=grid.dtype, device=grid.device)
        off = (side - crop) // 2
        canvas[off:off + crop, off:off + crop] = small
        left, top = (side - width) // 2, (side - height) // 2
        out = canvas[top:top + height, left:left + width]
    else:
        raise NotImplementedError(processor_kind)
    return out.unsqueeze(0)


def grounding_point(similarity_score: torch.Tensor, image_size: Tuple[int, int],
                    processor_kind: str = "blip") -> Tuple[int, int]:
    """get_grounding_point, grounding_utils.py:166-261: first global argmax -> (x, y)."""
    height, width = image_size
    m = interpolate_similarity_scores(similarity_score, image_size, processor_kind)[0]
    flat = int(torch.argmax(m.reshape(-1)).item())               # grounding_utils.py:254
    return flat % width, flat // width                           # :256-259  (x, y)


def zero_shot_labels(logits: torch.Tensor) -> torch.Tensor:
    """argmax over prompts per image -- external/CARZero/inference.py:328-331."""
    return torch.argmax(logits, dim=1)


# --------------------------------------------------------------------------- C4 helper
def contrastive_step_reference(text: torch.Tensor, group_map: torch.Tensor,
                               vision_tokens: torch.Tensor, gamma: torch.Tensor,
                               beta: torch.Tensor, log_tau: torch.Tensor, sim_op: str = "cos"):
    """Forward + backward of the contrastive step through autograd on the oracle.

    Returns (loss, dict of grads for text / vision_tokens / gamma / beta / log_tau).
    Mirrors RadZeroLoss.forward(compute_loss=True) (losses.py:71-124) for one process.
    """
    leaves = [t.detach().clone().requires_grad_(True)
              for t in (text, vision_tokens, gamma, beta, log_tau)]
    t, x, g, b, lt = leaves
    tn = layer_norm_rows(t, g, b)
    xn = layer_norm_rows(x, g, b)
    tau = torch.exp(lt)
    z, _ = similarity_logit(tn, xn, temperature=tau, sim_op=sim_op, squeeze_quirk=False)
    loss = multi_positive_nce_loss(z, group_map, temperature=tau)
    loss.backward()
    return loss.detach(), {
        "text": t.grad, "vision_tokens": x.grad, "gamma": g.grad, "beta": b.grad,
        "log_tau": lt.grad, "t2i_logits": z.detach(),
    }


# --------------------------------------------------------------------------- a7 consumers
def dice_sweep_stats(similarity_scores: torch.Tensor, masks: torch.Tensor, origin_size: Tuple[int, int],
                     image_processor="blip", thresholds=None) -> Dict[str, torch.Tensor]:
    """Counts behind the Dice threshold sweep of exp/cxr_pt/inference/segmentation_utils.py:255-261:
    for every map, prob = sigmoid(interpolate(scores)) (:222-225) and, for every
    t in np.arange(0, 1.01, 0.01), pred_t = (prob > t) (:258); returns |pred_t|, |pred_t & mask|, |mask|
    and max prob per map.  (The reference feeds pred_t to torchmetrics' DiceScore, which is not
    vendored; the counts are what any Dice aggregation is computed from.)"""
    import numpy as np
    thresholds = np.arange(0, 1.01, 0.01) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    s = similarity_scores.reshape(-1, similarity_scores.shape[-1])
    pred, inter, mx = [], [], []
    for m in range(s.shape[0]):
        prob = torch.sigmoid(interpolate_similarity_scores(s[m], origin_size, image_processor)[0])
        g = masks[m] != 0
        pred.append(torch.stack([(prob > float(t)).sum() for t in thresholds]))
        inter.append(torch.stack([((prob > float(t)) & g).sum() for t in thresholds]))
        mx.append(prob.max())
    return {"pred": torch.stack(pred), "inter": torch.stack(inter), "gt": (masks != 0).flatten(1).sum(1),
            "max_prob": torch.stack(mx), "thresholds": torch.as_tensor(thresholds)}


def dice_score_samplewise(pred_masks: torch.Tensor, target_masks: torch.Tensor) -> torch.Tensor:
    """What ``DiceScore(num_classes=1)(preds, target)`` of the reference's sweep returns
    (segmentation_utils.py:255-258).  torchmetrics is a third-party dependency the reference does not
    vendor (``torchmetrics==1.6.1``, requirements.txt:242) and is absent here, so its published algorithm
    is restated (functional/segmentation/dice.py): per sample and class, numerator = 2 * sum(preds *
    target), denominator = sum(preds) + sum(target) over the spatial axes; ``average="micro"`` sums both
    over the class axis (one class here); dice = numerator / denominator with 1.0 where the denominator is
    0; the metric is the nan-mean over samples.  Parity unpinned against the library itself -- the
    formula is pinned by a hand-computed case in tests/test_oracle_golden.py."""
    p = (pred_masks != 0).flatten(1).double()
    g = (target_masks != 0).flatten(1).double()
    num = 2.0 * (p * g).sum(1)
    den = p.sum(1) + g.sum(1)
    dice = torch.where(den > 0, num / den.clamp_min(1.0), torch.ones_like(den))
    return dice.nanmean()


def compute_specificity(negative_probs: torch.Tensor, threshold: float) -> float:
    """segmentation_utils.py:136-158: share of negative images with no pixel above the threshold."""
    tn = ((negative_probs > threshold).long().flatten(1).sum(-1) == 0).sum()
    return (tn / len(negative_probs)).item()
