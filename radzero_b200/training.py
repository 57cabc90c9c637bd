"""The contrastive training step of the VL-CABS path as ONE autograd node.

Replaces what autograd does in the reference for ``RadZeroLoss.forward(compute_loss=True)``
(exp/cxr_pt/model/losses.py:71-124): LayerNorm + L2 (rz_prep_rows), the fused similarity
forward (rz_sim_fwd, keeping log-sum-exp, |o| and the pooled vectors), the MP-NCE loss and its
closed-form dL/dZ (rz_mpnce_*), the closed-form similarity backward as three tensor-core GEMM
passes (rz_sim_bwd) and the normalisation backward (rz_prep_rows_bwd).

Multi-GPU: the reference all-gathers the vision tokens of every rank and
computes the full (N_total x B_global) problem on each of them (losses.py:87-88, 156-161).
Here the IMAGES stay where they are -- rank r owns columns [r*B_local, (r+1)*B_local) of the
logit matrix -- and only small tensors cross NVLink:
    forward   all-gather of the fp16 normalised sentence embeddings (N_total x 768) + group_map,
              all-reduce(sum) of the per-sentence row sums / positives, all-reduce of 3 scalars
    backward  all-reduce(sum) of dL/dq (N_total x 768 fp32); each rank keeps its own rows
Gradients returned to autograd are multiplied by the world size when ``ddp_compatible`` (the
default under torch.distributed): the reference's ``dist.nn.all_gather`` backward sums the
identical global loss of all W ranks and DDP then averages parameter gradients over W; the
scale makes this sharded step a drop-in under the same DDP wrapper.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import RzError


def _gather_rows(t: torch.Tensor, group=None):
    """Ragged all-gather along dim 0 -> (cat, sizes list)."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = torch.empty(world, device=t.device, dtype=torch.int64)
    dist.all_gather_into_tensor(sizes, n_local, group=group)
    sizes_l = [int(s) for s in sizes.tolist()]
    n_max = max(sizes_l)
    padded = t.new_zeros((n_max,) + tuple(t.shape[1:]))
    padded[: t.shape[0]] = t
    out = t.new_empty((world * n_max,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view((world, n_max) + tuple(t.shape[1:]))
    return torch.cat([out[r, : sizes_l[r]] for r in range(world)], dim=0), sizes_l


class _ContrastiveStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text, tokens, gamma, beta, log_tau, attn_log_tau, group_map, cfg):
        distributed = cfg["distributed"]
        K = cfg.get("ops", ops)                   # kernels; tests of the host logic inject a stand-in
        if cfg["sim_op"] != "cos":
            raise NotImplementedError("the fused training step implements sim_op='cos' (radzero.yaml:44)")
        world = dist.get_world_size() if distributed else 1
        rank = dist.get_rank() if distributed else 0
        B, L, _ = tokens.shape
        Lp = K.padded_tokens_bwd(L)
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        k16, _, _ = K.prep_rows(tokens.detach(), g, b, rows_per_group=L, rows_per_group_padded=Lp)
        k16 = k16.view(B, Lp, ops.HIDDEN)
        q16_local, _, _ = K.prep_rows(text.detach(), g, b)
        n_local = q16_local.shape[0]
        if distributed:
            q16, sizes = _gather_rows(q16_local)
            gm, _ = _gather_rows(group_map)
            row0 = sum(sizes[:rank])
        else:
            q16, gm, row0 = q16_local, group_map, 0
        a_lt = attn_log_tau if attn_log_tau is not None else log_tau
        fwd = K.sim_fwd(k16, q16, L, 1.0, log_tau_scale=a_lt.detach(), want_scores=cfg["need_scores"],
                        drop_cls=False, want_stats=True, want_pooled=True)
        z = fwd["z"]                                   # (N_total, B_local)
        n_total = z.shape[0]
        b_global = B * world
        col0 = rank * B
        loss = None
        dz = terms = None
        if cfg["compute_loss"]:
            inv_tau = float(torch.exp(-log_tau.detach()))
            rs, ps, cn, cp = K.mpnce_partials(z, gm, col0, inv_tau)
            if distributed:
                both = torch.stack([rs, ps])
                dist.all_reduce(both)
                rs, ps = both[0], both[1]
            terms, dz = K.mpnce_finish(z, gm, col0, b_global, inv_tau, rs, ps, cn, cp,
                                       row_sum=cfg["row_sum"], col_sum=cfg["col_sum"], want_dz=True)
            tsum = terms.clone()
            if distributed:
                dist.all_reduce(tsum)
            n_row = b_global if cfg["row_sum"] else n_total
            n_col = b_global if cfg["col_sum"] else n_total
            loss = (tsum[0] / n_row + tsum[1] / n_col) * 0.5
        ctx.K = K
        ctx.cfg = dict(cfg)
        ctx.meta = (B, L, Lp, n_local, row0, world, attn_log_tau is not None)
        ctx.save_for_backward(text, tokens, gamma, beta, log_tau, attn_log_tau, k16, q16, z, dz,
                              fwd["lse"], fwd["onorm"], fwd["pooled"], terms, fwd.get("p"), fwd.get("mref"),
                              fwd.get("lsum"))
        ctx.mark_non_differentiable(z)
        scores = fwd["scores"]
        if scores is None:
            scores = z.new_empty(0)
        ctx.mark_non_differentiable(scores)
        if loss is None:
            loss = z.new_zeros(())
        return loss, z, scores

    @staticmethod
    def backward(ctx, g_loss, _gz, _gs):
        (text, tokens, gamma, beta, log_tau, attn_log_tau, k16, q16, z, dz, lse, onorm, pooled,
         terms, p_un, mref, lsum) = ctx.saved_tensors
        K = ctx.K
        B, L, Lp, n_local, row0, world, has_attn = ctx.meta
        if dz is None:
            raise RzError("backward through the contrastive step needs compute_loss=True")
        distributed = ctx.cfg["distributed"]
        ddp_scale = float(world) if (distributed and ctx.cfg["ddp_compatible"]) else 1.0
        gl = g_loss.reshape(()).float() * ddp_scale
        dzs = dz * gl
        a_lt = attn_log_tau if has_attn else log_tau
        dq, dk, dlt_attn = K.sim_bwd(k16, q16, L, 1.0, z, dzs, lse, onorm, pooled, log_tau=a_lt.detach(),
                                     p=p_un, mref=mref, lsum=lsum)
        if distributed:
            dist.all_reduce(dq)
        dq_local = dq[row0: row0 + n_local].contiguous()
        dgamma = dbeta = None
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        dx_tok, dgamma, dbeta = K.prep_rows_bwd(tokens.detach(), g, b, dk, rows_per_group=L,
                                                rows_per_group_padded=Lp, native_dx=True)
        dx_txt, dgamma, dbeta = K.prep_rows_bwd(text.detach(), g, b, dq_local, dgamma=dgamma,
                                                dbeta=dbeta, accumulate=True, native_dx=True)
        # d/dlog(tau): the loss temperature through exp(Z/tau) (= -sum dZ*Z) and, when the
        # attention shares it (attn_temperature: null, radzero.yaml:43), the softmax scores
        d_loss_lt = -(terms[2] * gl).reshape(1)
        if has_attn:
            d_lt, d_alt = d_loss_lt, dlt_attn.reshape(1)
        else:
            d_lt, d_alt = d_loss_lt + dlt_attn.reshape(1), None
        dtok = dx_tok.view(B, L, ops.HIDDEN).to(tokens.dtype)
        dtxt = dx_txt.to(text.dtype)
        dgm = dgamma.to(gamma.dtype) if gamma is not None else None
        dbt = dbeta.to(beta.dtype) if beta is not None else None
        return dtxt, dtok, dgm, dbt, d_lt.to(log_tau.dtype), d_alt, None, None


def contrastive_step(loss_fn, text: torch.Tensor, group_map: torch.Tensor, tokens: torch.Tensor, *,
                     distributed: bool, need_attn_weights: bool = False, compute_loss: bool = True,
                     ddp_compatible: Optional[bool] = None, kernel_ops=None):
    """Run the fused step for a ``RadZeroLoss`` module; returns dict(loss, z, scores)."""
    gamma, beta = (loss_fn.layer_norm.weight, loss_fn.layer_norm.bias) if loss_fn.layer_norm is not None \
        else (None, None)
    cfg = dict(distributed=distributed, sim_op=loss_fn.sim_op, need_scores=need_attn_weights,
               compute_loss=compute_loss, row_sum=loss_fn.mpnce_row_sum, col_sum=loss_fn.mpnce_col_sum,
               ddp_compatible=distributed if ddp_compatible is None else ddp_compatible)
    if kernel_ops is not None:
        cfg["ops"] = kernel_ops
    loss, z, scores = _ContrastiveStep.apply(text, tokens, gamma, beta, loss_fn.loss_temperature,
                                             loss_fn.attn_temperature, group_map, cfg)
    return {"loss": loss if compute_loss else None, "z": z,
            "scores": scores if need_attn_weights else None}


class _SimilarityLogitFn(torch.autograd.Function):
    """SimilarityLogit alone under autograd (inputs already LayerNorm-ed, sim_op 'cos')."""

    @staticmethod
    def forward(ctx, queries, tokens, scale, need_scores):
        B, L, _ = tokens.shape
        Lp = ops.padded_tokens_bwd(L)
        k16, _, _ = ops.prep_rows(tokens.detach(), None, None, rows_per_group=L, rows_per_group_padded=Lp)
        k16 = k16.view(B, Lp, ops.HIDDEN)
        q16, _, _ = ops.prep_rows(queries.detach(), None, None)
        fwd = ops.sim_fwd(k16, q16, L, scale, want_scores=need_scores, drop_cls=False, want_stats=True,
                          want_pooled=True)
        ctx.save_for_backward(queries, tokens, k16, q16, fwd["z"], fwd["lse"], fwd["onorm"], fwd["pooled"],
                              fwd.get("p"), fwd.get("mref"), fwd.get("lsum"))
        ctx.meta = (B, L, Lp, scale)
        scores = fwd["scores"] if fwd["scores"] is not None else fwd["z"].new_empty(0)
        ctx.mark_non_differentiable(scores)
        return fwd["z"], scores

    @staticmethod
    def backward(ctx, gz, _gs):
        queries, tokens, k16, q16, z, lse, onorm, pooled, p_un, mref, lsum = ctx.saved_tensors
        B, L, Lp, scale = ctx.meta
        dq, dk, _ = ops.sim_bwd(k16, q16, L, scale, z, gz.float().contiguous(), lse, onorm, pooled,
                                p=p_un, mref=mref, lsum=lsum)
        dxt, _, _ = ops.prep_rows_bwd(tokens.detach(), None, None, dk, rows_per_group=L,
                                      rows_per_group_padded=Lp)
        dxq, _, _ = ops.prep_rows_bwd(queries.detach(), None, None, dq)
        return dxq.to(queries.dtype), dxt.view(B, L, ops.HIDDEN).to(tokens.dtype), None, None


def similarity_logit_autograd(queries, tokens, scale: float, l2: bool, need_scores: bool):
    if not l2:
        raise NotImplementedError("training through sim_op='dot' is not implemented on the B200 path")
    z, scores = _SimilarityLogitFn.apply(queries, tokens, scale, need_scores)
    return z, (scores if need_scores else None)
