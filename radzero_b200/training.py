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
    forward   ONE all-gather of the fp16 normalised sentence embeddings (N_total x 768) + group_map
              (fixed-capacity records, on a side stream under the token prep; the sentence counts
              travel over a CPU-side gloo channel, so the GPU queue is never drained),
              all-reduce(sum) of the per-sentence row sums / positives, all-reduce of the loss scalar
    backward  reduce-scatter(sum) of dL/dq (N_total x 768 fp32) to the ranks that own the rows
No call in the step reads device memory from the host: the temperatures are read by the kernels from
the parameters (``log_tau`` pointers).
Gradients returned to autograd are multiplied by the world size when ``ddp_compatible`` (the
default under torch.distributed): the reference's ``dist.nn.all_gather`` backward sums the
identical global loss of all W ranks and DDP then averages parameter gradients over W; the
scale makes this sharded step a drop-in under the same DDP wrapper.
"""
from __future__ import annotations

from typing import List, Optional

import math

import torch
import torch.distributed as dist

from . import ops
from ._lib import RzError


class _Comm:
    """Per-process helper of the sharded step: a CPU-side (gloo) channel for the sentence counts, so the
    ragged exchange needs NO device-to-host read (the GPU queue is never drained), and a side stream on
    which the text all-gather overlaps the token prep."""

    _inst = None

    def __init__(self):
        self.side_group = None
        self.side_failed = False
        self.stream = {}

    @classmethod
    def get(cls):
        if cls._inst is None or cls._inst.world != dist.get_world_size():
            cls._inst = cls()
            cls._inst.world = dist.get_world_size()
        return cls._inst

    def sizes(self, n_local: int, device) -> List[int]:
        """Sentence counts of all ranks.  gloo default group: a CPU all-gather; NCCL: a gloo side group
        created once (collective on first use); only if that fails, the NCCL exchange + a host read."""
        world = dist.get_world_size()
        backend = dist.get_backend()
        grp = None
        if backend != "gloo" and not self.side_failed:
            if self.side_group is None:
                try:
                    self.side_group = dist.new_group(backend="gloo")
                except Exception:          # pragma: no cover - gloo not built in
                    self.side_failed = True
            grp = self.side_group
        if backend == "gloo" or grp is not None:
            mine = torch.tensor([int(n_local)], dtype=torch.int64)
            out = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(out, mine, group=grp)
            return [int(t.item()) for t in out]
        n = torch.tensor([n_local], device=device, dtype=torch.int64)
        sizes = torch.empty(world, device=device, dtype=torch.int64)
        dist.all_gather_into_tensor(sizes, n)
        return [int(v) for v in sizes.tolist()]

    def side_stream(self, device):
        if device.type != "cuda":
            return None
        st = self.stream.get(device.index)
        if st is None:
            st = self.stream[device.index] = torch.cuda.Stream(device=device)
        return st


def _gather_text(q_local: torch.Tensor, gm_local: torch.Tensor, sizes: List[int], side=None):
    """ONE all-gather for the ragged (rows, group_map) pair: each rank contributes a fixed-capacity
    byte record [cap rows | cap int64 indices]; the valid prefixes are concatenated afterwards.  With
    ``side`` (a CUDA stream) the collective runs there; the caller waits on the returned event."""
    world = len(sizes)
    cap = max(max(sizes), 1)
    n_local = q_local.shape[0]
    row_b = q_local.shape[1] * q_local.element_size()
    rec = cap * (row_b + 8)
    send = torch.empty(rec, dtype=torch.uint8, device=q_local.device)
    recv = torch.empty((world, rec), dtype=torch.uint8, device=q_local.device)
    gm_local = gm_local.to(torch.int64)

    def run():
        send[: n_local * row_b].copy_(q_local.reshape(-1).view(torch.uint8))
        send[cap * row_b: cap * row_b + n_local * 8].copy_(gm_local.view(torch.uint8))
        dist.all_gather_into_tensor(recv.view(-1), send)

    ev = None
    if side is not None:
        side.wait_stream(torch.cuda.current_stream(q_local.device))
        with torch.cuda.stream(side):
            run()
            ev = torch.cuda.Event()
            ev.record(side)
    else:
        run()

    def unpack():
        if ev is not None:
            torch.cuda.current_stream(q_local.device).wait_event(ev)
        rows = [recv[r, : sizes[r] * row_b].view(q_local.dtype).view(sizes[r], -1) for r in range(world)]
        gms = [recv[r, cap * row_b: cap * row_b + sizes[r] * 8].view(torch.int64) for r in range(world)]
        return torch.cat(rows, dim=0), torch.cat(gms, dim=0)

    return unpack


def _scatter_dq(dq: torch.Tensor, sizes: List[int], rank: int) -> torch.Tensor:
    """dL/dq rows back to their owners: reduce-scatter (sum) of the per-rank sections, padded to a
    common capacity.  Half the NVLink traffic of the round-1 all-reduce."""
    world = len(sizes)
    cap = max(max(sizes), 1)
    if all(s == cap for s in sizes):
        send = dq
    else:
        send = torch.empty((world * cap,) + tuple(dq.shape[1:]), dtype=dq.dtype, device=dq.device)
        o = 0
        for r, s in enumerate(sizes):
            send[r * cap: r * cap + s].copy_(dq[o: o + s])
            o += s
    out = torch.empty((cap,) + tuple(dq.shape[1:]), dtype=dq.dtype, device=dq.device)
    if dist.get_backend() == "gloo":       # CPU tests: gloo has no reduce_scatter
        tmp = send.clone()
        dist.all_reduce(tmp)
        out.copy_(tmp[rank * cap: (rank + 1) * cap])
    else:
        dist.reduce_scatter_tensor(out, send)
    return out[: sizes[rank]]


_DOT_SCALE = 1.0 / math.sqrt(ops.HIDDEN)          # losses.py:215


def _inv_norms(q16: torch.Tensor) -> torch.Tensor:
    """1/|q_n| of the fp16 rows the kernels multiply with (F.normalize of the queries, losses.py:226)."""
    return 1.0 / q16.float().norm(dim=-1).clamp_min(1e-12)


class _ContrastiveStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text, tokens, gamma, beta, log_tau, attn_log_tau, group_map, cfg):
        distributed = cfg["distributed"]
        K = cfg.get("ops", ops)                   # kernels; tests of the host logic inject a stand-in
        if cfg["sim_op"] not in ("cos", "dot"):
            raise NotImplementedError(cfg["sim_op"])           # losses.py:216-217
        # sim_op "dot" (the constructor default, losses.py:45, 214-215): rows are LayerNorm-ed but not
        # L2-normalised, scores are divided by sqrt(768) and no attention temperature exists
        dot = cfg["sim_op"] == "dot"
        l2kw = dict(l2=False) if dot else {}
        world = dist.get_world_size() if distributed else 1
        rank = dist.get_rank() if distributed else 0
        B, L, _ = tokens.shape
        Lp = K.padded_tokens_bwd(L)
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        q16_local, _, _ = K.prep_rows(text.detach(), g, b, **l2kw)
        n_local = q16_local.shape[0]
        sizes = None
        unpack = None
        if distributed:
            comm = _Comm.get()
            sizes = comm.sizes(n_local, q16_local.device)
            # the text all-gather runs on a side stream while this stream normalises the tokens
            unpack = _gather_text(q16_local, group_map, sizes, comm.side_stream(q16_local.device))
        k16, _, _ = K.prep_rows(tokens.detach(), g, b, rows_per_group=L, rows_per_group_padded=Lp, **l2kw)
        k16 = k16.view(B, Lp, ops.HIDDEN)
        if distributed:
            q16, gm = unpack()
            row0 = sum(sizes[:rank])
        else:
            q16, gm, row0 = q16_local, group_map, 0
        if dot:
            fwd = K.sim_fwd(k16, q16, L, _DOT_SCALE, q_inv_norm=_inv_norms(q16), want_scores=cfg["need_scores"],
                            drop_cls=False, want_stats=True, want_pooled=True)
        else:
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
            # temperature read from the parameter on the device: no host synchronisation in the step
            lt = log_tau.detach().reshape(1)
            lt = lt.float() if z.is_cuda else lt
            rowpos = torch.empty((2, n_total), dtype=torch.float32 if z.is_cuda else z.dtype,
                                 device=z.device)      # rowsum | pos: ONE all-reduce message
            rs, ps, cn, cp = K.mpnce_partials(z, gm, col0, log_tau=lt, rowpos=rowpos, col_sum=cfg["col_sum"],
                                              b_global=b_global)
            if distributed:
                dist.all_reduce(rowpos)
            terms, dz = K.mpnce_finish(z, gm, col0, b_global, 1.0, rs, ps, cn, cp, log_tau=lt,
                                       row_sum=cfg["row_sum"], col_sum=cfg["col_sum"], want_dz=True)
            loss = terms[3].clone()            # this rank's share; the sum over ranks is the loss
            if distributed:
                dist.all_reduce(loss)
        ctx.K = K
        ctx.cfg = dict(cfg)
        ctx.meta = (B, L, Lp, n_local, row0, world, attn_log_tau is not None, sizes, rank)
        ctx.save_for_backward(text, tokens, gamma, beta, log_tau, attn_log_tau, k16, q16, z, dz,
                              fwd["lse"], fwd["onorm"], fwd["pooled"], terms, fwd.get("p"), fwd.get("mref"),
                              fwd.get("lsum"))
        z_out = z
        if distributed and cfg.get("gather_logits"):
            # the reference returns the full (N_total, B_global) matrix on every rank
            blocks = torch.empty((world,) + tuple(z.shape), dtype=z.dtype, device=z.device)
            dist.all_gather_into_tensor(blocks.view(-1), z.contiguous().view(-1))
            z_out = blocks.permute(1, 0, 2).reshape(n_total, b_global)
        ctx.mark_non_differentiable(z_out)
        scores = fwd["scores"]
        if scores is None:
            scores = z.new_empty(0)
        ctx.mark_non_differentiable(scores)
        if loss is None:
            loss = z.new_zeros(())
        return loss, z_out, scores

    @staticmethod
    def backward(ctx, g_loss, _gz, _gs):
        (text, tokens, gamma, beta, log_tau, attn_log_tau, k16, q16, z, dz, lse, onorm, pooled,
         terms, p_un, mref, lsum) = ctx.saved_tensors
        K = ctx.K
        B, L, Lp, n_local, row0, world, has_attn, sizes, rank = ctx.meta
        if dz is None:
            raise RzError("backward through the contrastive step needs compute_loss=True")
        distributed = ctx.cfg["distributed"]
        ddp_scale = float(world) if (distributed and ctx.cfg["ddp_compatible"]) else 1.0
        gl = g_loss.reshape(()).float() * ddp_scale
        dzs = dz * gl
        dot = ctx.cfg["sim_op"] == "dot"
        l2kw = dict(l2=False) if dot else {}
        if dot:
            dq, dk, _ = K.sim_bwd(k16, q16, L, _DOT_SCALE, z, dzs, lse, onorm, pooled, p=p_un, mref=mref,
                                  lsum=lsum, q_inv_norm=_inv_norms(q16))
            dlt_attn = None                        # no temperature inside the attention (losses.py:214-215)
        else:
            a_lt = attn_log_tau if has_attn else log_tau
            dq, dk, dlt_attn = K.sim_bwd(k16, q16, L, 1.0, z, dzs, lse, onorm, pooled, log_tau=a_lt.detach(),
                                         p=p_un, mref=mref, lsum=lsum)
        if distributed:
            dq_local = _scatter_dq(dq, sizes, rank).contiguous()
        else:
            dq_local = dq
        dgamma = dbeta = None
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        dx_tok, dgamma, dbeta = K.prep_rows_bwd(tokens.detach(), g, b, dk, rows_per_group=L,
                                                rows_per_group_padded=Lp, native_dx=True, **l2kw)
        dx_txt, dgamma, dbeta = K.prep_rows_bwd(text.detach(), g, b, dq_local, dgamma=dgamma,
                                                dbeta=dbeta, accumulate=True, native_dx=True, **l2kw)
        # d/dlog(tau): the loss temperature through exp(Z/tau) (= -sum dZ*Z) and, when the
        # attention shares it (attn_temperature: null, radzero.yaml:43), the softmax scores
        d_loss_lt = -(terms[2] * gl).reshape(1)
        if dlt_attn is None:
            d_lt, d_alt = d_loss_lt, None
        elif has_attn:
            d_lt, d_alt = d_loss_lt, dlt_attn.reshape(1)
        else:
            d_lt, d_alt = d_loss_lt + dlt_attn.reshape(1), None
        dtok = dx_tok.view(B, L, ops.HIDDEN).to(tokens.dtype)
        dtxt = dx_txt.to(text.dtype)
        dgm = dgamma.to(gamma.dtype) if gamma is not None else None
        dbt = dbeta.to(beta.dtype) if beta is not None else None
        if d_alt is None and has_attn:
            d_alt = torch.zeros_like(attn_log_tau)
        return dtxt, dtok, dgm, dbt, d_lt.to(log_tau.dtype), d_alt, None, None


def contrastive_step(loss_fn, text: torch.Tensor, group_map: torch.Tensor, tokens: torch.Tensor, *,
                     distributed: bool, need_attn_weights: bool = False, compute_loss: bool = True,
                     ddp_compatible: Optional[bool] = None, kernel_ops=None, gather_logits: bool = False):
    """Run the fused step for a ``RadZeroLoss`` module; returns dict(loss, z, scores)."""
    gamma, beta = (loss_fn.layer_norm.weight, loss_fn.layer_norm.bias) if loss_fn.layer_norm is not None \
        else (None, None)
    cfg = dict(distributed=distributed, sim_op=loss_fn.sim_op, need_scores=need_attn_weights,
               compute_loss=compute_loss, row_sum=loss_fn.mpnce_row_sum, col_sum=loss_fn.mpnce_col_sum,
               ddp_compatible=distributed if ddp_compatible is None else ddp_compatible,
               gather_logits=gather_logits)
    if kernel_ops is not None:
        cfg["ops"] = kernel_ops
    loss, z, scores = _ContrastiveStep.apply(text, tokens, gamma, beta, loss_fn.loss_temperature,
                                             loss_fn.attn_temperature, group_map, cfg)
    return {"loss": loss if compute_loss else None, "z": z,
            "scores": scores if need_attn_weights else None}


class _SimilarityLogitFn(torch.autograd.Function):
    """SimilarityLogit alone under autograd (inputs already LayerNorm-ed; sim_op 'cos' = l2, 'dot' = not)."""

    @staticmethod
    def forward(ctx, queries, tokens, scale, need_scores, l2):
        B, L, _ = tokens.shape
        Lp = ops.padded_tokens_bwd(L)
        k16, _, _ = ops.prep_rows(tokens.detach(), None, None, rows_per_group=L, rows_per_group_padded=Lp, l2=l2)
        k16 = k16.view(B, Lp, ops.HIDDEN)
        q16, _, _ = ops.prep_rows(queries.detach(), None, None, l2=l2)
        fwd = ops.sim_fwd(k16, q16, L, scale, want_scores=need_scores, drop_cls=False, want_stats=True,
                          want_pooled=True, q_inv_norm=None if l2 else _inv_norms(q16))
        ctx.save_for_backward(queries, tokens, k16, q16, fwd["z"], fwd["lse"], fwd["onorm"], fwd["pooled"],
                              fwd.get("p"), fwd.get("mref"), fwd.get("lsum"))
        ctx.meta = (B, L, Lp, scale, l2)
        scores = fwd["scores"] if fwd["scores"] is not None else fwd["z"].new_empty(0)
        ctx.mark_non_differentiable(scores)
        return fwd["z"], scores

    @staticmethod
    def backward(ctx, gz, _gs):
        queries, tokens, k16, q16, z, lse, onorm, pooled, p_un, mref, lsum = ctx.saved_tensors
        B, L, Lp, scale, l2 = ctx.meta
        gz = gz.float()
        if gz.stride() != z.stride():
            gz = gz.contiguous()
            z = z.contiguous()
        dq, dk, _ = ops.sim_bwd(k16, q16, L, scale, z, gz, lse, onorm, pooled, p=p_un, mref=mref, lsum=lsum,
                                q_inv_norm=None if l2 else _inv_norms(q16))
        dxt, _, _ = ops.prep_rows_bwd(tokens.detach(), None, None, dk, rows_per_group=L,
                                      rows_per_group_padded=Lp, l2=l2)
        dxq, _, _ = ops.prep_rows_bwd(queries.detach(), None, None, dq, l2=l2)
        return dxq.to(queries.dtype), dxt.view(B, L, ops.HIDDEN).to(tokens.dtype), None, None, None


def similarity_logit_autograd(queries, tokens, scale: float, l2: bool, need_scores: bool):
    z, scores = _SimilarityLogitFn.apply(queries, tokens, scale, need_scores, l2)
    return z, (scores if need_scores else None)
