"""Mirror of the reference's ``AlignTransformer`` (exp/cxr_pt/model/align_transformers.py:23-45).

The reference wraps a transformers ``Dinov2Encoder`` of ``num_hidden_layers`` layers (2 in
radzero.yaml:29-33) and an optional final ``nn.LayerNorm`` (``use_layer_norm``, False in the released
configuration); ``forward(vision_tokens) -> vision_tokens``.  It is the only trainable vision
compute and the producer of the VL-CABS path's input.

Here the module keeps the SAME sub-module names (``transformer_layers`` = the HF encoder object,
``layer_norm``), so a reference state dict loads unchanged, but its forward never calls them: the
weights are packed once to fp16 GEMM operands and every layer runs on the hand-written sm_100a
kernels behind the C ABI (``rz_ln_rows``, ``rz_linear``, ``rz_attention``):

    h16 = LN1(x)                       rz_ln_rows          (eps 1e-6)
    qkv = h16 [Wq/8 | Wk | Wv]^T + b   rz_linear  "bias"   (one GEMM, 1/sqrt(64) folded into Wq, bq)
    a16 = softmax(q k^T) v             rz_attention        (tcgen05, online softmax)
    x   = x + ls1 * (a16 Wo^T + bo)    rz_linear  "residual" (in place on the fp32 residual stream)
    h16 = LN2(x)                       rz_ln_rows
    g16 = gelu(h16 W1^T + b1)          rz_linear  "gelu"
    x   = x + ls2 * (g16 W2^T + b2)    rz_linear  "residual"

Training (`module_to_update: [align_transformer, ...]`, radzero.yaml): when autograd needs gradients
through the module, ``forward`` runs the same kernels under ``_AlignFn`` (a ``torch.autograd.Function``
that keeps each layer's activations) and its backward runs the hand-written chain of ``layer_backward``:
the nine GEMM-shaped products per layer on ``rz_linear`` (dX = dY W with pre-transposed weights;
dW = dY^T X with both operands transposed to K-major by ``rz_transpose_pad``, accumulated in fp32 by the
residual epilogue), ``rz_attention_bwd``, ``rz_ln_rows_bwd``, ``rz_gelu_bwd``, ``rz_ls_cast_bwd``.  The fp16
gradient chain carries one power-of-two scale chosen on the device (``rz_grad_scale``), so small loss
gradients do not fall into fp16 subnormals and no host synchronisation is needed.  The stock HF modules
are never called (``kernel_backward = False`` restores ``stock_forward`` for comparisons).
No CPU fallback: CPU tensors raise ``RzError``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import RzError

HEAD_DIM = 64


def pack_layer(layer: nn.Module, device=None) -> Dict[str, torch.Tensor]:
    """fp16 GEMM operands + fp32 vectors of one ``Dinov2Layer`` (transformers modeling_dinov2)."""
    att = layer.attention.attention
    if att.attention_head_size != HEAD_DIM:
        raise RzError(f"the attention kernel is built for head dim {HEAD_DIM}")
    if not hasattr(layer.mlp, "fc1"):
        raise RzError("SwiGLU feed-forward layers are not on the RadZero path (use_swiglu_ffn=False)")
    s = float(att.scaling)            # 1/8: a power of two, so folding it into Wq / bq is exact (up to fp16 subnormals)
    dev = device or att.query.weight.device
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    f16 = lambda t: t.detach().to(device=dev, dtype=torch.float32).to(torch.float16).contiguous()

    def bias_of(lin):
        return lin.bias if lin.bias is not None else torch.zeros(lin.out_features, device=lin.weight.device)

    return {
        "heads": att.num_attention_heads,
        "eps1": layer.norm1.eps, "g1": f32(layer.norm1.weight), "b1": f32(layer.norm1.bias),
        "eps2": layer.norm2.eps, "g2": f32(layer.norm2.weight), "b2": f32(layer.norm2.bias),
        "wqkv": f16(torch.cat([att.query.weight * s, att.key.weight, att.value.weight], dim=0)),
        "bqkv": f32(torch.cat([bias_of(att.query) * s, bias_of(att.key), bias_of(att.value)], dim=0)),
        "wo": f16(layer.attention.output.dense.weight), "bo": f32(layer.attention.output.dense.bias),
        "ls1": f32(layer.layer_scale1.lambda1),
        "w1": f16(layer.mlp.fc1.weight), "bf1": f32(layer.mlp.fc1.bias),
        "w2": f16(layer.mlp.fc2.weight), "bf2": f32(layer.mlp.fc2.bias),
        "ls2": f32(layer.layer_scale2.lambda1),
        "q_scale": s,
    }


def pack_layer_bwd(layer: nn.Module, device=None) -> Dict[str, torch.Tensor]:
    """The transposed fp16 weights the dX = dY W products read ((in, out) row-major = rz_linear's
    (N, K) operand).  The query block is the UNSCALED weight: ``rz_attention_bwd`` hands back the gradient
    of the unscaled projections."""
    att = layer.attention.attention
    dev = device or att.query.weight.device
    t16 = lambda t: t.detach().to(device=dev, dtype=torch.float32).t().to(torch.float16).contiguous()
    return {
        "wqkv_t": t16(torch.cat([att.query.weight, att.key.weight, att.value.weight], dim=0)),   # (768, 2304)
        # the LayerScale factors ride in the transposed weights: (ls * dy) W = dy (diag(ls) W)
        "wo_t": t16(layer.attention.output.dense.weight * layer.layer_scale1.lambda1[:, None]),  # (768, 768)
        "w1_t": t16(layer.mlp.fc1.weight),                                                      # (768, 3072)
        "w2_t": t16(layer.mlp.fc2.weight * layer.layer_scale2.lambda1[:, None]),                 # (3072, 768)
        # fp32 weights for dls = rowsum(W * dW_unscaled) (rz_ls_weight_bwd)
        "wo32": layer.attention.output.dense.weight.detach().to(device=dev, dtype=torch.float32).contiguous(),
        "w232": layer.mlp.fc2.weight.detach().to(device=dev, dtype=torch.float32).contiguous(),
    }


def layer_params(layer: nn.Module) -> List[Optional[nn.Parameter]]:
    """The parameters of one Dinov2Layer in the order ``layer_backward`` returns their gradients."""
    att = layer.attention.attention
    return [layer.norm1.weight, layer.norm1.bias,
            att.query.weight, att.query.bias, att.key.weight, att.key.bias, att.value.weight, att.value.bias,
            layer.attention.output.dense.weight, layer.attention.output.dense.bias, layer.layer_scale1.lambda1,
            layer.norm2.weight, layer.norm2.bias,
            layer.mlp.fc1.weight, layer.mlp.fc1.bias, layer.mlp.fc2.weight, layer.mlp.fc2.bias,
            layer.layer_scale2.lambda1]


def layer_forward_train(x2: torch.Tensor, B: int, L: int, w: Dict[str, torch.Tensor]):
    """``layer_forward`` out of place, keeping what the backward reads: returns (z (B L, 768) fp32, saved)."""
    D = x2.shape[1]
    h1 = ops.ln_rows(x2, w["g1"], w["b1"], w["eps1"])
    qkv = ops.linear(h1, w["wqkv"], w["bqkv"], "bias")
    a = ops.attention(qkv.view(B, L, 3 * D), w["heads"]).view(B * L, D)
    y = ops.linear(a, w["wo"], w["bo"], "residual", scale=w["ls1"], residual=x2)
    h2 = ops.ln_rows(y, w["g2"], w["b2"], w["eps2"])
    g = ops.linear(h2, w["w1"], w["bf1"], "gelu")
    z = ops.linear(g, w["w2"], w["bf2"], "residual", scale=w["ls2"], residual=y)
    return z, (x2, h1, qkv, a, y, h2, g)


def _dweight(dyT: torch.Tensor, xT: torch.Tensor, sc: torch.Tensor) -> torch.Tensor:
    """dW (out, in) fp32 = 2^-k dY^T X from the K-major transposes (K = rows, zero padded)."""
    out = torch.zeros((dyT.shape[0], xT.shape[0]), dtype=torch.float32, device=dyT.device)
    return ops.linear(dyT, xT, None, "residual", scale=sc[4:4 + xT.shape[0]], residual=out, out=out)


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
OVERLAP_WEIGHT_GRADS = os.environ.get("RZ_ALIGN_BWD_OVERLAP", "1") != "0"


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One high-priority stream per device for the weight-gradient branch of the backward."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=dev, priority=-1)
    return _SIDE_STREAMS[idx]


class _Branch:
    """The weight-gradient branch of a layer's backward.  dW = dY^T X (and the bias / LayerScale
    gradients hanging off it) feeds nothing else in the chain, and its GEMMs are short of CTAs (768 x 768
    outputs = 18 tiles on 148 SMs), so it runs on a high-priority side stream next to the dX chain and the
    attention backward instead of leaving SMs idle in line.  Inputs made on the main stream are handed over
    with an event + ``record_stream`` (the caching allocator must not recycle them while the side stream
    still reads them); ``join`` makes the main stream wait and hands the results back the same way."""

    def __init__(self, dev: torch.device, enabled: bool):
        enabled = enabled and dev.type == "cuda"
        self.main = torch.cuda.current_stream(dev) if enabled else None
        self.side = _side_stream(dev) if enabled else None
        self.outs: List[torch.Tensor] = []

    def run(self, fn, *inputs: torch.Tensor):
        if self.side is None:
            return fn()
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)
        for t in inputs:
            t.record_stream(self.side)
        with torch.cuda.stream(self.side):
            res = fn()
        self.outs.extend(r for r in res if torch.is_tensor(r))
        return res

    def join(self) -> None:
        if self.side is None:
            return
        self.main.wait_stream(self.side)
        for t in self.outs:
            t.record_stream(self.main)
        self.outs = []


def layer_backward(dz: torch.Tensor, saved, B: int, L: int, w: Dict[str, torch.Tensor],
                   wb: Dict[str, torch.Tensor], sc: torch.Tensor, branch: Optional[_Branch] = None,
                   dz16: Optional[torch.Tensor] = None, want_dx16: bool = False):
    """Autograd of one Dinov2Layer.  ``dz`` (B L, 768) fp32 = dL/d(layer output); returns
    (dL/d(layer input) fp32, gradients in ``layer_params`` order, fp32, fp16(2^k dL/d(layer input)) | None).
    ``dz16`` = fp16(2^k dz) when the layer above already produced it (``want_dx16``: the LayerNorm backward
    writes the fp16 operand of the next product from the registers that hold the fp32 result).  The caller
    joins ``branch`` before it reads the gradients."""
    x2, h1, qkv, a, y, h2, g = saved
    D = x2.shape[1]
    dev = dz.device
    br = branch or _Branch(dev, False)
    zeros = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)

    def wgrad(dy16, x16, bias_n, ls=None, w32=None, bias=None):
        """(dW, db, dls | None) of ``o = x W^T + b`` (times LayerScale) from dy16 = fp16(2^k dL/do) and x."""
        db = zeros(bias_n)
        dw = _dweight(ops.transpose_pad(dy16, sc, db), ops.transpose_pad(x16), sc)
        # LayerScale is folded into the transposed weights of the dX product and finished here from the
        # unscaled weight gradient (dW, db, dls), so the scaled product is never recomputed
        dls = ops.ls_weight_bwd(dw, w32, bias, db, ls) if ls is not None else None
        return dw, db, dls

    # ---- x = y + ls2 * (gelu(h2 W1^T + b1) W2^T + b2)
    do2 = dz16 if dz16 is not None else ops.ls_cast_bwd(dz, None, None, sc, None)
    dw2, db2, dls2 = br.run(lambda: wgrad(do2, g, D, w["ls2"], wb["w232"], w["bf2"]), do2, g, sc)
    dg = ops.linear(do2, wb["w2_t"], None, "bias")
    del do2
    u = ops.linear(h2, w["w1"], w["bf1"], "bias")                    # recomputed pre-activation
    du = ops.gelu_bwd(dg, u)
    del dg, u
    dw1, db1, _ = br.run(lambda: wgrad(du, h2, 4 * D), du, h2, sc)
    dh2 = ops.linear(du, wb["w1_t"], None, "bias")
    del du
    dg2, dbeta2 = zeros(D), zeros(D)
    do1 = torch.empty((dz.shape[0], D), dtype=torch.float16, device=dev)
    dy = ops.ln_rows_bwd(y, dh2, w["g2"], w["eps2"], dz, sc, dg2, dbeta2, out16=do1)
    del dh2
    # ---- y = x + ls1 * (attention(LN1(x)) Wo^T + bo)
    dwo, dbo, dls1 = br.run(lambda: wgrad(do1, a, D, w["ls1"], wb["wo32"], w["bo"]), do1, a, sc)
    da = ops.linear(do1, wb["wo_t"], None, "bias")
    del do1
    dqkv = ops.attention_bwd(qkv.view(B, L, 3 * D), a.view(B, L, D), da.view(B, L, D), w["heads"],
                             w["q_scale"]).view(B * L, 3 * D)
    del da
    dwqkv, dbqkv, _ = br.run(lambda: wgrad(dqkv, h1, 3 * D), dqkv, h1, sc)
    dh1 = ops.linear(dqkv, wb["wqkv_t"], None, "bias")
    del dqkv
    dg1, dbeta1 = zeros(D), zeros(D)
    dx16 = torch.empty((dz.shape[0], D), dtype=torch.float16, device=dev) if want_dx16 else None
    dx = ops.ln_rows_bwd(x2, dh1, w["g1"], w["eps1"], dy, sc, dg1, dbeta1, out=dy, out16=dx16)
    grads = [dg1, dbeta1,
             dwqkv[:D], dbqkv[:D], dwqkv[D:2 * D], dbqkv[D:2 * D], dwqkv[2 * D:], dbqkv[2 * D:],
             dwo, dbo, dls1, dg2, dbeta2, dw1, db1, dw2, db2, dls2]
    return dx, grads, dx16


class _AlignFn(torch.autograd.Function):
    """The encoder layers under autograd on the B200 kernels (forward keeps the activations)."""

    @staticmethod
    def forward(ctx, tokens, module, *params):
        B, L, D = tokens.shape
        layers = module._weights(tokens.device)
        x = tokens.detach()
        x = (x if x.dtype == torch.float32 else x.to(torch.float32)).contiguous().view(B * L, D)
        saved = []
        for w in layers:
            x, sv = layer_forward_train(x, B, L, w)
            saved.append(sv)
        ctx.module, ctx.saved, ctx.shape, ctx.params = module, saved, (B, L, D), params
        ctx.in_dtype = tokens.dtype
        return x.view(B, L, D).to(tokens.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        B, L, D = ctx.shape
        module = ctx.module
        layers = module._weights(dout.device)
        back = module._weights_bwd(dout.device)
        dz = dout.to(torch.float32).contiguous().view(B * L, D)
        sc = ops.grad_scale(dz)
        per_layer = []
        branch = _Branch(dout.device, OVERLAP_WEIGHT_GRADS)
        dz16 = None
        for i, (w, wb, sv) in enumerate(zip(reversed(layers), reversed(back), reversed(ctx.saved))):
            dz, grads, dz16 = layer_backward(dz, sv, B, L, w, wb, sc, branch, dz16, want_dx16=i + 1 < len(layers))
            per_layer.append(grads)
        del dz16
        branch.join()
        ctx.saved = None
        flat = [g for grads in reversed(per_layer) for g in grads]
        out = []
        for p, g, need in zip(ctx.params, flat, ctx.needs_input_grad[2:]):
            out.append(None if (p is None or not need) else g.to(p.dtype).view(p.shape))
        dx = dz.view(B, L, D).to(ctx.in_dtype) if ctx.needs_input_grad[0] else None
        return (dx, None, *out)


def layer_forward(x: torch.Tensor, w: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None,
                  f16_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One Dinov2Layer on the fp32 residual stream ``x`` (B, L, 768).

    ``out`` None: ``x`` is updated in place.  Otherwise ``x`` is only read and the new residual stream
    is written to ``out`` (the first residual GEMM goes out of place, which is how the caller's tokens
    stay untouched without a copy).  ``f16_out`` (B, L, 768) fp16: the layer's result is written ONLY
    there by the last GEMM's epilogue (the tokens reach the similarity
    kernel in fp16, half the bytes); the fp32 stream then holds the intermediate state ``h``."""
    B, L, D = x.shape
    x2 = x.view(B * L, D)
    y2 = x2 if out is None else out.view(B * L, D)
    h = ops.ln_rows(x2, w["g1"], w["b1"], w["eps1"])
    qkv = ops.linear(h, w["wqkv"], w["bqkv"], "bias")
    a = ops.attention(qkv.view(B, L, 3 * D), w["heads"])
    ops.linear(a.view(B * L, D), w["wo"], w["bo"], "residual", scale=w["ls1"], residual=x2, out=y2)
    h = ops.ln_rows(y2, w["g2"], w["b2"], w["eps2"])
    g = ops.linear(h, w["w1"], w["bf1"], "gelu")
    if f16_out is not None:
        ops.linear(g, w["w2"], w["bf2"], "residual_f16", scale=w["ls2"], residual=y2, out=f16_out.view(B * L, D))
        return f16_out
    ops.linear(g, w["w2"], w["bf2"], "residual", scale=w["ls2"], residual=y2, out=y2)
    return x if out is None else out


class AlignTransformer(nn.Module):
    """Drop-in for align_transformers.py:23-45 (inference forward on the B200 kernels)."""

    def __init__(self, transformer_layers: Optional[nn.Module] = None, layer_norm: Optional[nn.LayerNorm] = None):
        super().__init__()
        self.transformer_layers = transformer_layers     # transformers Dinov2Encoder (weights only)
        self.layer_norm = layer_norm
        self._packed: Optional[List[Dict[str, torch.Tensor]]] = None
        self._packed_stamp = None
        self._packed_bwd: Optional[List[Dict[str, torch.Tensor]]] = None
        self._packed_bwd_stamp = None
        self.kernel_backward = True      # False: gradients through the stock HF modules (comparisons only)

    def refresh(self) -> None:
        """Drop the packed fp16 operands.  Not normally needed: in-place weight changes (load_state_dict,
        optimizer steps) are detected through the parameters' version counters."""
        self._packed = None
        self._packed_bwd = None

    def _stamp(self):
        # (identity, in-place version) of every parameter: load_state_dict / optimizer steps bump versions
        # data_ptr catches `p.data = other` swaps (EMA, sharded optimizers) that leave the version untouched;
        # `p.data.copy_()` into the same storage is invisible to both: call refresh() after such writes
        return tuple((id(p), p._version, p.data_ptr()) for p in self.parameters())

    def _weights(self, device) -> List[Dict[str, torch.Tensor]]:
        stamp = self._stamp()
        stale = self._packed is None or self._packed_stamp != stamp or (
            self._packed and self._packed[0]["wo"].device != device)
        if stale:
            layers = [] if self.transformer_layers is None else list(self.transformer_layers.layer)
            self._packed = [pack_layer(l, device) for l in layers]
            self._packed_stamp = stamp
        return self._packed

    def _weights_bwd(self, device) -> List[Dict[str, torch.Tensor]]:
        stamp = self._stamp()
        if self._packed_bwd is None or self._packed_bwd_stamp != stamp or (
                self._packed_bwd and self._packed_bwd[0]["wo_t"].device != device):
            layers = [] if self.transformer_layers is None else list(self.transformer_layers.layer)
            self._packed_bwd = [pack_layer_bwd(l, device) for l in layers]
            self._packed_bwd_stamp = stamp
        return self._packed_bwd

    def train_forward(self, vision_tokens: torch.Tensor) -> torch.Tensor:
        """Forward under autograd on the kernels (``_AlignFn``); the optional final ``layer_norm``
        (use_layer_norm=True, not the released configuration) stays a torch module on top."""
        if not vision_tokens.is_cuda:
            raise RzError("radzero_b200 ops run on CUDA tensors only (there is no CPU fallback)")
        if vision_tokens.dim() != 3 or vision_tokens.shape[-1] != ops.HIDDEN:
            raise RzError("vision tokens must be (B, L, 768)")
        x = vision_tokens
        if self.transformer_layers is not None and len(self.transformer_layers.layer):
            params = [p for l in self.transformer_layers.layer for p in layer_params(l)]
            x = _AlignFn.apply(x, self, *params)
        if self.layer_norm is not None:
            x = self.layer_norm(x)
        return x

    def stock_forward(self, vision_tokens: torch.Tensor) -> torch.Tensor:
        """The reference's own forward (align_transformers.py:37-45) through the stock HF modules
        (``kernel_backward = False``: the comparison arm of the backward tests and benches)."""
        x = vision_tokens
        if self.transformer_layers is not None:
            x = self.transformer_layers(x)["last_hidden_state"]
        if self.layer_norm is not None:
            x = self.layer_norm(x)
        return x

    def _needs_grad(self, vision_tokens: torch.Tensor) -> bool:
        # a module in TRAIN mode under autograd takes the stock path; in eval mode (the state every
        # inference script of the reference puts the model in) the kernels run, so a missing
        # torch.no_grad() cannot silently select the slow path -- unless the INPUT itself asks for a
        # gradient (saliency / Grad-CAM in eval mode): both cases run `train_forward`
        if not torch.is_grad_enabled():
            return False
        if vision_tokens.requires_grad:
            return True
        return self.training and any(p.requires_grad for p in self.parameters())

    def forward(self, vision_tokens: torch.Tensor, inplace: bool = False,
                handoff_f16: bool = False) -> torch.Tensor:
        """``inplace=True`` lets the kernels update the caller's fp32 token buffer (no copy).  The result
        comes back in the input's dtype, as the reference's module does (bf16 / fp16 models).
        ``handoff_f16=True`` (inference, when the consumer is the similarity kernel): the last layer's
        tokens are emitted by its fc2 epilogue as fp16 and returned as such -- the fp32 copy is never
        written and the similarity kernel reads half the bytes."""
        if self._needs_grad(vision_tokens):
            return self.train_forward(vision_tokens) if self.kernel_backward else self.stock_forward(vision_tokens)
        with torch.no_grad():
            out = self._forward_kernels(vision_tokens, inplace, handoff_f16)
        if handoff_f16 and out.dtype == torch.float16:
            return out
        return out if out.dtype == vision_tokens.dtype else out.to(vision_tokens.dtype)

    def _forward_kernels(self, vision_tokens: torch.Tensor, inplace: bool, handoff_f16: bool = False) -> torch.Tensor:
        if not vision_tokens.is_cuda:
            raise RzError("radzero_b200 ops run on CUDA tensors only (there is no CPU fallback)")
        if vision_tokens.dim() != 3 or vision_tokens.shape[-1] != ops.HIDDEN:
            raise RzError("vision tokens must be (B, L, 768)")
        x = vision_tokens.detach()
        layers = self._weights(x.device)
        fresh = None
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()                  # already a private copy
        elif not inplace and layers:
            fresh = torch.empty_like(x)                           # first layer writes here: no copy of the input
        elif not inplace:
            x = x.clone()
        f16 = None
        if handoff_f16 and layers and self.layer_norm is None:
            f16 = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        for i, w in enumerate(layers):
            last16 = f16 if i == len(layers) - 1 else None
            if i == 0 and fresh is not None:
                x = layer_forward(x, w, out=fresh, f16_out=last16)
            else:
                x = layer_forward(x, w, f16_out=last16)
        if f16 is not None:
            return f16
        if self.layer_norm is not None:
            # use_layer_norm=True (not the released configuration): one more row LayerNorm, fp32 out
            g, b = self.layer_norm.weight.detach(), self.layer_norm.bias.detach()
            if abs(self.layer_norm.eps - 1e-5) > 1e-12:
                raise RzError("AlignTransformer.layer_norm: only the nn.LayerNorm default eps is supported")
            _, x32, _ = ops.prep_rows(x, g, b, want_f16=False, want_f32=True, l2=False)
            x = x32.view(x.shape)
        return x
