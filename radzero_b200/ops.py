"""Thin torch-facing wrappers over the C ABI (device pointers + current CUDA stream).

PyTorch is plumbing here: it owns device memory and streams.  Every function takes CUDA
tensors, allocates its outputs, calls librz_b200.so on torch's current stream and returns.
No CPU path exists -- CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import RzError

HIDDEN = 768
_DTYPES = {torch.float32: _lib.RZ_F32, torch.bfloat16: _lib.RZ_BF16, torch.float16: _lib.RZ_F16}


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _guard(fn):
    """Run ``fn`` with the CUDA device of its tensor arguments current (kernels, TMA descriptors and the
    stream all belong to that device, whatever ``torch.cuda.current_device()`` was) and refuse tensors
    that live on different devices."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RzError(f"{fn.__name__}: tensors on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RzError("radzero_b200 ops run on CUDA tensors only (there is no CPU fallback)")


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------- K1 + K2
@_guard
def prep_rows(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], *,
              rows_per_group: Optional[int] = None, rows_per_group_padded: Optional[int] = None,
              want_f16: bool = True, want_f32: bool = False, want_stats: bool = False,
              l2: bool = True):
    """LayerNorm + L2-normalise rows of ``x`` (..., 768).  Returns (f16, f32, stats)."""
    _need_cuda(x, gamma, beta)
    if x.shape[-1] != HIDDEN:
        raise RzError(f"hidden size must be {HIDDEN}, got {x.shape[-1]}")
    if x.dtype not in _DTYPES:
        raise RzError(f"unsupported input dtype {x.dtype}")
    x2 = _contig(x).view(-1, HIDDEN)
    rows = x2.shape[0]
    rpg = rows_per_group or max(rows, 1)
    rpp = rows_per_group_padded or rpg
    g = _contig(gamma.float()) if gamma is not None else None
    b = _contig(beta.float()) if beta is not None else None
    groups = rows // rpg if rows else 0
    f16 = torch.empty((groups * rpp, HIDDEN), dtype=torch.float16, device=x.device) if want_f16 else None
    f32 = torch.empty((rows, HIDDEN), dtype=torch.float32, device=x.device) if want_f32 else None
    st = torch.empty((rows, 3), dtype=torch.float32, device=x.device) if want_stats else None
    rc = _lib.load().rz_prep_rows(_p(x2), _DTYPES[x.dtype], _p(g), _p(b), rows, rpg, rpp,
                                  _p(f16), _p(f32), _p(st), 1 if l2 else 0, _stream())
    _lib.check(rc, "rz_prep_rows")
    return f16, f32, st


# ----------------------------------------------------------------------------- K3 - K6
TOKEN_TILE = 128          # operands are padded to the GEMM token tile (1370 -> 1408)
LARGE_N_THRESHOLD = 64    # above this many prompts the three-pass GEMM forward is used


def padded_tokens(tokens: int) -> int:
    return (tokens + TOKEN_TILE - 1) // TOKEN_TILE * TOKEN_TILE


def _log_tau_ptr(t):
    """A device fp32 scalar holding a log-temperature (or None)."""
    if t is None:
        return None
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.numel() == 1):
        raise RzError("log-temperature must be a 1-element fp32 CUDA tensor")
    return t.detach()


def _sim_outputs(B, N, tokens, dev, want_scores, drop_cls, want_z, z_out, z_image_major):
    drop = 1 if drop_cls else 0
    scores = torch.empty((B, N, tokens - drop), dtype=torch.float32, device=dev) if want_scores else None
    z = None
    if want_z:
        zshape = (B, N) if z_image_major else (N, B)
        z = z_out if z_out is not None else torch.empty(zshape, dtype=torch.float32, device=dev)
        if z.dtype != torch.float32 or tuple(z.shape) != zshape:
            raise RzError(f"z_out must be fp32 {zshape}")
    zs_text, zs_img = (0, 0) if z is None else ((z.stride(1), z.stride(0)) if z_image_major
                                                else (z.stride(0), z.stride(1)))
    return drop, scores, z, zs_text, zs_img


@_guard
def sim_fwd(k_f16: torch.Tensor, q_f16: torch.Tensor, tokens: int, scale: float, *,
            want_scores: bool = False, drop_cls: bool = True, want_z: bool = True,
            want_stats: bool = False, want_pooled: bool = False,
            q_inv_norm: Optional[torch.Tensor] = None, z_out: Optional[torch.Tensor] = None,
            z_scale: float = 1.0, z_sigmoid: bool = False, z_image_major: bool = False,
            log_tau_scale: Optional[torch.Tensor] = None, log_tau_z: Optional[torch.Tensor] = None):
    """Fused similarity forward.  k_f16 (B, Lp, 768) fp16, q_f16 (N, 768) fp16.

    Returns dict(scores (B, N, L - drop) | None, z (N, B) | None, lse, onorm (B, N) | None,
    pooled (B, N, 768) fp16 | None).  ``z_out`` lets the caller place Z directly into a
    column block of a larger (N, ld) matrix; ``z_image_major`` makes z (B, N) instead, and
    ``z_scale`` / ``z_sigmoid`` fuse the ``/ tau`` and sigmoid of compute_logits /
    similarity_prob into the epilogue.  ``log_tau_scale`` / ``log_tau_z`` (1-element fp32
    CUDA tensors, e.g. the ``loss_temperature`` parameter) make the kernel read the
    temperature on the device: scale = exp(-log_tau), no host sync.
    """
    _need_cuda(k_f16, q_f16, q_inv_norm, z_out)
    if k_f16.dtype != torch.float16 or q_f16.dtype != torch.float16:
        raise RzError("sim_fwd operands must be fp16 rows from prep_rows")
    if k_f16.dim() != 3 or k_f16.shape[-1] != HIDDEN or not k_f16.is_contiguous():
        raise RzError("k_f16 must be contiguous (B, Lp, 768)")
    if q_f16.dim() != 2 or q_f16.shape[-1] != HIDDEN or not q_f16.is_contiguous():
        raise RzError("q_f16 must be contiguous (N, 768)")
    B, Lp, _ = k_f16.shape
    N = q_f16.shape[0]
    dev = k_f16.device
    large = N > LARGE_N_THRESHOLD and Lp % 128 == 0
    drop, scores, z, zs_text, zs_img = _sim_outputs(B, N, tokens, dev, want_scores and not large, drop_cls,
                                                    want_z, z_out, z_image_major)
    lse = torch.empty((B, N), dtype=torch.float32, device=dev) if want_stats else None
    onorm = torch.empty((B, N), dtype=torch.float32, device=dev) if want_stats else None
    pooled = torch.empty((B, N, HIDDEN), dtype=torch.float16, device=dev) if want_pooled else None
    qin = _contig(q_inv_norm.float()) if q_inv_norm is not None else None
    lts, ltz = _log_tau_ptr(log_tau_scale), _log_tau_ptr(log_tau_z)
    lib = _lib.load()
    if N > LARGE_N_THRESHOLD and Lp % 128 == 0:
        # more prompts than one SM's TMEM can pool (64 x 768 fp32): two full-rate GEMM passes.
        # The similarity map leaves the kernel through TMA stores: rows of `pitch` floats (a
        # multiple of 32 = 128 B) holding all tokens, CLS at column 0; the caller gets a view.
        store = None
        if want_scores:
            pitch = (int(tokens) + 31) // 32 * 32
            store = torch.empty((B, N, pitch), dtype=torch.float32, device=dev)
            scores = store[:, :, drop:int(tokens)]
        nbytes = int(lib.rz_sim_fwd_large_workspace_bytes(B, N, Lp))
        pt = mref = lsum = None
        if want_pooled:     # the training path keeps P~ and its row statistics for rz_sim_bwd
            pt = torch.empty((B, N, Lp), dtype=torch.float16, device=dev)
            mref = torch.empty((B, N), dtype=torch.float32, device=dev)
            lsum = torch.empty((B, N), dtype=torch.float32, device=dev)
            nbytes -= B * N * Lp * 2
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        rc = lib.rz_sim_fwd_large(
            _p(k_f16), B, int(tokens), Lp, _p(q_f16), N, float(scale), _p(lts), _p(qin),
            _p(store), store.stride(0) if store is not None else 0,
            store.stride(1) if store is not None else 0, 0,
            _p(z), zs_text, zs_img, float(z_scale), _p(ltz), 1 if z_sigmoid else 0,
            _p(lse), _p(onorm), _p(pooled), _p(pt), _p(mref), _p(lsum), 1 if want_z else 0, _p(ws),
            C.c_size_t(ws.numel()), _stream())
        _lib.check(rc, "rz_sim_fwd_large")
        return dict(scores=scores, z=z, lse=lse, onorm=onorm, pooled=pooled, p=pt, mref=mref, lsum=lsum)
    rc = lib.rz_sim_fwd(
        _p(k_f16), B, int(tokens), Lp, _p(q_f16), N, float(scale), _p(lts), _p(qin),
        _p(scores), scores.stride(0) if scores is not None else 0,
        scores.stride(1) if scores is not None else 0, drop,
        _p(z), zs_text, zs_img, float(z_scale), _p(ltz), 1 if z_sigmoid else 0,
        _p(lse), _p(onorm), _p(pooled), _stream())
    _lib.check(rc, "rz_sim_fwd")
    return dict(scores=scores, z=z, lse=lse, onorm=onorm, pooled=pooled, p=None, mref=None, lsum=None)


FUSED_PREP_MAX_TEXT = 16
# Small prompt sets (N <= 16) go through the single-kernel rz_sim_fwd_tokens: the raw tokens cross
# HBM once (no fp16 copy is written) and the work is balanced over all SMs for any batch size.
USE_FUSED_PREP = True


@_guard
def sim_fwd_tokens(tokens_raw: torch.Tensor, gamma: Optional[torch.Tensor],
                   beta: Optional[torch.Tensor], q_f16: Optional[torch.Tensor], scale: float, *, l2: bool = True,
                   text_raw: Optional[torch.Tensor] = None,
                   want_scores: bool = False, drop_cls: bool = True, want_z: bool = True,
                   q_inv_norm: Optional[torch.Tensor] = None, z_out: Optional[torch.Tensor] = None,
                   z_scale: float = 1.0, z_sigmoid: bool = False, z_image_major: bool = False,
                   log_tau_scale: Optional[torch.Tensor] = None,
                   log_tau_z: Optional[torch.Tensor] = None):
    """Small-prompt-set forward straight from the RAW tokens (B, L, 768) fp32/bf16/fp16:
    LayerNorm + L2 + the whole similarity path in ONE kernel (N <= 16).  ``text_raw`` (N, 768) fp32 instead
    of ``q_f16``: the prompts' own LayerNorm + L2 also run inside the kernel (no prep launch)."""
    _need_cuda(tokens_raw, gamma, beta, q_f16, q_inv_norm, z_out, text_raw)
    if tokens_raw.dtype not in _DTYPES:
        raise RzError(f"unsupported token dtype {tokens_raw.dtype}")
    if tokens_raw.dim() != 3 or tokens_raw.shape[-1] != HIDDEN:
        raise RzError("tokens must be (B, L, 768)")
    x = _contig(tokens_raw)
    B, L, _ = x.shape
    if (q_f16 is None) == (text_raw is None):
        raise RzError("give either q_f16 (prepared rows) or text_raw (fp32 rows before LayerNorm + L2)")
    txt = None
    if text_raw is not None:
        if q_inv_norm is not None:
            raise RzError("text_raw is for sim_op 'cos' (no q_inv_norm)")
        if text_raw.dtype != torch.float32 or text_raw.dim() != 2 or text_raw.shape[1] != HIDDEN:
            raise RzError("text_raw must be fp32 (N, 768)")
        txt = _contig(text_raw.detach())
    N = (q_f16 if q_f16 is not None else txt).shape[0]
    if N > FUSED_PREP_MAX_TEXT:
        raise RzError(f"sim_fwd_tokens handles at most {FUSED_PREP_MAX_TEXT} prompts")
    if q_f16 is not None and (q_f16.dtype != torch.float16 or not q_f16.is_contiguous()):
        raise RzError("q_f16 must be contiguous fp16 (N, 768)")
    dev = x.device
    drop, scores, z, zs_text, zs_img = _sim_outputs(B, N, L, dev, want_scores, drop_cls, want_z,
                                                    z_out, z_image_major)
    g = _contig(gamma.detach().float()) if gamma is not None else None
    b = _contig(beta.detach().float()) if beta is not None else None
    qin = _contig(q_inv_norm.float()) if q_inv_norm is not None else None
    lts, ltz = _log_tau_ptr(log_tau_scale), _log_tau_ptr(log_tau_z)
    lib = _lib.load()
    nbytes = int(lib.rz_sim_fwd_tokens_workspace_bytes(B, N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = lib.rz_sim_fwd_tokens(
        _p(x), _DTYPES[x.dtype], _p(g), _p(b), 1 if l2 else 0, B, L, _p(q_f16), N, float(scale),
        _p(lts), _p(qin), _p(scores), scores.stride(0) if scores is not None else 0,
        scores.stride(1) if scores is not None else 0, drop, _p(z), zs_text, zs_img,
        float(z_scale), _p(ltz), 1 if z_sigmoid else 0, _p(txt), _p(ws), C.c_size_t(nbytes), _stream())
    _lib.check(rc, "rz_sim_fwd_tokens")
    return dict(scores=scores, z=z)


# ----------------------------------------------------------------------------- backward
BWD_TOKEN_TILE = 128


def padded_tokens_bwd(tokens: int) -> int:
    return (tokens + BWD_TOKEN_TILE - 1) // BWD_TOKEN_TILE * BWD_TOKEN_TILE


@_guard
def sim_bwd(k_f16: torch.Tensor, q_f16: torch.Tensor, tokens: int, inv_tau: float, z: torch.Tensor,
            dz: torch.Tensor, lse: Optional[torch.Tensor], onorm: torch.Tensor, pooled: torch.Tensor, *,
            log_tau: Optional[torch.Tensor] = None, p: Optional[torch.Tensor] = None,
            mref: Optional[torch.Tensor] = None, lsum: Optional[torch.Tensor] = None,
            q_inv_norm: Optional[torch.Tensor] = None):
    """Closed-form backward of the fused similarity.  Returns (dq (N,768), dk (B,Lp,768), dlog_tau (1,)).
    ``p`` / ``mref`` / ``lsum`` (kept by the large-N forward) select the single-GEMM coefficient pass.
    ``q_inv_norm`` (N,) = 1/|q_n|: sim_op "dot" -- operands without L2 normalisation, ``inv_tau`` = 1/sqrt(768);
    the radial term of dq is added here."""
    _need_cuda(k_f16, q_f16, z, dz, lse, onorm, pooled, q_inv_norm)
    B, Lp, _ = k_f16.shape
    N = q_f16.shape[0]
    if Lp % BWD_TOKEN_TILE:
        raise RzError("training path needs tokens padded to a multiple of 128")
    if z.shape != (N, B) or dz.shape != (N, B) or z.stride(1) != 1 or dz.stride() != z.stride():
        raise RzError("z / dz must be fp32 (N, B) with identical row pitch")
    dev = k_f16.device
    lib = _lib.load()
    nbytes = int(lib.rz_sim_bwd_workspace_bytes(B, N, Lp))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dq = torch.empty((N, HIDDEN), dtype=torch.float32, device=dev)
    dk = torch.empty((B, Lp, HIDDEN), dtype=torch.float32, device=dev)
    dlt = torch.empty(1, dtype=torch.float32, device=dev)
    qin = _contig(q_inv_norm.float()) if q_inv_norm is not None else None
    rc = lib.rz_sim_bwd(_p(k_f16), B, int(tokens), Lp, _p(q_f16), N, float(inv_tau),
                        _p(_log_tau_ptr(log_tau)), _p(z), _p(dz), z.stride(0), _p(lse), _p(onorm),
                        _p(pooled), _p(p), _p(mref), _p(lsum), _p(qin), _p(dq), _p(dk), _p(dlt), _p(ws),
                        C.c_size_t(nbytes), _stream())
    _lib.check(rc, "rz_sim_bwd")
    if qin is not None:
        # Z = <q/|q|, .>: the derivative of the query normalisation (losses.py:226), a rank-one term per prompt
        radial = (dz[:, :B] * z[:, :B]).sum(dim=1) * qin * qin
        dq.addcmul_(q_f16.float(), radial.unsqueeze(1), value=-1.0)
    return dq, dk, dlt


@_guard
def prep_rows_bwd(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                  dnorm: torch.Tensor, *, rows_per_group: Optional[int] = None,
                  rows_per_group_padded: Optional[int] = None, l2: bool = True,
                  dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None,
                  accumulate: bool = False, native_dx: bool = False):
    """Backward of prep_rows: returns (dx (rows, 768), dgamma, dbeta).  dx is fp32, or -- with
    ``native_dx`` and a 16-bit input -- written directly in the input's dtype."""
    _need_cuda(x, gamma, beta, dnorm, dgamma, dbeta)
    x2 = _contig(x).view(-1, HIDDEN)
    rows = x2.shape[0]
    rpg = rows_per_group or max(rows, 1)
    rpp = rows_per_group_padded or rpg
    if dnorm.dtype != torch.float32 or not dnorm.is_contiguous():
        raise RzError("dnorm must be contiguous fp32")
    if dnorm.numel() != (rows // rpg) * rpp * HIDDEN:
        raise RzError("dnorm does not match the padded group layout")
    dev = x.device
    lib = _lib.load()
    native = bool(native_dx) and x.dtype in (torch.bfloat16, torch.float16)
    dx = torch.empty((rows, HIDDEN), dtype=x.dtype if native else torch.float32, device=dev)
    g = _contig(gamma.detach().float()) if gamma is not None else None
    b = _contig(beta.detach().float()) if beta is not None else None
    part = None
    if g is not None:
        part = torch.empty((int(lib.rz_prep_rows_bwd_blocks(rows)), 2, HIDDEN), dtype=torch.float32, device=dev)
        if dgamma is None:
            dgamma = torch.zeros(HIDDEN, dtype=torch.float32, device=dev)
            dbeta = torch.zeros(HIDDEN, dtype=torch.float32, device=dev)
            accumulate = False
    rc = lib.rz_prep_rows_bwd(_p(x2), _DTYPES[x.dtype], _p(g), _p(b), rows, rpg, rpp, _p(dnorm),
                              1 if l2 else 0, _p(dx), 1 if native else 0, _p(part), _p(dgamma), _p(dbeta),
                              1 if accumulate else 0, 1.0, _stream())
    _lib.check(rc, "rz_prep_rows_bwd")
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------- K8 + K9
def _as_maps(t: torch.Tensor) -> torch.Tensor:
    """View (..., G*G) as (maps, G*G) with unit inner stride and ONE map stride, without copying
    when the leading dims collapse (e.g. the padded-pitch similarity map of the large-N forward)."""
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.stride(-1) != 1:
        return t.reshape(-1, t.shape[-1]).contiguous()
    if t.dim() == 2:
        return t
    lead, strides = list(t.shape[:-1]), list(t.stride()[:-1])
    ok = all(strides[i] == strides[i + 1] * lead[i + 1] for i in range(len(lead) - 1))
    if not ok:
        return t.reshape(-1, t.shape[-1]).contiguous()
    maps = 1
    for d in lead:
        maps *= d
    return t.as_strided((maps, t.shape[-1]), (strides[-1], 1), t.storage_offset())


@_guard
def upsample_maps(scores: torch.Tensor, out_hw: Tuple[int, int], *, mode: int = _lib.RZ_UP_RAW,
                  interp_hw: Optional[Tuple[int, int]] = None, offset: Tuple[int, int] = (0, 0),
                  fill: float = -999.0, threshold: float = 0.5, grid: Optional[int] = None):
    """Upsample ``scores`` (maps, grid*grid) fp32 -> (maps, H, W) (or (maps, 2) for argmax, or the bit-packed
    mask (maps, H, ceil(W / 32)) int32 for RZ_UP_MASK_BITS: bit x % 32 of word x // 32)."""
    _need_cuda(scores)
    if scores.dtype != torch.float32:
        raise RzError("scores must be fp32")
    scores = _as_maps(scores)
    maps, n = scores.shape
    g = grid or int(round(n ** 0.5))
    if g * g != n:
        raise RzError(f"scores last dim {n} is not a square grid")
    H, W = int(out_hw[0]), int(out_hw[1])
    ih, iw = (H, W) if interp_hw is None else (int(interp_hw[0]), int(interp_hw[1]))
    if mode in (_lib.RZ_UP_RAW, _lib.RZ_UP_SIGMOID):
        out = torch.empty((maps, H, W), dtype=torch.float32, device=scores.device)
    elif mode == _lib.RZ_UP_MASK:
        out = torch.empty((maps, H, W), dtype=torch.uint8, device=scores.device)
    elif mode == _lib.RZ_UP_MASK_BITS:
        out = torch.empty((maps, H, (W + 31) // 32), dtype=torch.int32, device=scores.device)
    elif mode == _lib.RZ_UP_ARGMAX:
        out = torch.empty((maps, 2), dtype=torch.int64, device=scores.device)
    else:
        raise RzError(f"bad upsample mode {mode}")
    lib = _lib.load()
    step = 32768  # gridDim.y limit; chunk very large batches
    for m0 in range(0, maps, step):
        m1 = min(maps, m0 + step)
        rc = lib.rz_upsample_maps(_p(scores[m0:]), scores.stride(0), m1 - m0, g, H, W, ih, iw,
                                  int(offset[0]), int(offset[1]), float(fill), mode,
                                  float(threshold), _p(out[m0:]), _stream())
        _lib.check(rc, "rz_upsample_maps")
    return out


@_guard
def map_threshold_stats(scores: torch.Tensor, out_hw: Tuple[int, int], thresholds_logit: torch.Tensor, *,
                        gt_masks: Optional[torch.Tensor] = None,
                        interp_hw: Optional[Tuple[int, int]] = None, offset: Tuple[int, int] = (0, 0),
                        fill: float = -999.0, grid: Optional[int] = None):
    """Threshold statistics of upsampled maps without materialising them.

    Returns (hist_all, hist_gt) int64 (maps, T + 1) and max_score fp32 (maps,); see
    rz_map_threshold_stats in include/rz_b200.h."""
    _need_cuda(scores, thresholds_logit, gt_masks)
    scores = _as_maps(scores.float() if scores.dtype != torch.float32 else scores)
    maps, n = scores.shape
    g = grid or int(round(n ** 0.5))
    if g * g != n:
        raise RzError(f"scores last dim {n} is not a square grid")
    H, W = int(out_hw[0]), int(out_hw[1])
    ih, iw = (H, W) if interp_hw is None else (int(interp_hw[0]), int(interp_hw[1]))
    thr = _contig(thresholds_logit.float())
    T = thr.numel()
    if gt_masks is not None:
        if gt_masks.dtype != torch.uint8 or tuple(gt_masks.shape) != (maps, H, W) or not gt_masks.is_contiguous():
            raise RzError("gt_masks must be contiguous uint8 (maps, H, W)")
    dev = scores.device
    ha = torch.empty((maps, T + 1), dtype=torch.int32, device=dev)
    hg = torch.empty((maps, T + 1), dtype=torch.int32, device=dev)
    mx = torch.empty(maps, dtype=torch.float32, device=dev)
    lib = _lib.load()
    step = 32768
    for m0 in range(0, maps, step):
        m1 = min(maps, m0 + step)
        rc = lib.rz_map_threshold_stats(_p(scores[m0:]), scores.stride(0), m1 - m0, g, H, W, ih, iw,
                                        int(offset[0]), int(offset[1]), float(fill),
                                        _p(gt_masks[m0:]) if gt_masks is not None else None, _p(thr), T,
                                        _p(ha[m0:]), _p(hg[m0:]), _p(mx[m0:]), _stream())
        _lib.check(rc, "rz_map_threshold_stats")
    return ha.long(), hg.long(), mx


# ----------------------------------------------------------------------------- K10
def _colstate_rows(bl: int) -> int:
    return (bl + 3) // 4 * 4


@_guard
def mpnce_partials(z: torch.Tensor, group_map: torch.Tensor, col0: int, inv_tau: float = 1.0, *,
                   log_tau: Optional[torch.Tensor] = None, rowpos: Optional[torch.Tensor] = None,
                   eps: float = 1e-8, col_sum: bool = False, b_global: Optional[int] = None):
    """Launch 1 of MP-NCE on the local column block.  Returns (rowsum, pos, colneg, colpos); colneg / colpos
    are the first two rows of the (5, b4) column-state block this launch writes (the backward's column
    coefficients sit behind them and are picked up by ``mpnce_finish`` through the same buffer).
    ``log_tau`` (1-element fp32 CUDA tensor): temperature read on the device instead of ``inv_tau``.
    ``rowpos`` (2, n) fp32: caller-owned buffer for rowsum / pos (one all-reduce message).
    ``eps`` / ``col_sum`` / ``b_global`` must be what ``mpnce_finish`` is given."""
    _need_cuda(z, group_map)
    assert z.dtype == torch.float32 and z.dim() == 2 and z.stride(1) == 1
    n, bl = z.shape
    gm = group_map if (group_map.dtype == torch.int64 and group_map.is_contiguous()) \
        else _contig(group_map.to(torch.int64))
    dev = z.device
    if rowpos is None:
        rowpos = torch.empty((2, n), dtype=torch.float32, device=dev)
    rowsum, pos = rowpos[0], rowpos[1]
    col = torch.empty((5, _colstate_rows(bl)), dtype=torch.float32, device=dev)
    lib = _lib.load()
    scratch = torch.empty(int(lib.rz_mpnce_partials_scratch_floats(n, bl)), dtype=torch.float32, device=dev)
    rc = lib.rz_mpnce_partials(_p(z), z.stride(0), n, bl, int(b_global if b_global is not None else bl), _p(gm),
                               int(col0), float(inv_tau), _p(_log_tau_ptr(log_tau)), float(eps), int(col_sum),
                               _p(rowsum), _p(pos), _p(col), _p(scratch), _stream())
    _lib.check(rc, "rz_mpnce_partials")
    return rowsum, pos, col[0, :bl], col[1, :bl]


@_guard
def mpnce_finish(z: torch.Tensor, group_map: torch.Tensor, col0: int, b_global: int, inv_tau: float,
                 rowsum, pos, colneg, colpos, *, eps: float = 1e-8, row_sum: bool = False,
                 col_sum: bool = False, want_dz: bool = True, log_tau: Optional[torch.Tensor] = None):
    """Launch 2: returns (loss_terms[4], dz or None) for the local column block.  ``colneg`` / ``colpos`` must
    be the views ``mpnce_partials`` returned (rows 0 and 1 of its column-state block)."""
    _need_cuda(z, group_map, rowsum, pos, colneg, colpos)
    n, bl = z.shape
    b4 = _colstate_rows(bl)
    if colpos.data_ptr() != colneg.data_ptr() + 4 * b4 or colneg.dtype != torch.float32:
        raise RzError("colneg / colpos must be the column-state views returned by mpnce_partials")
    gm = group_map if (group_map.dtype == torch.int64 and group_map.is_contiguous()) \
        else _contig(group_map.to(torch.int64))
    dev = z.device
    dz = torch.empty_like(z) if want_dz else None
    if dz is not None:
        assert dz.stride(0) == z.stride(0)
    terms = torch.empty(4, dtype=torch.float32, device=dev)
    lib = _lib.load()
    scratch = torch.empty(int(lib.rz_mpnce_finish_scratch_floats(n, bl, int(b_global))), dtype=torch.float32,
                          device=dev)
    rc = lib.rz_mpnce_finish(_p(z), z.stride(0), n, bl, int(b_global), _p(gm), int(col0),
                             float(inv_tau), _p(_log_tau_ptr(log_tau)), float(eps), int(row_sum),
                             int(col_sum), _p(rowsum), _p(pos), _p(colneg), _p(scratch),
                             _p(dz), _p(terms), _stream())
    _lib.check(rc, "rz_mpnce_finish")
    return terms, dz


# ----------------------------------------------------------------------------- diagnostics
@_guard
def umma_probe(a_image: torch.Tensor, b_image: torch.Tensor, a_desc: int, b_desc: int,
               a_step_bytes: int, b_step_bytes: int, k_steps: int, idesc: int,
               d_tmem_offset: int = 0, ncols: int = 32) -> torch.Tensor:
    """Run the tcgen05 probe; a_image / b_image are uint8 CUDA tensors (smem byte images)."""
    _need_cuda(a_image, b_image)
    out = torch.empty((128, ncols), dtype=torch.float32, device=a_image.device)
    rc = _lib.load().rz_umma_probe(_p(a_image), a_image.numel(), _p(b_image), b_image.numel(),
                                   C.c_ulonglong(a_desc), C.c_ulonglong(b_desc), a_step_bytes,
                                   b_step_bytes, k_steps, C.c_uint(idesc), C.c_uint(d_tmem_offset),
                                   ncols, _p(out), _stream())
    _lib.check(rc, "rz_umma_probe")
    return out


# ----------------------------------------------------------------------------- A0 - A2
@_guard
def ln_rows(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    """nn.LayerNorm(768, eps) of the rows of ``x`` (..., 768) -> fp16 rows (a GEMM operand)."""
    _need_cuda(x, gamma, beta)
    if x.shape[-1] != HIDDEN or x.dtype not in _DTYPES:
        raise RzError(f"ln_rows needs (..., {HIDDEN}) fp32/bf16/fp16 rows, got {tuple(x.shape)} {x.dtype}")
    x2 = _contig(x).view(-1, HIDDEN)
    out = torch.empty(x2.shape, dtype=torch.float16, device=x.device)
    if x2.shape[0] == 0:
        return out.view(x.shape)
    rc = _lib.load().rz_ln_rows(_p(x2), _DTYPES[x.dtype], _p(_contig(gamma.float())), _p(_contig(beta.float())),
                                float(eps), x2.shape[0], _p(out), _stream())
    _lib.check(rc, "rz_ln_rows")
    return out.view(x.shape)


@_guard
def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: str = "bias", *,
           scale: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``epilogue(a @ w.T + bias)`` on the tcgen05 GEMM skeleton.

    a (M, K) fp16, w (N, K) fp16 (nn.Linear.weight layout), bias fp32 (N,).  ``epilogue``:
    "bias" / "gelu" -> fp16 (M, N); "residual" -> fp32 ``residual + scale * (a @ w.T + bias)``
    (``out`` may be ``residual`` itself: in-place update of the residual stream); "residual_f16" -> the
    same value written only as fp16 (the last layer's hand-off to the similarity kernel).
    """
    _need_cuda(a, w, bias, scale, residual, out)
    if a.dtype != torch.float16 or w.dtype != torch.float16 or a.dim() != 2 or w.dim() != 2:
        raise RzError("linear operands must be 2-d fp16")
    if not (a.is_contiguous() and w.is_contiguous()) or a.shape[1] != w.shape[1]:
        raise RzError("linear operands must be contiguous (M, K) and (N, K)")
    m, k = a.shape
    n = w.shape[0]
    ep = {"bias": _lib.RZ_LIN_BIAS, "gelu": _lib.RZ_LIN_GELU, "residual": _lib.RZ_LIN_RESIDUAL,
          "residual_f16": _lib.RZ_LIN_RESIDUAL_F16}[epilogue]
    for v in (bias, scale):
        if v is not None and (v.dtype != torch.float32 or v.numel() != n or not v.is_contiguous()):
            raise RzError("bias / scale must be contiguous fp32 (N,)")
    if ep in (_lib.RZ_LIN_RESIDUAL, _lib.RZ_LIN_RESIDUAL_F16):
        if residual is None or residual.dtype != torch.float32 or tuple(residual.shape) != (m, n) \
                or not residual.is_contiguous():
            raise RzError("residual must be contiguous fp32 (M, N)")
        odt = torch.float32 if ep == _lib.RZ_LIN_RESIDUAL else torch.float16
    else:
        odt = torch.float16
    if out is None:
        out = torch.empty((m, n), dtype=odt, device=a.device)
    elif out.dtype != odt or tuple(out.shape) != (m, n) or not out.is_contiguous():
        raise RzError(f"out must be contiguous {odt} ({m}, {n})")
    if m == 0:
        return out
    rc = _lib.load().rz_linear(_p(a), m, k, _p(w), n, _p(bias), ep, _p(scale), _p(residual), _p(out), _stream())
    _lib.check(rc, "rz_linear")
    return out


@_guard
def attention(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T) v per (image, head); qkv (B, L, 3 * heads * 64) fp16 = [q | k | v] with the
    1/sqrt(64) scale folded into q.  Returns (B, L, heads * 64) fp16."""
    _need_cuda(qkv)
    if qkv.dtype != torch.float16 or qkv.dim() != 3 or qkv.shape[-1] != 3 * heads * 64 or not qkv.is_contiguous():
        raise RzError("qkv must be contiguous fp16 (B, L, 3 * heads * 64)")
    B, L, _ = qkv.shape
    out = torch.empty((B, L, heads * 64), dtype=torch.float16, device=qkv.device)
    if B == 0 or L == 0:
        return out
    rc = _lib.load().rz_attention(_p(qkv), B, L, heads, _p(out), _stream())
    _lib.check(rc, "rz_attention")
    return out


# ----------------------------------------------------------------------------- A3 (AlignTransformer backward)
_ATTN_BWD_MAX_BH = 65535      # (image, head) pairs per rz_attention_bwd launch

def _f16c(t, what):
    if t.dtype != torch.float16 or not t.is_contiguous():
        raise RzError(f"{what} must be contiguous fp16")


@_guard
def grad_scale(grad: torch.Tensor) -> torch.Tensor:
    """Power-of-two scale of the fp16 gradient chain from max|grad| (on the device, no sync):
    returns sc fp32 with sc[0] = 2^k, sc[1] = 2^-k, sc[2] = max|grad|, sc[3] = k, sc[4:] = 2^-k."""
    _need_cuda(grad)
    if grad.dtype != torch.float32 or not grad.is_contiguous():
        raise RzError("grad must be contiguous fp32")
    lib = _lib.load()
    sc = torch.empty(int(lib.rz_grad_scale_floats()), dtype=torch.float32, device=grad.device)
    _lib.check(lib.rz_grad_scale(_p(grad), grad.numel(), _p(sc), _stream()), "rz_grad_scale")
    return sc


@_guard
def ls_cast_bwd(dy: torch.Tensor, ls: Optional[torch.Tensor], o16: Optional[torch.Tensor], sc: torch.Tensor,
                dls: Optional[torch.Tensor]) -> torch.Tensor:
    """Dinov2LayerScale backward: returns fp16(2^k ls dy) (rows, 768); ``dls`` (768,) += sum_rows dy * o16."""
    _need_cuda(dy, ls, o16, sc, dls)
    if dy.dtype != torch.float32 or not dy.is_contiguous() or dy.dim() != 2 or dy.shape[1] != HIDDEN:
        raise RzError("dy must be contiguous fp32 (rows, 768)")
    if o16 is not None:
        _f16c(o16, "o16")
    out = torch.empty(dy.shape, dtype=torch.float16, device=dy.device)
    rc = _lib.load().rz_ls_cast_bwd(_p(dy), _p(ls), _p(o16), _p(sc), dy.shape[0], _p(out), _p(dls), _stream())
    _lib.check(rc, "rz_ls_cast_bwd")
    return out


@_guard
def ls_weight_bwd(g: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], colsum: torch.Tensor,
                  ls: torch.Tensor) -> torch.Tensor:
    """Dinov2LayerScale backward from ``g = dy^T x`` (n, k) fp32 and ``colsum = sum_rows dy``: in place
    ``g -> dW = ls g``, ``colsum -> db = ls colsum``; returns ``dls`` (n,) = rowsum(w g) + bias colsum."""
    _need_cuda(g, w, bias, colsum, ls)
    for t in (g, w):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 2:
            raise RzError("g / w must be contiguous fp32 (n, k)")
    if g.shape != w.shape:
        raise RzError("g and w differ in shape")
    n, k = g.shape
    dls = torch.empty(n, dtype=torch.float32, device=g.device)
    rc = _lib.load().rz_ls_weight_bwd(_p(g), _p(w), _p(bias), _p(colsum), _p(ls), n, k, _p(dls), _stream())
    _lib.check(rc, "rz_ls_weight_bwd")
    return dls


@_guard
def transpose_pad(x16: torch.Tensor, sc: Optional[torch.Tensor] = None, colsum: Optional[torch.Tensor] = None,
                  want_out: bool = True) -> Optional[torch.Tensor]:
    """(rows, cols) fp16 -> (cols, rows padded to 64) fp16, the K-major operand of a dW GEMM (K = rows);
    ``colsum`` (cols,) += 2^-k sum_rows x16 (the bias gradient) from the same read."""
    _need_cuda(x16, sc, colsum)
    _f16c(x16, "x16")
    rows, cols = x16.shape
    rp = (rows + 63) // 64 * 64
    out = torch.empty((cols, rp), dtype=torch.float16, device=x16.device) if want_out else None
    rc = _lib.load().rz_transpose_pad(_p(x16), rows, cols, rp, _p(out), _p(colsum), _p(sc), _stream())
    _lib.check(rc, "rz_transpose_pad")
    return out


@_guard
def gelu_bwd(dg16: torch.Tensor, u16: torch.Tensor) -> torch.Tensor:
    """dg * gelu_erf'(u), fp16 in and out."""
    _need_cuda(dg16, u16)
    _f16c(dg16, "dg16")
    _f16c(u16, "u16")
    if dg16.shape != u16.shape:
        raise RzError("gelu_bwd operands differ in shape")
    out = torch.empty_like(dg16)
    _lib.check(_lib.load().rz_gelu_bwd(_p(dg16), _p(u16), dg16.numel(), _p(out), _stream()), "rz_gelu_bwd")
    return out


@_guard
def ln_rows_bwd(x: torch.Tensor, dh16: torch.Tensor, gamma: torch.Tensor, eps: float, dres: Optional[torch.Tensor],
                sc: torch.Tensor, dgamma: Optional[torch.Tensor], dbeta: Optional[torch.Tensor],
                out: Optional[torch.Tensor] = None, out16: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.LayerNorm backward: ``dres + 2^-k LN'(dh16)`` fp32 (rows, 768); dgamma / dbeta accumulated;
    ``out16`` (rows, 768) fp16, optional, receives ``fp16(2^k result)`` (the next product's operand)."""
    _need_cuda(x, dh16, gamma, dres, sc, dgamma, dbeta, out, out16)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 2 or x.shape[1] != HIDDEN:
        raise RzError("x must be contiguous fp32 (rows, 768)")
    _f16c(dh16, "dh16")
    if dres is not None and (dres.dtype != torch.float32 or not dres.is_contiguous() or dres.shape != x.shape):
        raise RzError("dres must be contiguous fp32 like x")
    if out is None:
        out = torch.empty_like(x)
    if out16 is not None and (out16.dtype != torch.float16 or not out16.is_contiguous() or out16.shape != x.shape):
        raise RzError("out16 must be contiguous fp16 like x")
    rc = _lib.load().rz_ln_rows_bwd(_p(x), _p(dh16), _p(gamma), float(eps), _p(dres), _p(sc), x.shape[0],
                                    _p(out), _p(out16), _p(dgamma), _p(dbeta), _stream())
    _lib.check(rc, "rz_ln_rows_bwd")
    return out


@_guard
def attention_bwd(qkv: torch.Tensor, out16: torch.Tensor, dout16: torch.Tensor, heads: int,
                  q_scale: float) -> torch.Tensor:
    """Backward of ``attention``: dqkv fp16 like qkv (q block times ``q_scale``)."""
    _need_cuda(qkv, out16, dout16)
    for t, n in ((qkv, "qkv"), (out16, "out16"), (dout16, "dout16")):
        _f16c(t, n)
    B, L, W = qkv.shape
    if W != 3 * heads * 64 or tuple(out16.shape) != (B, L, heads * 64) or out16.shape != dout16.shape:
        raise RzError("attention_bwd shapes: qkv (B, L, 3 * heads * 64), out / dout (B, L, heads * 64)")
    dqkv = torch.empty_like(qkv)
    if B == 0 or L == 0:
        return dqkv
    # one launch covers at most 65 535 (image, head) pairs (the grid's y extent): more images go in slices
    step = max(1, _ATTN_BWD_MAX_BH // heads)
    lp = (L + 63) // 64 * 64
    ws = torch.empty((2, min(B, step) * heads * lp), dtype=torch.float32, device=qkv.device)
    lib = _lib.load()
    for i0 in range(0, B, step):
        i1 = min(B, i0 + step)
        rc = lib.rz_attention_bwd(_p(qkv[i0:i1]), _p(out16[i0:i1]), _p(dout16[i0:i1]), i1 - i0, L, heads,
                                  float(q_scale), _p(ws[0]), _p(ws[1]), _p(dqkv[i0:i1]), _stream())
        _lib.check(rc, "rz_attention_bwd")
    return dqkv


# ----------------------------------------------------------------------------- T0
@_guard
def text_pool(hidden: torch.Tensor, attention_mask: torch.Tensor, gamma: Optional[torch.Tensor] = None,
              beta: Optional[torch.Tensor] = None, *, l2: bool = True, want_feats: bool = True,
              want_q16: bool = True):
    """Masked mean pooling of ``hidden`` (n, T, 768) + LayerNorm + L2 of the pooled rows, one launch.

    Returns ``(text_features_wo_l2_norm fp32 (n, 768) | None, q16 fp16 (n, 768) | None)``.
    """
    _need_cuda(hidden, attention_mask, gamma, beta)
    if hidden.dim() != 3 or hidden.shape[-1] != HIDDEN or hidden.dtype not in _DTYPES:
        raise RzError("hidden must be (n, T, 768) fp32/bf16/fp16")
    if tuple(attention_mask.shape) != tuple(hidden.shape[:2]):
        raise RzError("attention_mask must be (n, T)")
    h = _contig(hidden)
    m = _contig(attention_mask.to(torch.int64))
    n, t, _ = h.shape
    g = _contig(gamma.detach().float()) if gamma is not None else None
    b = _contig(beta.detach().float()) if beta is not None else None
    feats = torch.empty((n, HIDDEN), dtype=torch.float32, device=h.device) if want_feats else None
    q16 = torch.empty((n, HIDDEN), dtype=torch.float16, device=h.device) if want_q16 else None
    if n == 0:
        return feats, q16
    rc = _lib.load().rz_text_pool(_p(h), _DTYPES[h.dtype], _p(m), n, t, _p(g), _p(b), 1 if l2 else 0,
                                  _p(feats), _p(q16), _stream())
    _lib.check(rc, "rz_text_pool")
    return feats, q16


# ----------------------------------------------------------------------------- K11
_IMG_DTYPES = {torch.uint8: _lib.RZ_IMG_U8, torch.uint16: _lib.RZ_IMG_U16, torch.int16: _lib.RZ_IMG_I16,
               torch.int32: _lib.RZ_IMG_I32, torch.float32: _lib.RZ_IMG_F32}


@_guard
def preprocess_images(raw: torch.Tensor, out_hw: Tuple[int, int] = (518, 518), *, mean, std,
                      rescale_factor: float = 1.0 / 255.0, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """min-max stretch to uint8 (cv2.normalize) -> RGB -> PIL bicubic resize -> rescale -> normalise, on the
    GPU, bit-identical to the reference's host chain.  raw (B, H, W) or (B, H, W, 3) uint8 / uint16 / int16 /
    int32 / float32 on CUDA -> pixel_values (B, 3, out_h, out_w)."""
    _need_cuda(raw)
    if raw.dtype not in _IMG_DTYPES:
        raise RzError(f"unsupported raw image dtype {raw.dtype}")
    if raw.dim() == 3:
        ch = 1
    elif raw.dim() == 4 and raw.shape[-1] in (1, 3):
        ch = int(raw.shape[-1])
    else:
        raise RzError("raw images must be (B, H, W) or (B, H, W, 3)")
    if out_dtype not in _DTYPES:
        raise RzError(f"unsupported output dtype {out_dtype}")
    x = _contig(raw)
    B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    oh, ow = int(out_hw[0]), int(out_hw[1])
    out = torch.empty((B, 3, oh, ow), dtype=out_dtype, device=x.device)
    if B == 0:
        return out
    import numpy as np
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(mean, dtype=np.float32), (3,)))
    sd = np.ascontiguousarray(np.broadcast_to(np.asarray(std, dtype=np.float32), (3,)))
    lib = _lib.load()
    step = 4096                                   # images per launch (gridDim.y)
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        nbytes = int(lib.rz_preprocess_workspace_bytes(b1 - b0, H, W, ch, oh, ow))
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=x.device)
        off = (-ws.data_ptr()) % 256
        rc = lib.rz_preprocess_images(_p(x[b0:]), _IMG_DTYPES[x.dtype], b1 - b0, H, W, ch, oh, ow,
                                      m.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p),
                                      C.c_double(float(rescale_factor)), _p(out[b0:]), _DTYPES[out_dtype],
                                      C.c_void_p(ws.data_ptr() + off), C.c_size_t(nbytes), _stream())
        _lib.check(rc, "rz_preprocess_images")
    return out


# ----------------------------------------------------------------------------- group map
def group_map_from_counts(counts, first_image: int, device) -> torch.Tensor:
    """int64 (sum(counts),) tensor on ``device``: sentence j -> first_image + its image (losses.py:131-151),
    written by a kernel that receives the counts as launch parameters (no host-to-device copy)."""
    import numpy as np
    c = np.ascontiguousarray(np.asarray(counts, dtype=np.int32))
    total = int(c.sum())
    device = torch.device(device)
    if device.type != "cuda":
        raise RzError("radzero_b200 ops run on CUDA tensors only (there is no CPU fallback)")
    out = torch.empty(total, dtype=torch.int64, device=device)
    if total == 0:
        return out
    with torch.cuda.device(device):
        rc = _lib.load().rz_group_map(c.ctypes.data_as(C.c_void_p), int(c.size), int(first_image), _p(out), _stream())
    _lib.check(rc, "rz_group_map")
    return out
