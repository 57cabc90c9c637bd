"""Parity of the training path (closed-form similarity backward as tcgen05 GEMM passes,
normalisation backward, the fused contrastive autograd node) against the oracle's fp64
autograd and the gradients frozen from the reference's own RadZeroLoss (golden fixtures)."""
import math

import pytest
import torch

import oracle
from radzero_b200 import losses, ops, synthetic, training
from tests.golden_util import T, case_inputs, split

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------ prep backward
@pytest.mark.parametrize("use_ln,l2", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_prep_rows_bwd(use_ln, l2, dtype):
    rows, L, Lp = 3 * 50, 50, 128
    tok, _, gamma, beta, _ = synthetic.make_inputs(3, 1, tokens_per_image=L, seed=3)
    x = tok.reshape(rows, 768).to(dtype)
    g = torch.Generator().manual_seed(5)
    d_pad = torch.randn(3, Lp, 768, generator=g)
    gm, bt = (gamma, beta) if use_ln else (None, None)
    dx, dg, db = ops.prep_rows_bwd(x.to(DEV), gm.to(DEV) if use_ln else None, bt.to(DEV) if use_ln else None,
                                   d_pad.to(DEV), rows_per_group=L, rows_per_group_padded=Lp, l2=l2)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    y = oracle.layer_norm_rows(xd, gd, bd) if use_ln else xd
    k = oracle.l2_normalize_rows(y) if l2 else y
    (k * d_pad[:, :L].reshape(rows, 768).double()).sum().backward()
    assert _rel(dx, xd.grad) < 2e-5
    if use_ln:
        assert _rel(dg, gd.grad) < 2e-5
        assert _rel(db, bd.grad) < 2e-5
    # dx written straight in the input's 16-bit type (what the training step uses): same values,
    # rounded once
    if dtype != torch.float32:
        dx16, _, _ = ops.prep_rows_bwd(x.to(DEV), gm.to(DEV) if use_ln else None,
                                       bt.to(DEV) if use_ln else None, d_pad.to(DEV), rows_per_group=L,
                                       rows_per_group_padded=Lp, l2=l2, native_dx=True)
        assert dx16.dtype == dtype
        assert torch.equal(dx16, dx.to(dtype))


# ------------------------------------------------------------------------------ similarity backward
@pytest.mark.parametrize("use_p", [False, True])
@pytest.mark.parametrize("B,N,L", [(2, 5, 50), (3, 130, 200), (2, 70, 1370), (2, 256, 1370)])
def test_sim_bwd_vs_autograd(B, N, L, use_p):
    """use_p: the single-GEMM coefficient pass fed by the forward's unnormalised probabilities
    (large-N forward only); otherwise the pass that recomputes the scores."""
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=50 + N)
    tau = 0.07
    Lp = ops.padded_tokens_bwd(L)
    k16, _, _ = ops.prep_rows(tok.to(DEV), gamma.to(DEV), beta.to(DEV), rows_per_group=L,
                              rows_per_group_padded=Lp)
    k16 = k16.view(B, Lp, 768)
    q16, _, _ = ops.prep_rows(text.to(DEV), gamma.to(DEV), beta.to(DEV))
    fwd = ops.sim_fwd(k16, q16, L, 1.0 / tau, want_stats=True, want_pooled=True)
    g = torch.Generator().manual_seed(9)
    dz = torch.randn(N, B, generator=g) * 1e-3
    if use_p and fwd["p"] is None:
        pytest.skip("the small-N forward keeps no probabilities")
    extra = dict(p=fwd["p"], mref=fwd["mref"], lsum=fwd["lsum"]) if use_p else {}
    dq, dk, dlt = ops.sim_bwd(k16, q16, L, 1.0 / tau, fwd["z"], dz.to(DEV), fwd["lse"], fwd["onorm"],
                              fwd["pooled"], **extra)
    # checker: fp64 autograd through the same (fp16-rounded) normalised operands
    k = k16[:, :L].double().cpu().requires_grad_(True)
    q = q16.double().cpu().requires_grad_(True)
    lt = torch.tensor([math.log(tau)], dtype=torch.float64, requires_grad=True)
    s = torch.einsum("nd,bld->bnl", q, k) * torch.exp(-lt)
    p = torch.softmax(s, -1)
    o = torch.einsum("bnl,bld->bnd", p, k)
    z = (q.unsqueeze(0) * torch.nn.functional.normalize(o, dim=-1)).sum(-1).T
    (z * dz.double()).sum().backward()
    assert _rel(dq, q.grad) < 5e-3
    assert _rel(dk[:, :L], k.grad) < 5e-3
    assert float(dk[:, L:].abs().max()) == 0.0 if Lp > L else True
    # d/dlog(tau) = -sum_l dL/ds_l * s_l with sum_l dL/ds_l = 0 per row: a heavily cancelling sum, so the
    # fp16 rounding of the pooled vectors (5e-4 relative on T = <o, k>) shows up amplified; the reference
    # computes this gradient under bf16 autocast (4e-3 relative per operand element)
    assert abs(dlt.item() - lt.grad.item()) < 1e-2 * abs(lt.grad.item()) + 1e-9


# ------------------------------------------------------------------------------ the fused step
def _loss_fn(gamma, beta, log_tau, sim_op="cos"):
    fn = losses.RadZeroLoss(sim_op=sim_op).to(DEV)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma)
        fn.layer_norm.bias.copy_(beta)
        fn.loss_temperature.copy_(log_tau)
    return fn


@pytest.mark.parametrize("B,counts,L", [(3, [2, 1, 3], 50), (4, [40, 35, 50, 25], 300)])
def test_contrastive_step_vs_oracle(B, counts, L):
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts), tokens_per_image=L, seed=31)
    fn = _loss_fn(gamma, beta, log_tau)
    t = text.to(DEV).requires_grad_(True)
    x = tok.to(DEV).requires_grad_(True)
    feats = split(t, counts)
    out = fn(list(range(B)), x, lambda i: {"text_features_wo_l2_norm": feats[i]})
    loss = out["losses"]["loss"]
    loss.backward()
    gm = oracle.build_group_map(counts)
    want, grads = oracle.contrastive_step_reference(text.double(), gm, tok.double(), gamma.double(),
                                                    beta.double(), log_tau.double())
    assert abs(loss.item() - want.item()) < 1e-3 * abs(want.item())
    assert (out["t2i_logits"].cpu().double() - grads["t2i_logits"]).abs().max() < 2e-4
    assert _rel(t.grad, grads["text"]) < 1e-2
    assert _rel(x.grad, grads["vision_tokens"]) < 1e-2
    assert _rel(fn.layer_norm.weight.grad, grads["gamma"]) < 1e-2
    assert _rel(fn.layer_norm.bias.grad, grads["beta"]) < 1e-2
    assert abs(fn.loss_temperature.grad.item() - grads["log_tau"].item()) < 1e-2 * abs(grads["log_tau"].item())


@pytest.mark.parametrize("B,counts,L", [(3, [2, 1, 3], 50), (4, [40, 35, 50, 25], 300), (2, [70, 90], 257)])
def test_contrastive_step_dot_vs_oracle(B, counts, L):
    """sim_op='dot' (the RadZeroLoss constructor default, losses.py:45, 214-215) through the same kernels:
    operands without L2 normalisation, per-prompt 1/|q| folded into the pair coefficients."""
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts), tokens_per_image=L, seed=77)
    fn = _loss_fn(gamma, beta, log_tau, sim_op="dot")
    t = text.to(DEV).requires_grad_(True)
    x = tok.to(DEV).requires_grad_(True)
    feats = split(t, counts)
    out = fn(list(range(B)), x, lambda i: {"text_features_wo_l2_norm": feats[i]})
    loss = out["losses"]["loss"]
    loss.backward()
    want, grads = oracle.contrastive_step_reference(text.double(), oracle.build_group_map(counts), tok.double(),
                                                    gamma.double(), beta.double(), log_tau.double(), sim_op="dot")
    assert abs(loss.item() - want.item()) < 1e-3 * abs(want.item())
    assert (out["t2i_logits"].cpu().double() - grads["t2i_logits"]).abs().max() < 3e-4
    assert _rel(t.grad, grads["text"]) < 1e-2
    assert _rel(x.grad, grads["vision_tokens"]) < 1e-2
    assert _rel(fn.layer_norm.weight.grad, grads["gamma"]) < 1e-2
    assert _rel(fn.layer_norm.bias.grad, grads["beta"]) < 1e-2
    assert abs(fn.loss_temperature.grad.item() - grads["log_tau"].item()) < 1e-2 * abs(grads["log_tau"].item())


def test_similarity_logit_autograd_dot():
    B, N, L = 2, 6, 90
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=4)
    tn = oracle.layer_norm_rows(text, gamma, beta)
    xn = oracle.layer_norm_rows(tok, gamma, beta)
    q = tn.to(DEV).requires_grad_(True)
    k = xn.to(DEV).requires_grad_(True)
    z, _ = losses.SimilarityLogit("dot")(q, k)
    w = torch.randn(N, B, generator=torch.Generator().manual_seed(1))
    (z * w.to(DEV)).sum().backward()
    qd = tn.double().requires_grad_(True)
    kd = xn.double().requires_grad_(True)
    zd, _ = oracle.similarity_logit(qd, kd, sim_op="dot", squeeze_quirk=False)
    (zd * w.double()).sum().backward()
    assert (z.detach().cpu().double() - zd.detach()).abs().max() < 3e-4
    assert _rel(q.grad, qd.grad) < 1e-2
    assert _rel(k.grad, kd.grad) < 1e-2


def test_contrastive_step_golden_full():
    """Loss and gradients frozen from the reference's own RadZeroLoss.forward + autograd (fp64)."""
    tok, text, gamma, beta, log_tau, counts = case_inputs("full")
    fn = _loss_fn(gamma, beta, log_tau)
    t = text.to(DEV).requires_grad_(True)
    x = tok.to(DEV).requires_grad_(True)
    feats = split(t, counts)
    out = fn(list(range(len(counts))), x, lambda i: {"text_features_wo_l2_norm": feats[i]})
    out["losses"]["loss"].backward()
    want = float(T("full.loss"))
    assert abs(out["losses"]["loss"].item() - want) < 1e-3 * abs(want)
    assert _rel(t.grad, T("full.grad_text")) < 1e-2
    assert _rel(x.grad[:, ::13], T("full.grad_tokens_stride13")) < 1e-2
    assert _rel(fn.layer_norm.weight.grad, T("full.grad_gamma")) < 1e-2
    assert _rel(fn.layer_norm.bias.grad, T("full.grad_beta")) < 1e-2
    assert abs(fn.loss_temperature.grad.item() - float(T("full.grad_log_tau"))) < 1e-2 * abs(float(T("full.grad_log_tau")))


def test_similarity_logit_autograd():
    B, N, L = 2, 6, 90
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=2)
    tn = oracle.layer_norm_rows(text, gamma, beta)
    xn = oracle.layer_norm_rows(tok, gamma, beta)
    q = tn.to(DEV).requires_grad_(True)
    k = xn.to(DEV).requires_grad_(True)
    sl = losses.SimilarityLogit("cos")
    z, _ = sl(q, k, temperature=torch.tensor(0.07))
    w = torch.randn(N, B, generator=torch.Generator().manual_seed(1))
    (z * w.to(DEV)).sum().backward()
    qd = tn.double().requires_grad_(True)
    kd = xn.double().requires_grad_(True)
    zz, _ = oracle.similarity_logit(qd, kd, temperature=0.07, squeeze_quirk=False)
    (zz * w.double()).sum().backward()
    assert (z.detach().cpu().double() - zz.detach()).abs().max() < 2e-4
    assert _rel(q.grad, qd.grad) < 1e-2
    assert _rel(k.grad, kd.grad) < 1e-2


def test_multi_gpu_sharded_step_equals_single_gpu():
    """NCCL run of the image-sharded step on all visible GPUs (needs >= 2) vs one GPU."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
