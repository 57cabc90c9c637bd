"""Parity of the fused tcgen05 similarity forward (rz_sim_fwd) against the oracle and the
golden fixtures frozen from the reference.  Tolerances are BASELINE.json's: 2e-3 abs on
cosine similarities / maps (at the 1/tau scale the reference emits), 1e-3 relative on
similarity_prob; identical argmax labels."""
import math

import numpy as np
import pytest
import torch

import oracle
from radzero_b200 import ops, synthetic
from tests.golden_util import T, case_inputs, golden, split

pytestmark = pytest.mark.gpu
DEV = "cuda"
TAU = math.exp(math.log(0.07))


def _run(tok, text, gamma, beta, scale=1.0 / TAU, **kw):
    B, L, _ = tok.shape
    Lp = ops.padded_tokens(L)
    g = gamma.to(DEV) if gamma is not None else None
    b = beta.to(DEV) if beta is not None else None
    k16, _, _ = ops.prep_rows(tok.to(DEV), g, b, rows_per_group=L, rows_per_group_padded=Lp)
    q16, _, _ = ops.prep_rows(text.to(DEV), g, b)
    return ops.sim_fwd(k16.view(B, Lp, 768), q16, L, scale, **kw)


def _oracle(tok, text, gamma, beta, tau=TAU):
    tn = oracle.layer_norm_rows(text.double(), gamma.double(), beta.double())
    xn = oracle.layer_norm_rows(tok.double(), gamma.double(), beta.double())
    z, sc = oracle.similarity_logit(tn, xn, temperature=tau, need_scores=True, squeeze_quirk=False)
    q = oracle.l2_normalize_rows(tn)
    k = oracle.l2_normalize_rows(xn)
    s = sc[0]
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bnl,bld->bnd", p, k)
    return dict(z=z, scores=s, lse=torch.logsumexp(s, dim=-1), onorm=o.norm(dim=-1), pooled=o)


@pytest.mark.parametrize("B,N,L", [(2, 3, 1370), (3, 14, 1370), (2, 16, 50), (1, 1, 130),
                                   (2, 20, 200), (2, 40, 1370), (1, 64, 64), (2, 130, 333), (2, 300, 1370)])
def test_sim_fwd_vs_oracle(B, N, L):
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=100 + N)
    out = _run(tok, text, gamma, beta, want_scores=True, drop_cls=True, want_stats=True,
               want_pooled=True)
    torch.cuda.synchronize()
    want = _oracle(tok, text, gamma, beta)
    sc = out["scores"].cpu().double()
    assert sc.shape == (B, N, L - 1)
    assert (sc - want["scores"][:, :, 1:]).abs().max() < 2e-3
    z = out["z"].cpu().double()
    assert z.shape == (N, B)
    assert (z - want["z"]).abs().max() < 2e-4
    assert (out["lse"].cpu().double() - want["lse"]).abs().max() < 2e-3
    assert ((out["onorm"].cpu().double() - want["onorm"]) / want["onorm"]).abs().max() < 1e-3
    assert (out["pooled"].cpu().double() - want["pooled"]).abs().max() < 2e-3
    # similarity_prob = sigmoid(Z / tau): 1e-3 relative
    pr = torch.sigmoid(z.T / TAU)
    pw = torch.sigmoid(want["z"].T / TAU)
    assert ((pr - pw) / pw).abs().max() < 1e-3
    if N > 1:
        assert torch.equal(pr.argmax(1), pw.argmax(1))


def test_sim_fwd_keep_cls_and_z_only():
    tok, text, gamma, beta, _ = synthetic.make_inputs(2, 5, tokens_per_image=70, seed=3)
    out = _run(tok, text, gamma, beta, want_scores=True, drop_cls=False)
    want = _oracle(tok, text, gamma, beta)
    assert out["scores"].shape == (2, 5, 70)
    assert (out["scores"].cpu().double() - want["scores"]).abs().max() < 2e-3
    out2 = _run(tok, text, gamma, beta)
    assert out2["scores"] is None and out2["lse"] is None
    assert (out2["z"].cpu().double() - want["z"]).abs().max() < 2e-4


def test_sim_fwd_lazy_rescale_path():
    """Scores that keep growing along the token axis force the running maximum (and the
    pooled accumulator rescale) to fire on many tiles."""
    B, N, L = 2, 6, 700
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=77)
    g = torch.ones(768)
    bta = torch.zeros(768)
    ramp = torch.linspace(-1.0, 3.0, L).view(1, L, 1)
    tok = tok + ramp * text[:1].view(1, 1, 768)          # similarity to prompt 0 rises with l
    tok[:, :, :] = tok - tok.mean(-1, keepdim=True)
    out = _run(tok, text, g, bta, want_scores=True, want_stats=True)
    want = _oracle(tok, text, g, bta)
    assert (out["scores"].cpu().double() - want["scores"][:, :, 1:]).abs().max() < 2e-3
    assert (out["z"].cpu().double() - want["z"]).abs().max() < 2e-4
    assert (out["lse"].cpu().double() - want["lse"]).abs().max() < 2e-3


def test_sim_fwd_negated_queries_low_scores():
    """All cosines strongly negative: a fixed-shift softmax would underflow fp16 here."""
    B, N, L = 1, 4, 300
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=5)
    g, bta = torch.ones(768), torch.zeros(768)
    tok = tok * 0.05 + text[0].view(1, 1, 768)           # every token ~ parallel to prompt 0
    text = -text                                          # ... and prompt 0 points the other way
    out = _run(tok, text, g, bta, want_stats=True)
    want = _oracle(tok, text, g, bta)
    assert (out["z"].cpu().double() - want["z"]).abs().max() < 2e-4
    assert (out["lse"].cpu().double() - want["lse"]).abs().max() < 2e-3


@pytest.mark.parametrize("name", ["full", "cls14"])
def test_sim_fwd_golden(name):
    """Against tensors frozen from the reference's own RadZeroLoss.forward (fp64)."""
    tok, text, gamma, beta, log_tau, counts = case_inputs(name)
    tau = float(torch.exp(log_tau))
    out = _run(tok, text, gamma, beta, scale=1.0 / tau, want_scores=True, drop_cls=False)
    z_ref = T(f"{name}.t2i_logits").double()
    assert (out["z"].cpu().double() - z_ref).abs().max() < 2e-4
    s_ref = T(f"{name}.scores_stride7").double()
    assert (out["scores"].cpu().double()[:, :, ::7] - s_ref).abs().max() < 2e-3
    logits = out["z"].T.cpu().double() / tau
    l_ref = T(f"{name}.logits").double()
    pr, pw = torch.sigmoid(logits), torch.sigmoid(l_ref)
    assert ((pr - pw) / pw).abs().max() < 1e-3
    assert torch.equal(logits.argmax(1), l_ref.argmax(1))


@pytest.mark.parametrize("sim_op", ["cos", "dot"])
@pytest.mark.parametrize("B,N", [(3, 5), (1, 4), (2, 1), (2, 20)])
def test_per_image_queries(sim_op, B, N):
    """SimilarityLogit(..., repeat=False): queries (B, N, D) (losses.py:204-206), incl. the squeeze quirk."""
    from radzero_b200 import losses
    g = torch.Generator().manual_seed(40 + B * 7 + N)
    q = torch.randn(B, N, 768, generator=g)
    tok = torch.randn(B, 90, 768, generator=g)
    sl = losses.SimilarityLogit(sim_op)
    with torch.no_grad():
        z, sc = sl(q.to(DEV), tok.to(DEV), need_attn_weights=True, repeat=False, temperature=torch.tensor(TAU))
    want, ws = oracle.similarity_logit(q.double(), tok.double(), sim_op=sim_op, temperature=TAU, need_scores=True)
    assert tuple(z.shape) == tuple(want.shape)
    assert (z.cpu().double() - want).abs().max() < 3e-4
    assert (sc[0].cpu().double() - ws[0]).abs().max() < (2e-3 if sim_op == "cos" else 2e-2)
    with pytest.raises(NotImplementedError):
        sl(q.to(DEV).requires_grad_(True), tok.to(DEV), repeat=False, temperature=torch.tensor(TAU))


@pytest.mark.parametrize("B,N,L", [(3, 14, 1370), (1, 1, 77), (5, 16, 200), (37, 5, 333)])
@pytest.mark.parametrize("with_ln", [True, False])
def test_sim_fwd_tokens_with_the_prompts_prepared_in_the_kernel(B, N, L, with_ln):
    """text_raw: the prompts' LayerNorm + L2 run in the prologue of every CTA (no prep launch); the same
    results as preparing them with rz_prep_rows up to the fp16 rounding of a prompt element (the two
    routines sum the row statistics in a different order), and within tolerance of the oracle."""
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=500 + N)
    g = gamma.to(DEV) if with_ln else None
    b = beta.to(DEV) if with_ln else None
    q16, _, _ = ops.prep_rows(text.to(DEV), g, b)
    two = ops.sim_fwd_tokens(tok.to(DEV), g, b, q16, 1.0 / TAU, want_scores=True, drop_cls=True)
    one = ops.sim_fwd_tokens(tok.to(DEV), g, b, None, 1.0 / TAU, text_raw=text.to(DEV), want_scores=True,
                             drop_cls=True)
    assert (one["scores"] - two["scores"]).abs().max() < 5e-4          # scores are at the 1/tau = 14.3 scale
    assert (one["z"] - two["z"]).abs().max() < 2e-5
    if with_ln:
        want = _oracle(tok, text, gamma, beta)
        assert (one["scores"].cpu().double() - want["scores"][:, :, 1:]).abs().max() < 2e-3
        assert (one["z"].cpu().double() - want["z"]).abs().max() < 2e-4
    with pytest.raises(Exception):
        ops.sim_fwd_tokens(tok.to(DEV), g, b, q16, 1.0, text_raw=text.to(DEV))


def test_sim_fwd_dot_mode():
    """sim_op='dot' (the RadZeroLoss constructor default, losses.py:45, 214-215)."""
    B, N, L = 2, 7, 150
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=9)
    Lp = ops.padded_tokens(L)
    k16, _, _ = ops.prep_rows(tok.to(DEV), gamma.to(DEV), beta.to(DEV), rows_per_group=L,
                              rows_per_group_padded=Lp, l2=False)
    q16, q32, _ = ops.prep_rows(text.to(DEV), gamma.to(DEV), beta.to(DEV), l2=False, want_f32=True)
    qin = 1.0 / q16.float().norm(dim=-1)
    out = ops.sim_fwd(k16.view(B, Lp, 768), q16, L, 1.0 / math.sqrt(768), q_inv_norm=qin,
                      want_scores=True, drop_cls=False)
    tn = oracle.layer_norm_rows(text.double(), gamma.double(), beta.double())
    xn = oracle.layer_norm_rows(tok.double(), gamma.double(), beta.double())
    z, sc = oracle.similarity_logit(tn, xn, sim_op="dot", need_scores=True, squeeze_quirk=False)
    assert (out["scores"].cpu().double() - sc[0]).abs().max() < 2e-2   # |s| ~ 30: fp16 operand rounding
    assert (out["z"].cpu().double() - z).abs().max() < 3e-4


# ------------------------------------------------------------------- fused-prep variant (N <= 16)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,N,L", [(3, 14, 1370), (2, 8, 1370), (1, 1, 77), (5, 16, 200), (1, 14, 1370),
                                   (37, 5, 333)])
def test_sim_fwd_tokens_vs_oracle(dtype, B, N, L):
    """One kernel from the raw tokens.  The shapes cover the stream-K partition: one image cut into
    many pieces (B = 1), pieces + whole images in one CTA range (B = 37), short images, N = 1."""
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=300 + N)
    tok = tok.to(dtype)
    q16, _, _ = ops.prep_rows(text.to(DEV), gamma.to(DEV), beta.to(DEV))
    out = ops.sim_fwd_tokens(tok.to(DEV), gamma.to(DEV), beta.to(DEV), q16, 1.0 / TAU,
                             want_scores=True, drop_cls=True)
    want = _oracle(tok.float(), text, gamma, beta)
    assert (out["scores"].cpu().double() - want["scores"][:, :, 1:]).abs().max() < 2e-3
    assert (out["z"].cpu().double() - want["z"]).abs().max() < 2e-4
    # the two-kernel path (prep -> fp16 -> TMA) must agree with the fused one to fp16 noise
    two = _run(tok, text, gamma, beta, want_scores=True, drop_cls=True)
    # (same fp16 operands up to the LayerNorm's summation order; different tile / merge order)
    assert (two["z"] - out["z"]).abs().max() < 5e-5
    assert (two["scores"] - out["scores"]).abs().max() < 5e-4


def test_sim_fwd_tokens_device_temperature_and_prob():
    from radzero_b200 import losses
    B, N, L = 4, 14, 1370
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, tokens_per_image=L, seed=12)
    fn = losses.RadZeroLoss(sim_op="cos").to(DEV)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma)
        fn.layer_norm.bias.copy_(beta)
        fn.loss_temperature.fill_(math.log(0.05))          # not the init value
    prob = fn.similarity_prob(text.to(DEV), tok.to(DEV))
    logits, scores, z = fn.similarity(text.to(DEV), tok.to(DEV))
    want = _oracle(tok, text, gamma, beta, tau=0.05)
    pw = torch.sigmoid(want["z"].T / 0.05)
    assert prob.shape == (B, N)
    assert ((prob.cpu().double() - pw) / pw).abs().max() < 1e-3
    assert (scores.cpu().double() - want["scores"][:, :, 1:]).abs().max() < 2e-3
    assert (logits.cpu().double() - want["z"].T / 0.05).abs().max() < 2e-3
    assert torch.equal(logits.cpu().argmax(1), (want["z"].T).argmax(1))


def test_many_images_persistent_loop():
    """More work items than SMs: every CTA walks several images (barrier phases wrap)."""
    B, N, L = 330, 14, 200
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, tokens_per_image=L, seed=4)
    q16, _, _ = ops.prep_rows(text.to(DEV), gamma.to(DEV), beta.to(DEV))
    out = ops.sim_fwd_tokens(tok.to(DEV), gamma.to(DEV), beta.to(DEV), q16, 1.0 / TAU)
    two = _run(tok, text, gamma, beta)
    want = _oracle(tok, text, gamma, beta)
    assert (out["z"].cpu().double() - want["z"]).abs().max() < 2e-4
    assert (two["z"].cpu().double() - want["z"]).abs().max() < 2e-4


def test_known_answers_on_the_kernels():
    """SURVEY section 8c's self-made KATs through the CUDA path (the oracle's copies: test_oracle_golden.py)."""
    from radzero_b200 import losses
    g = torch.Generator().manual_seed(0)
    q = torch.randn(3, 768, generator=g)
    one = torch.randn(1, 1, 768, generator=g)
    sl = losses.SimilarityLogit("cos")
    tau = torch.tensor(0.07)
    with torch.no_grad():
        # (i) all tokens identical -> Z = cos(q, k);  (ii) scores within +-1/tau
        z, sc = sl(q.to(DEV), one.expand(2, 137, 768).contiguous().to(DEV), need_attn_weights=True, temperature=tau)
        cos = torch.nn.functional.cosine_similarity(q.double(), one[0].double().expand(3, 768), dim=-1)
        assert (z[:, 0].cpu().double() - cos).abs().max() < 3e-4
        assert sc[0].abs().max().item() <= 1 / 0.07 + 2e-3
        # (iv) permuting the images permutes the columns
        tok = torch.randn(5, 300, 768, generator=g)
        perm = torch.tensor([2, 0, 4, 1, 3])
        z1, _ = sl(q.to(DEV), tok.to(DEV), temperature=tau)
        z2, _ = sl(q.to(DEV), tok[perm].contiguous().to(DEV), temperature=tau)
        # (not bit for bit: the stream-K cut points and the lazy softmax references move with the image order)
        assert (z1[:, perm] - z2).abs().max().item() < 1e-4
    # (iii) one sentence per image, perfectly separated logits -> loss = log(1 + (B - 1) e^{-2/tau})
    B = 4
    zi = (2 * torch.eye(B) - 1).to(DEV)
    loss = losses.multi_positive_nce_loss(zi, torch.arange(B, device=DEV), temperature=0.07)
    assert abs(loss.item() - math.log(1 + (B - 1) * math.exp(-2 / 0.07))) < 1e-6
