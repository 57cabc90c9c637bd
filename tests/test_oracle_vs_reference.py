"""Pin the oracle against the reference's own code, executed in place.

Runs only where the reference checkout exists (the build container).  The same
comparisons are frozen into ``tests/golden/*.npz`` by ``tests/golden/make_golden.py`` so
that they also run on the GPU box (``test_oracle_golden.py``).
"""
import pytest
import torch
import torch.nn.functional as F

import oracle
from radzero_b200 import synthetic
from tests._refload import make_reference_loss, text_callback


def _split(text, counts):
    out, o = [], 0
    for c in counts:
        out.append(text[o:o + c])
        o += c
    return out


@pytest.mark.parametrize("B,counts,L", [(3, [2, 1, 3], 50), (2, [1, 2], 1370), (4, [3, 3, 3, 3], 101)])
def test_forward_and_loss_match_reference_fp64(reference_losses, B, counts, L):
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(
        B, sum(counts), tokens_per_image=L, seed=7, dtype=torch.float64)
    feats = _split(text, counts)
    ref = make_reference_loss(reference_losses, gamma, beta, log_tau)
    with torch.no_grad():
        r = ref(list(range(B)), tok, text_callback(feats), ddp_gather=False,
                need_attn_weights=True, compute_loss=True)
        o = oracle.radzero_forward(feats, tok, gamma, beta, log_tau.double(),
                                   need_attn_weights=True, compute_loss=True)
    assert r["t2i_logits"].shape == o["t2i_logits"].shape
    assert (r["t2i_logits"] - o["t2i_logits"]).abs().max() < 1e-9
    assert (r["t2i_attn_weights"][0] - o["t2i_attn_weights"][0]).abs().max() < 1e-7
    assert abs(r["losses"]["loss"].item() - o["losses"]["loss"].item()) < 1e-9


def test_forward_matches_reference_fp32(reference_losses):
    B, counts = 2, [4, 3]
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, 7, tokens_per_image=200, seed=3)
    feats = _split(text, counts)
    ref = make_reference_loss(reference_losses, gamma, beta, log_tau)
    with torch.no_grad():
        r = ref(list(range(B)), tok, text_callback(feats), ddp_gather=False,
                need_attn_weights=True, compute_loss=True)
        o = oracle.radzero_forward(feats, tok, gamma, beta, log_tau, need_attn_weights=True)
    assert (r["t2i_logits"] - o["t2i_logits"]).abs().max() < 2e-6
    assert (r["t2i_attn_weights"][0] - o["t2i_attn_weights"][0]).abs().max() < 5e-5
    assert abs(r["losses"]["loss"].item() - o["losses"]["loss"].item()) < 1e-5


@pytest.mark.parametrize("B,N", [(1, 4), (3, 1), (1, 1), (2, 3)])
def test_squeeze_quirk_shapes(reference_losses, B, N):
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, tokens_per_image=30, seed=1)
    sl = reference_losses.SimilarityLogit("cos")
    tau = torch.exp(log_tau)
    r, _ = sl(text, tok, need_attn_weights=False, temperature=tau)
    o, _ = oracle.similarity_logit(text, tok, temperature=tau)
    assert tuple(r.shape) == tuple(o.shape)
    assert torch.allclose(r, o, atol=2e-6)


@pytest.mark.parametrize("sim_op", ["cos", "dot"])
@pytest.mark.parametrize("B,N", [(3, 4), (1, 5), (4, 1)])
def test_per_image_queries(reference_losses, sim_op, B, N):
    """repeat=False (losses.py:204-206): queries (B, N, D), one prompt set per image."""
    g = torch.Generator().manual_seed(B * 10 + N)
    q = torch.randn(B, N, 768, generator=g, dtype=torch.float64)
    tok = torch.randn(B, 37, 768, generator=g, dtype=torch.float64)
    sl = reference_losses.SimilarityLogit(sim_op)
    r, rs = sl(q, tok, need_attn_weights=True, repeat=False, temperature=torch.tensor(0.07, dtype=torch.float64))
    o, os_ = oracle.similarity_logit(q, tok, sim_op=sim_op, temperature=0.07, need_scores=True)
    assert tuple(r.shape) == tuple(o.shape)
    assert (r - o).abs().max() < 1e-10 and (rs[0] - os_[0]).abs().max() < 1e-8


def test_dot_branch(reference_losses):
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(2, 3, tokens_per_image=40, seed=5,
                                                            dtype=torch.float64)
    sl = reference_losses.SimilarityLogit("dot")
    r, rs = sl(text, tok, need_attn_weights=True)
    o, os_ = oracle.similarity_logit(text, tok, sim_op="dot", need_scores=True)
    assert (r - o).abs().max() < 1e-10 and (rs[0] - os_[0]).abs().max() < 1e-9


@pytest.mark.parametrize("row_sum,col_sum", [(False, False), (True, False), (False, True), (True, True)])
def test_mpnce_variants(reference_losses, row_sum, col_sum):
    g = torch.Generator().manual_seed(11)
    # fp32: the reference's row_sum branch allocates fp32 accumulators (losses.py:305-310)
    z = (torch.rand(13, 5, generator=g, dtype=torch.float32) * 2 - 1)
    gm = torch.tensor([0, 0, 1, 1, 1, 2, 3, 3, 3, 3, 4, 4, 4])
    r = reference_losses.multi_positive_nce_loss(z, gm, temperature=0.07, row_sum=row_sum, col_sum=col_sum)
    o = oracle.multi_positive_nce_loss(z, gm, temperature=0.07, row_sum=row_sum, col_sum=col_sum)
    assert abs(r.item() - o.item()) < 2e-6 * max(1.0, abs(r.item()))


@pytest.mark.parametrize("sim_op", ["cos", "dot"])
def test_contrastive_grads_match_reference(reference_losses, sim_op):
    B, counts = 3, [2, 1, 3]
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, 6, tokens_per_image=50, seed=9,
                                                            dtype=torch.float64)
    log_tau = log_tau.double()
    ref = make_reference_loss(reference_losses, gamma, beta, log_tau, sim_op=sim_op)
    t = text.clone().requires_grad_(True)
    x = tok.clone().requires_grad_(True)
    r = ref(list(range(B)), x, text_callback(_split(t, counts)), ddp_gather=False)
    r["losses"]["loss"].backward()
    loss, grads = oracle.contrastive_step_reference(
        text, oracle.build_group_map(counts), tok, gamma, beta, log_tau, sim_op=sim_op)
    assert abs(loss.item() - r["losses"]["loss"].item()) < 1e-10
    assert (grads["text"] - t.grad).abs().max() < 1e-10
    assert (grads["vision_tokens"] - x.grad).abs().max() < 1e-10
    assert (grads["gamma"] - ref.layer_norm.weight.grad).abs().max() < 1e-9
    assert (grads["beta"] - ref.layer_norm.bias.grad).abs().max() < 1e-9
    assert (grads["log_tau"] - ref.loss_temperature.grad).abs().max() < 1e-9


@pytest.mark.parametrize("size", [(518, 518), (1024, 1024), (300, 417), (64, 80), (37, 37), (20, 25)])
def test_bilinear_matches_aten(size):
    g = torch.Generator().manual_seed(2)
    grid = torch.randn(37, 37, generator=g)
    ref = F.interpolate(grid.view(1, 1, 37, 37), size=size, mode="bilinear", align_corners=False)[0, 0]
    out = oracle.bilinear_upsample(grid, *size)
    assert (ref - out).abs().max() < 2e-5
    ref64 = F.interpolate(grid.double().view(1, 1, 37, 37), size=size, mode="bilinear",
                          align_corners=False)[0, 0]
    out64 = oracle.bilinear_upsample(grid.double(), *size)
    assert (ref64 - out64).abs().max() < 1e-12


@pytest.mark.parametrize("kind", ["blip", "aspect_blip", "bit", "m3ae"])
@pytest.mark.parametrize("size", [(518, 518), (300, 417), (417, 300)])
def test_interpolate_variants_match_reference(kind, size):
    from tests._refload import load_reference_function, reference_present
    if not reference_present():
        pytest.skip("reference checkout not present")
    fn, kinds = load_reference_function("exp/cxr_pt/inference/segmentation_utils.py",
                                        "interpolate_similarity_scores")
    gp, _ = load_reference_function("exp/cxr_pt/inference/grounding_utils.py", "get_grounding_point")
    g = torch.Generator().manual_seed(4)
    s = torch.randn(1369, generator=g)
    r = fn(s, size, kinds[kind])
    o = oracle.interpolate_similarity_scores(s, size, kind)
    assert tuple(r.shape) == tuple(o.shape) == (1, size[0], size[1])
    assert (r - o).abs().max() < 2e-5
    assert gp(s, size, kinds[kind]) == oracle.grounding_point(s, size, kind)
