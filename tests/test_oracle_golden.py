"""The oracle against the fixtures frozen from the reference (runs on any machine)."""
import numpy as np
import pytest
import torch

import oracle
from tests.golden_util import T, case_inputs, golden, split


@pytest.mark.parametrize("name", ["small", "full", "cls14"])
def test_forward_golden(name):
    tok, text, gamma, beta, log_tau, counts = case_inputs(name, dtype=torch.float64)
    o = oracle.radzero_forward(split(text, counts), tok, gamma, beta, log_tau.double(),
                               need_attn_weights=True, compute_loss=(name != "cls14"))
    assert (o["t2i_logits"] - T(f"{name}.t2i_logits", torch.float64)).abs().max() < 1e-6
    glue = oracle.compute_logits_glue(o["t2i_logits"], o["t2i_attn_weights"][0], log_tau.double())
    assert (glue["logits"] - T(f"{name}.logits", torch.float64)).abs().max() < 1e-5
    if name == "small":
        assert (o["t2i_attn_weights"][0] - T("small.scores", torch.float64)).abs().max() < 1e-5
        assert (glue["similarity_scores"] - T("small.similarity_scores", torch.float64)).abs().max() < 1e-5
    else:
        assert (o["t2i_attn_weights"][0][:, :, ::7] - T(f"{name}.scores_stride7", torch.float64)).abs().max() < 1e-5
    if name != "cls14":
        assert abs(o["losses"]["loss"].item() - float(golden()[f"{name}.loss"])) < 1e-8


@pytest.mark.parametrize("name", ["small", "full"])
def test_grads_golden(name):
    tok, text, gamma, beta, log_tau, counts = case_inputs(name, dtype=torch.float64)
    loss, g = oracle.contrastive_step_reference(text, oracle.build_group_map(counts), tok,
                                                gamma, beta, log_tau.double())
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()
    assert rel(g["text"], T(f"{name}.grad_text", torch.float64)) < 1e-5
    if name == "small":
        assert rel(g["vision_tokens"], T("small.grad_tokens", torch.float64)) < 1e-5
    else:
        assert rel(g["vision_tokens"][:, ::13], T("full.grad_tokens_stride13", torch.float64)) < 1e-5
    assert rel(g["gamma"], T(f"{name}.grad_gamma", torch.float64)) < 1e-6
    assert rel(g["beta"], T(f"{name}.grad_beta", torch.float64)) < 1e-6
    assert rel(g["log_tau"], T(f"{name}.grad_log_tau", torch.float64)) < 1e-6


@pytest.mark.parametrize("rs,cs", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_mpnce_golden(rs, cs):
    z = T("mpnce.z").clone().requires_grad_(True)
    gm = T("mpnce.group_map")
    l = oracle.multi_positive_nce_loss(z, gm, temperature=0.07, row_sum=bool(rs), col_sum=bool(cs))
    l.backward()
    assert abs(l.item() - float(golden()[f"mpnce.loss_r{rs}c{cs}"])) < 5e-6
    assert (z.grad - T(f"mpnce.dz_r{rs}c{cs}")).abs().max() < 5e-5


@pytest.mark.parametrize("B,N", [(1, 4), (3, 1), (1, 1), (2, 3)])
def test_quirk_shapes_golden(B, N):
    from radzero_b200 import synthetic
    tok, text, *_ = synthetic.make_inputs(B, N, tokens_per_image=30, seed=1)
    z, _ = oracle.similarity_logit(text, tok, temperature=0.07)
    assert list(z.shape) == [int(v) for v in golden()[f"quirk.B{B}N{N}.shape"]]


@pytest.mark.parametrize("key", ["s64x80", "s518", "s1024", "s300x417"])
def test_upsample_golden(key):
    grid = T("upsample.grid")
    h, w, stride = [int(v) for v in golden()[f"upsample.{key}.size_stride"]]
    m = oracle.interpolate_similarity_scores(grid, (h, w), "blip")[0]
    assert (m[::stride, ::stride] - T(f"upsample.{key}.map")).abs().max() < 1e-4  # fp32 lerp-weight noise on |v|~10
    assert abs(m.double().sum().item() - float(golden()[f"upsample.{key}.sum"])) < 1e-2
    assert list(oracle.grounding_point(grid, (h, w))) == [int(v) for v in golden()[f"upsample.{key}.point"]]


@pytest.mark.parametrize("kind", ["aspect_blip", "bit", "m3ae"])
@pytest.mark.parametrize("size", [(300, 417), (417, 300)])
def test_upsample_variants_golden(kind, size):
    grid = T("upsample.grid")
    m = oracle.interpolate_similarity_scores(grid, size, kind)[0]
    assert (m[::3, ::3] - T(f"upsample.{kind}.{size[0]}x{size[1]}.map")).abs().max() < 1e-4
    assert list(oracle.grounding_point(grid, size, kind)) == \
        [int(v) for v in golden()[f"upsample.{kind}.{size[0]}x{size[1]}.point"]]


def test_known_answers():
    """Self-made KATs: identical tokens, bounds, permutation, constant map."""
    D = 768
    g = torch.Generator().manual_seed(0)
    q = torch.randn(3, D, generator=g, dtype=torch.float64)
    k1 = torch.randn(1, 1, D, generator=g, dtype=torch.float64).expand(2, 17, D)
    z, s = oracle.similarity_logit(q, k1, temperature=0.07, need_scores=True, squeeze_quirk=False)
    cos = torch.nn.functional.cosine_similarity(q, k1[0, :1].expand(3, D), dim=-1)
    assert (z[:, 0] - cos).abs().max() < 1e-12          # all tokens identical -> Z = cos(q, k)
    assert s[0].abs().max() <= 1 / 0.07 + 1e-9          # scores within +-1/tau
    tok = torch.randn(4, 9, D, generator=g, dtype=torch.float64)
    z1, _ = oracle.similarity_logit(q, tok, temperature=0.07, squeeze_quirk=False)
    perm = torch.tensor([2, 0, 3, 1])
    z2, _ = oracle.similarity_logit(q, tok[perm], temperature=0.07, squeeze_quirk=False)
    assert (z1[:, perm] - z2).abs().max() < 1e-12        # permuting images permutes columns
    const = oracle.bilinear_upsample(torch.full((37, 37), 2.5), 100, 333)
    assert (const - 2.5).abs().max() < 1e-6              # constant map stays constant
    # one sentence per image, perfectly separated logits -> loss ~ 2*log(1 + (B-1) e^{-2/tau}) / 2
    B = 4
    zi = 2 * torch.eye(B, dtype=torch.float64) - 1
    l = oracle.multi_positive_nce_loss(zi, torch.arange(B), temperature=0.07)
    import math
    assert abs(l.item() - math.log(1 + (B - 1) * math.exp(-2 / 0.07))) < 1e-7


def test_masked_mean_pool_known_answers():
    """oracle.masked_mean_pool (modeling.py:147-156): mean over the unmasked tokens; an all-masked
    sentence gives 0 (sum 0 over clamp(0, min=1e-9)); a full mask is the plain mean."""
    import oracle
    torch.manual_seed(0)
    h = torch.randn(3, 5, 8, dtype=torch.float64)
    m = torch.tensor([[1, 1, 1, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]])
    out = oracle.masked_mean_pool(h, m)
    assert torch.allclose(out[0], h[0, :3].mean(0))
    assert torch.equal(out[1], torch.zeros(8, dtype=torch.float64))
    assert torch.allclose(out[2], h[2].mean(0))


def test_samplewise_dice_known_answer():
    """torchmetrics 1.6.1 DiceScore(num_classes=1): per-image Dice, zero denominator -> 1, nan-mean over images.
    Hand-computed: image 0: |P|=4, |G|=2, |P&G|=2 -> 2*2/6; image 1: |P|=0, |G|=3 -> 0; image 2: empty both -> 1."""
    pred = torch.zeros(3, 1, 4, 4, dtype=torch.long)
    gt = torch.zeros(3, 1, 4, 4, dtype=torch.long)
    pred[0, 0, 0, :4] = 1
    gt[0, 0, 0, :2] = 1
    gt[1, 0, 1, :3] = 1
    got = oracle.dice_score_samplewise(pred, gt)
    assert abs(float(got) - (2 * 2 / 6 + 0.0 + 1.0) / 3) < 1e-12
    # pooling the counts first would give 2*2 / (4 + 5): a different number
    assert abs(float(got) - 4 / 9) > 0.1


def test_best_dice_selection_is_samplewise_on_cpu_statistics():
    """radzero_b200.inference.best_dice_and_specificity on hand-made count tables (pure host logic)."""
    from radzero_b200 import inference
    pos = {"pred": torch.tensor([[10, 4, 0], [100, 50, 2]]), "inter": torch.tensor([[2, 2, 0], [50, 40, 2]]),
           "gt": torch.tensor([2, 60]), "thresholds": torch.tensor([0.0, 0.5, 0.9])}
    res = inference.best_dice_and_specificity(pos)
    per = torch.tensor([[4 / 12, 4 / 6, 0.0], [100 / 160, 80 / 110, 4 / 62]], dtype=torch.float64).mean(0)
    assert res["best_threshold"] == 0.5 and abs(res["dice"] - float(per[1])) < 1e-12
    pooled = inference.best_dice_and_specificity(pos, aggregate="pooled")
    assert abs(pooled["dice"] - max(2 * 52 / 172, 2 * 42 / 116, 2 * 2 / 64)) < 1e-12
