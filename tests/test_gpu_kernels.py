"""Parity of the HBM-bound CUDA kernels (prep, upsample, MP-NCE) against the oracle and
the golden fixtures frozen from the reference.  All calls go through the C ABI."""
import numpy as np
import pytest
import torch

import oracle
from radzero_b200 import _lib, ops, synthetic
from tests.golden_util import T, case_inputs, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------------------- prep
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,L,Lp", [(7, 7, 7), (3 * 50, 50, 64), (2 * 1370, 1370, 1408)])
def test_prep_rows(dtype, rows, L, Lp):
    tok, text, gamma, beta, _ = synthetic.make_inputs(rows // L, 1, tokens_per_image=L, seed=5)
    x = tok.reshape(rows, 768).to(dtype)
    f16, f32, st = ops.prep_rows(x.to(DEV), gamma.to(DEV), beta.to(DEV), rows_per_group=L,
                                 rows_per_group_padded=Lp, want_f32=True, want_stats=True)
    xd = x.double()
    ln = oracle.layer_norm_rows(xd, gamma.double(), beta.double())
    want = oracle.l2_normalize_rows(ln)
    assert (f32.cpu().double() - want).abs().max() < 2e-6          # fp32 math on identical inputs
    h = f16.cpu().view(rows // L, Lp, 768)
    assert (h[:, :L].double().reshape(rows, 768) - want).abs().max() < 6e-4  # fp16 rounding of |v|<=1
    if Lp > L:
        assert h[:, L:].abs().max() == 0                            # padding rows are zero
    mu = xd.mean(-1)
    rstd = 1 / torch.sqrt(xd.var(-1, unbiased=False) + 1e-5)
    assert (st[:, 0].cpu().double() - mu).abs().max() < 1e-5
    assert ((st[:, 1].cpu().double() - rstd) / rstd).abs().max() < 1e-5
    assert ((st[:, 2].cpu().double() * ln.norm(dim=-1)) - 1).abs().max() < 1e-5


def test_prep_rows_no_layernorm_and_no_l2():
    x = torch.randn(33, 768, generator=torch.Generator().manual_seed(1))
    _, f32, _ = ops.prep_rows(x.to(DEV), None, None, want_f16=False, want_f32=True)
    assert (f32.cpu() - oracle.l2_normalize_rows(x)).abs().max() < 2e-6
    g, b = torch.rand(768) + 0.5, torch.rand(768) - 0.5
    _, f32, _ = ops.prep_rows(x.to(DEV), g.to(DEV), b.to(DEV), want_f16=False, want_f32=True, l2=False)
    assert (f32.cpu() - oracle.layer_norm_rows(x, g, b)).abs().max() < 2e-5


def test_cpu_tensors_are_rejected():
    with pytest.raises(_lib.RzError):
        ops.prep_rows(torch.zeros(2, 768), None, None)


# ------------------------------------------------------------------------------- upsample
@pytest.mark.parametrize("key", ["s64x80", "s518", "s1024", "s300x417"])
def test_upsample_golden(key):
    grid = T("upsample.grid").to(DEV)
    h, w, stride = [int(v) for v in golden()[f"upsample.{key}.size_stride"]]
    m = ops.upsample_maps(grid.view(1, -1), (h, w))[0].cpu()
    assert (m[::stride, ::stride] - T(f"upsample.{key}.map")).abs().max() < 2e-4
    assert abs(m.double().sum().item() - float(golden()[f"upsample.{key}.sum"])) < 0.05
    pt = ops.upsample_maps(grid.view(1, -1), (h, w), mode=_lib.RZ_UP_ARGMAX)[0].cpu().tolist()
    assert pt == [int(v) for v in golden()[f"upsample.{key}.point"]]
    mask = ops.upsample_maps(grid.view(1, -1), (h, w), mode=_lib.RZ_UP_MASK, threshold=0.7)[0].cpu()
    want_mask = torch.sigmoid(oracle.interpolate_similarity_scores(grid.cpu(), (h, w))[0]) > 0.7
    band = (torch.sigmoid(oracle.interpolate_similarity_scores(grid.cpu(), (h, w))[0]) - 0.7).abs() < 1e-4
    assert ((mask.bool() != want_mask) & ~band).sum() == 0          # identical outside the fp-noise band
    assert abs(int(mask.sum()) - int(golden()[f"upsample.{key}.mask_gt0p7_count"])) <= int(band.sum())


@pytest.mark.parametrize("kind", ["blip", "aspect_blip", "bit", "m3ae"])
@pytest.mark.parametrize("size", [(300, 417), (417, 300), (518, 518)])
def test_upsample_processor_variants(kind, size):
    from radzero_b200.inference import interpolate_params
    g = torch.Generator().manual_seed(8)
    scores = torch.randn(5, 1369, generator=g) * 4
    kw = interpolate_params(size, kind)
    out = ops.upsample_maps(scores.to(DEV), size, **kw).cpu()
    sig = ops.upsample_maps(scores.to(DEV), size, mode=_lib.RZ_UP_SIGMOID, **kw).cpu()
    pts = ops.upsample_maps(scores.to(DEV), size, mode=_lib.RZ_UP_ARGMAX, **kw).cpu()
    for i in range(5):
        want = oracle.interpolate_similarity_scores(scores[i], size, kind)[0]
        assert (out[i] - want).abs().max() < 2e-4
        # sigmoid through ONE tanh.approx (2^-11 relative): <= 2.5e-4 absolute; the path's tolerance is 2e-3
        assert (sig[i] - torch.sigmoid(want)).abs().max() < 4e-4
        x, y = pts[i].tolist()
        assert want[y, x] >= want.max() - 2e-4


def test_upsample_padded_pitch_batch():
    """(B, N, G*G) views of a padded-pitch buffer (the large-N similarity map) are upsampled in place."""
    g = torch.Generator().manual_seed(11)
    store = torch.randn(3, 5, 1376, generator=g).to(DEV)
    view = store[:, :, 1:1370]
    assert not view.is_contiguous()
    got = ops.upsample_maps(view, (70, 90))
    want = ops.upsample_maps(view.reshape(-1, 1369).contiguous(), (70, 90))
    assert got.shape == (15, 70, 90)
    assert torch.equal(got, want)


def test_upsample_constant_and_strided_input():
    s = torch.full((3, 1369), 2.5, device=DEV)
    assert (ops.upsample_maps(s, (100, 333)) - 2.5).abs().max() < 1e-6
    big = torch.randn(4, 6, 1370, device=DEV)
    view = big[:, :, 1:].reshape(24, 1369)      # what compute_logits hands over (CLS dropped)
    out = ops.upsample_maps(view, (74, 74)).cpu()
    want = oracle.bilinear_upsample(view.cpu().view(24, 37, 37), 74, 74)
    assert (out - want).abs().max() < 1e-4


# ------------------------------------------------------------------------------- MP-NCE
def _mpnce_cuda(z, gm, tau, b_global=None, col0=0, row_sum=False, col_sum=False):
    b_global = b_global or z.shape[1]
    rs, ps, cn, cp = ops.mpnce_partials(z, gm, col0, 1.0 / tau, col_sum=col_sum, b_global=b_global)
    terms, dz = ops.mpnce_finish(z, gm, col0, b_global, 1.0 / tau, rs, ps, cn, cp,
                                 row_sum=row_sum, col_sum=col_sum)
    return terms, dz, (rs, ps, cn, cp)


@pytest.mark.parametrize("rs,cs", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_mpnce_golden(rs, cs):
    z = T("mpnce.z").to(DEV)
    gm = T("mpnce.group_map").to(DEV)
    terms, dz, _ = _mpnce_cuda(z, gm, 0.07, row_sum=bool(rs), col_sum=bool(cs))
    n, b = z.shape
    loss = (terms[0] / (b if rs else n) + terms[1] / (b if cs else n)) / 2
    want = float(golden()[f"mpnce.loss_r{rs}c{cs}"])
    assert abs(loss.item() - want) < 1e-3 * abs(want)
    ref_dz = T(f"mpnce.dz_r{rs}c{cs}")
    assert (dz.cpu() - ref_dz).abs().max() < 1e-3 * ref_dz.abs().max()


@pytest.mark.parametrize("n,b", [(45, 8), (700, 300), (6144, 1024)])
def test_mpnce_vs_oracle_and_sharded(n, b):
    g = torch.Generator().manual_seed(n)
    counts = torch.randint(1, 2 * n // b + 1, (b,), generator=g)
    gm = torch.repeat_interleave(torch.arange(b), counts)[:n]
    gm = torch.cat([gm, torch.full((n - gm.numel(),), b - 1)]) if gm.numel() < n else gm
    z = (torch.rand(n, b, generator=g) * 2 - 1)
    z[torch.arange(n), gm] += 0.5
    z.clamp_(-1, 1)
    zd = z.double().requires_grad_(True)
    want = oracle.multi_positive_nce_loss(zd, gm, temperature=0.07)
    want.backward()
    terms, dz, _ = _mpnce_cuda(z.to(DEV), gm.to(DEV), 0.07)
    loss = (terms[0] + terms[1]) / (2 * n)
    assert abs(loss.item() - want.item()) < 1e-4 * abs(want.item())
    assert (dz.cpu().double() - zd.grad).abs().max() < 1e-4 * zd.grad.abs().max()
    # d loss / d log(tau) = -sum dZ*Z
    assert abs(terms[2].item() - float((zd.grad * zd.detach()).sum())) < 1e-3 * abs(float((zd.grad * zd.detach()).sum())) + 1e-6
    # image-sharded evaluation (emulating W ranks on one GPU): all-reduce = plain sum
    W = 4 if b % 4 == 0 else 1
    if W > 1:
        bl = b // W
        zc = z.to(DEV)
        parts = [ops.mpnce_partials(zc[:, r * bl:(r + 1) * bl].contiguous(), gm.to(DEV), r * bl, 1 / 0.07, b_global=b)
                 for r in range(W)]
        rowsum = sum(p[0] for p in parts)
        pos = sum(p[1] for p in parts)
        tot = torch.zeros(4, device=DEV)
        for r in range(W):
            zr = zc[:, r * bl:(r + 1) * bl].contiguous()
            t, dzr = ops.mpnce_finish(zr, gm.to(DEV), r * bl, b, 1 / 0.07, rowsum, pos, parts[r][2], parts[r][3])
            tot += t
            assert (dzr.cpu().double() - zd.grad[:, r * bl:(r + 1) * bl]).abs().max() < 1e-4 * zd.grad.abs().max()
        assert abs(((tot[0] + tot[1]) / (2 * n)).item() - want.item()) < 1e-4 * abs(want.item())


# ------------------------------------------------------------------------------ fused consumers
@pytest.mark.parametrize("kind,size", [("blip", (97, 131)), ("bit", (150, 120)), ("blip", (518, 518))])
def test_dice_sweep_stats_vs_oracle(kind, size):
    """Dice threshold sweep / specificity statistics straight from the patch-grid scores
    (segmentation_utils.py:255-261, 136-158) against the oracle's upsample -> sigmoid -> `> t` loop."""
    from radzero_b200 import inference
    g = torch.Generator().manual_seed(21)
    M = 6 if size[0] < 500 else 2
    scores = torch.randn(M, 1369, generator=g) * 3.0
    yy, xx = torch.meshgrid(torch.arange(size[0]), torch.arange(size[1]), indexing="ij")
    masks = torch.stack([(((yy - size[0] * (0.3 + 0.1 * m)) ** 2 + (xx - size[1] * 0.5) ** 2) < (8 + 4 * m) ** 2)
                         for m in range(M)]).to(torch.uint8)
    masks[-1] = 0                                                  # a negative image
    got = inference.dice_sweep_stats(scores.to(DEV), masks.to(DEV), size, kind)
    want = oracle.dice_sweep_stats(scores, masks, size, kind)
    assert torch.equal(got["gt"].cpu(), want["gt"])
    assert (got["max_prob"].cpu() - want["max_prob"]).abs().max() < 1e-5
    # counts agree except for pixels whose probability sits within fp32 noise of a threshold
    thr = want["thresholds"]
    for m in range(M):
        prob = torch.sigmoid(oracle.interpolate_similarity_scores(scores[m], size, kind)[0])
        near = torch.stack([((prob - float(t)).abs() < 2e-5).sum() for t in thr])
        assert ((got["pred"][m].cpu() - want["pred"][m]).abs() <= near).all()
        assert ((got["inter"][m].cpu() - want["inter"][m]).abs() <= near).all()
    pos = {k: (v[:-1] if v.dim() and v.shape[0] == M else v) for k, v in got.items()}
    neg = {k: (v[-1:] if v.dim() and v.shape[0] == M else v) for k, v in got.items()}
    res = inference.best_dice_and_specificity(pos, neg)
    # checker: the reference's loop itself -- DiceScore(num_classes=1)((probs > t).long(), masks) per threshold
    # (torchmetrics' samplewise Dice restated in the oracle), on the oracle's pixel maps
    probs = torch.stack([torch.sigmoid(oracle.interpolate_similarity_scores(scores[m], size, kind)[0])
                         for m in range(M - 1)])
    sweep = torch.stack([oracle.dice_score_samplewise(probs > float(t), masks[:-1]) for t in thr])
    assert abs(res["dice"] - float(sweep.max())) < 1e-3
    assert abs(float(sweep[int(round(res["best_threshold"] * 100))]) - float(sweep.max())) < 1e-3
    # the pooled aggregation is a different number and stays available as an explicit option
    wp, wi, wg = want["pred"][:-1].sum(0).double(), want["inter"][:-1].sum(0).double(), want["gt"][:-1].sum().double()
    pooled = inference.best_dice_and_specificity(pos, neg, aggregate="pooled")
    assert abs(pooled["dice"] - float((2 * wi / (wp + wg).clamp_min(1.0)).max())) < 1e-3
    negprob = torch.sigmoid(oracle.interpolate_similarity_scores(scores[-1], size, kind))
    assert res["specificity"] == oracle.compute_specificity(negprob, res["best_threshold"])


# ------------------------------------------------------------------------------- T0 text pooling
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,T", [(1, 1), (14, 9), (37, 32), (600, 17)])
def test_text_pool_matches_oracle(n, T, dtype):
    """rz_text_pool against the oracle's masked mean (modeling.py:147-156) + LN + L2 (ragged masks, a
    fully masked sentence, padding tokens holding garbage that must never be read into the mean)."""
    import oracle
    from radzero_b200 import ops
    torch.manual_seed(n * 31 + T)
    hidden = (torch.randn(n, T, 768, device="cuda") * 2 + 0.3).to(dtype)
    lens = torch.randint(1, T + 1, (n,), device="cuda")
    if n > 2:
        lens[1] = 0                                              # all-masked sentence: clamp(min=1e-9) branch
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).long()
    hidden = torch.where(mask.bool()[..., None], hidden, torch.full_like(hidden, float("nan")))
    gamma = torch.rand(768, device="cuda") + 0.5
    beta = torch.rand(768, device="cuda") * 0.4 - 0.2
    feats, q16 = ops.text_pool(hidden, mask, gamma, beta)
    clean = torch.nan_to_num(hidden.double(), nan=0.0)
    want = oracle.masked_mean_pool(clean, mask)
    assert torch.isfinite(feats).all() and (feats.double() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    wq = oracle.l2_normalize_rows(oracle.layer_norm_rows(feats.double(), gamma.double(), beta.double()))
    assert (q16.double() - wq).abs().max().item() <= 1e-3
    # identical to the two-step path of the product (prep_rows on the pooled rows)
    q_two, _, _ = ops.prep_rows(feats, gamma, beta)
    assert (q16.float() - q_two.float()).abs().max().item() <= 2e-3
    # no LayerNorm / no L2 variants
    _, q_plain = ops.text_pool(hidden, mask, None, None, l2=False, want_feats=False)
    assert (q_plain.double() - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("n_images,first", [(1, 0), (7, 14), (1024, 1024), (2500, 3)])
def test_group_map_from_counts(n_images, first):
    """rz_group_map: counts travel as launch parameters; result = np.repeat (losses.py:131-151)."""
    import numpy as np
    from radzero_b200 import ops
    rng = np.random.default_rng(n_images)
    counts = rng.integers(0, 12, size=n_images)
    counts[rng.integers(0, n_images)] = 300
    got = ops.group_map_from_counts(counts.tolist(), first, "cuda").cpu().numpy()
    want = np.repeat(np.arange(first, first + n_images, dtype=np.int64), counts)
    assert np.array_equal(got, want)
    assert ops.group_map_from_counts([0, 0], 0, "cuda").numel() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("size,kind", [((518, 518), "blip"), ((300, 417), "blip"), ((64, 80), "bit"), ((1024, 1024), "blip"),
                                       ((37, 33), "blip")])
def test_upsample_bit_packed_mask(size, kind):
    """RZ_UP_MASK_BITS: the thresholded mask with one bit per pixel (bit x % 32 of word x // 32, padding bits
    zero) equals the byte mask of RZ_UP_MASK (sigmoid(score) > t, segmentation_utils.py:225, 258)."""
    import numpy as np
    from radzero_b200 import inference, ops, _lib
    g = torch.Generator().manual_seed(size[0] + size[1])
    scores = torch.randn(5, 37 * 37, generator=g).cuda() * 3
    kw = inference.interpolate_params(size, kind)
    for thr in (0.5, 0.7, 0.0, 1.0):
        byte = ops.upsample_maps(scores, size, mode=_lib.RZ_UP_MASK, threshold=thr, **kw).cpu().numpy()
        bits = ops.upsample_maps(scores, size, mode=_lib.RZ_UP_MASK_BITS, threshold=thr, **kw).cpu().numpy()
        H, W = size
        assert bits.shape == (5, H, (W + 31) // 32) and bits.dtype == np.int32
        un = np.unpackbits(bits.view(np.uint8).reshape(5, H, -1), axis=-1, bitorder="little")
        assert np.array_equal(un[:, :, :W], byte)
        assert not un[:, :, W:].any()
