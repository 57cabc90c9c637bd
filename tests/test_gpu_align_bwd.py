"""Backward of the AlignTransformer on the B200 kernels (radzero_b200/csrc/rz_align_bwd.cu + rz_linear)
against torch autograd in fp64 and the golden gradients of transformers' Dinov2Encoder
(tests/golden/make_align_bwd_golden.py).  The gradient chain is fp16 operands with fp32 accumulation under
one power-of-two scale, so the bar is the contrastive step's: every gradient tensor within 1e-2 relative
(Frobenius), the reference's own bf16-autocast arithmetic being several times coarser."""
import os
import sys

import numpy as np
import pytest
import torch

from radzero_b200 import ops, synthetic
from radzero_b200.align import AlignTransformer

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_bwd_golden.npz")


def _rel(got, want):
    got, want = got.double(), want.double()
    return ((got - want).norm() / want.norm().clamp_min(1e-300)).item()


def _key_bias_ok(got, name):
    """dL/d(key bias) is exactly zero in exact arithmetic (a key bias shifts every score of a row by the
    same q . b, which softmax ignores): the per-key gradients cancel, so the bar is absolute, a fraction
    of the query-bias gradient (same sum, no cancellation)."""
    ref = got[name.replace("key.bias", "query.bias")].double().norm().item()
    return got[name].double().norm().item() <= 1e-2 * ref


def _unit_scale():
    return ops.grad_scale(torch.full((8,), 40.0, device=DEV))       # 40 is in [32, 64): k = 0


@pytest.mark.parametrize("mag", [1.0, 3e-7, 5e4])
def test_grad_scale_is_a_power_of_two_from_the_maximum(mag):
    g = torch.randn(3, 1370, 768, device=DEV) * mag
    sc = ops.grad_scale(g)
    amax = g.abs().max().item()
    assert sc[2].item() == amax
    up, down, k = sc[0].item(), sc[1].item(), int(sc[3].item())
    assert up == 2.0 ** k and down == 2.0 ** -k
    assert 32.0 <= amax * up < 64.0
    assert torch.all(sc[4:] == down) and sc.numel() == 4 + 3072
    zero = ops.grad_scale(torch.zeros(16, device=DEV))
    assert zero[0].item() == 1.0 and zero[1].item() == 1.0


@pytest.mark.parametrize("rows", [1, 63, 200, 2 * 1370])
@pytest.mark.parametrize("cols", [768, 2304])
def test_transpose_pad_and_bias_gradient(rows, cols):
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols, device=DEV).half()
    sc = ops.grad_scale(torch.full((8,), 5.0, device=DEV))           # k = 3
    colsum = torch.ones(cols, device=DEV)
    out = ops.transpose_pad(x, sc, colsum)
    rp = (rows + 63) // 64 * 64
    assert tuple(out.shape) == (cols, rp)
    assert torch.equal(out[:, :rows], x.t()) and (out[:, rows:] == 0).all()
    want = 1.0 + x.double().sum(0) * sc[1].item()
    assert (colsum.double() - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())


def test_ls_cast_and_gelu_backward():
    torch.manual_seed(3)
    rows = 777
    dy = torch.randn(rows, 768, device=DEV)
    ls = torch.rand(768, device=DEV) + 0.5
    o = torch.randn(rows, 768, device=DEV).half()
    sc = ops.grad_scale(dy)
    dls = torch.zeros(768, device=DEV)
    do = ops.ls_cast_bwd(dy, ls, o, sc, dls)
    assert torch.equal(do, (dy * ls * sc[0]).half())
    assert _rel(dls, (dy.double() * o.double()).sum(0)) <= 1e-5
    u = (torch.randn(rows, 3072, device=DEV) * 2).half()
    dg = torch.randn(rows, 3072, device=DEV).half()
    ud = u.double().requires_grad_(True)
    torch.nn.functional.gelu(ud).backward(dg.double())
    got = ops.gelu_bwd(dg, u)
    assert (got.double() - ud.grad).abs().max().item() <= 4e-3


@pytest.mark.parametrize("n,k", [(768, 768), (768, 3072)])
def test_layer_scale_backward_from_the_weight_gradient(n, k):
    """y = ls * (x W^T + b): dW, db and dls from G = dy^T x and colsum(dy), against autograd."""
    torch.manual_seed(n + k)
    rows = 300
    x = torch.randn(rows, k, device=DEV, dtype=torch.float64)
    dy = torch.randn(rows, n, device=DEV, dtype=torch.float64)
    W = (torch.randn(n, k, device=DEV, dtype=torch.float64) * 0.05).requires_grad_(True)
    b = torch.randn(n, device=DEV, dtype=torch.float64).requires_grad_(True)
    ls = (torch.rand(n, device=DEV, dtype=torch.float64) + 0.5).requires_grad_(True)
    (ls * (x @ W.T + b)).backward(dy)
    g = (dy.T @ x).float().contiguous()
    cs = dy.sum(0).float().contiguous()
    dls = ops.ls_weight_bwd(g, W.detach().float().contiguous(), b.detach().float(), cs, ls.detach().float())
    assert _rel(g, W.grad) <= 1e-5 and _rel(cs, b.grad) <= 1e-5 and _rel(dls, ls.grad) <= 1e-4


@pytest.mark.parametrize("rows", [1, 5, 1370, 2 * 1370 + 3])
def test_ln_rows_backward(rows):
    torch.manual_seed(rows)
    x = torch.randn(rows, 768, device=DEV) * 2 + 0.3
    g = torch.rand(768, device=DEV) + 0.5
    b = torch.rand(768, device=DEV) - 0.5
    dh = torch.randn(rows, 768, device=DEV).half()
    dres = torch.randn(rows, 768, device=DEV)
    sc = ops.grad_scale(torch.full((8,), 5.0, device=DEV))           # k = 3: the kernel must undo it
    down = sc[1].item()
    xd, gd, bd = x.double().requires_grad_(True), g.double().requires_grad_(True), b.double().requires_grad_(True)
    torch.nn.functional.layer_norm(xd, (768,), gd, bd, 1e-6).backward(dh.double() * down)
    dgamma, dbeta = torch.zeros(768, device=DEV), torch.zeros(768, device=DEV)
    dx = ops.ln_rows_bwd(x, dh, g, 1e-6, dres, sc, dgamma, dbeta)
    assert _rel(dx, dres.double() + xd.grad) <= 1e-5
    assert _rel(dgamma, gd.grad) <= 1e-5 and _rel(dbeta, bd.grad) <= 1e-5
    # in place on the residual-path gradient, and without one
    dx16 = torch.empty(rows, 768, device=DEV, dtype=torch.float16)
    dx2 = ops.ln_rows_bwd(x, dh, g, 1e-6, dres, sc, None, None, out=dres, out16=dx16)
    assert dx2.data_ptr() == dres.data_ptr() and torch.equal(dx2, dx)
    assert torch.equal(dx16, (dx * sc[0]).half())            # the next product's operand: 2^k dx, rounded once
    assert _rel(ops.ln_rows_bwd(x, dh, g, 1e-6, None, sc, None, None), xd.grad) <= 1e-5


def _attn_autograd(qkv, dout, heads):
    B, L, W = qkv.shape
    d = W // 3
    x = qkv.double().requires_grad_(True)
    q, k, v = [t.view(B, L, heads, 64).transpose(1, 2) for t in x.split(d, dim=-1)]
    p = torch.softmax(q @ k.transpose(2, 3), dim=-1)       # the 1/sqrt(64) is already folded into q
    o = (p @ v).transpose(1, 2).reshape(B, L, d)
    o.backward(dout.double())
    return o.detach(), x.grad


@pytest.mark.parametrize("B,L,heads", [(1, 64, 1), (2, 70, 12), (1, 1370, 12), (3, 129, 2), (1, 1, 12)])
def test_attention_backward(B, L, heads):
    torch.manual_seed(B * 1000 + L)
    d = heads * 64
    qkv = torch.randn(B, L, 3 * d, device=DEV)
    qkv[..., :d] *= 0.35                                   # scores of a few units, as after the folded 1/8
    qkv = qkv.half()
    dout = torch.randn(B, L, d, device=DEV).half()
    out = ops.attention(qkv, heads)
    want_o, want = _attn_autograd(qkv, dout, heads)
    assert (out.double() - want_o).abs().max().item() <= 4e-3 * max(1.0, want_o.abs().max().item())
    got = ops.attention_bwd(qkv, out, dout, heads, 1.0)
    for i, name in enumerate("qkv"):
        w_i = want[..., i * d:(i + 1) * d]
        err = (got[..., i * d:(i + 1) * d].double() - w_i).norm().item()
        # one key (L = 1): dq = dk = 0 exactly, so the bar needs an absolute part
        assert err <= 3e-3 * w_i.norm().item() + 1e-4 * dout.double().norm().item(), (name, err)
    # q_scale multiplies the query block only; run-to-run identical (no atomics)
    again = ops.attention_bwd(qkv, out, dout, heads, 0.125)
    assert torch.equal(again[..., d:], got[..., d:])
    assert (again[..., :d].double() - want[..., :d] * 0.125).norm().item() <= 3e-3 * 0.125 * want[..., :d].norm().item() + 1e-4
    assert torch.equal(ops.attention_bwd(qkv, out, dout, heads, 0.125), again)


def _stock_grads(enc, tok, up):
    enc = enc.double().train()
    t = tok.double().requires_grad_(True)
    (enc(t)["last_hidden_state"] * up.double()).sum().backward()
    return t.grad, {n: p.grad for n, p in enc.named_parameters()}


@pytest.mark.parametrize("B,L,seed,mag", [(2, 70, 31, 1.0), (1, 1370, 32, 1e-5), (3, 200, 33, 300.0)])
def test_align_transformer_backward_vs_autograd(B, L, seed, mag):
    """The whole module in train mode: dL/dtokens and EVERY parameter gradient against the stock HF
    modules under torch autograd in fp64 on the same device; gradient magnitudes from 1e-5 to 300
    (the device-side power-of-two scale keeps the fp16 chain in range)."""
    enc = synthetic.build_align_encoder(seed=seed, device=DEV)
    mod = AlignTransformer(enc).train()
    tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0].to(DEV)
    up = torch.randn(B, L, 768, device=DEV, generator=torch.Generator(DEV).manual_seed(seed)) * mag
    t = tok.clone().requires_grad_(True)
    before = ops._lib.launch_count()
    y = mod(t)
    assert y.requires_grad and y.grad_fn is not None and "AlignFn" in type(y.grad_fn).__name__
    (y * up).sum().backward()
    assert ops._lib.launch_count() - before >= 2 * (7 + 24)        # our kernels ran, forward and backward
    got = {n: p.grad.clone() for n, p in enc.named_parameters()}
    assert all(g is not None for g in got.values())
    import copy
    want_dx, want = _stock_grads(copy.deepcopy(enc), tok, up)
    assert _rel(t.grad, want_dx) <= 1e-2, _rel(t.grad, want_dx)
    worst = max((_rel(got[n], want[n]), n) for n in want if not n.endswith("key.bias"))
    assert worst[0] <= 1e-2, worst
    assert all(_key_bias_ok(got, n) for n in want if n.endswith("key.bias"))
    assert max(want[n].norm().item() / want[n.replace("key.bias", "query.bias")].norm().item()
               for n in want if n.endswith("key.bias")) <= 1e-9
    # the forward under autograd is the inference forward
    with torch.no_grad():
        assert torch.equal(mod.eval()(tok), y.detach())


def test_align_transformer_backward_golden():
    import make_align_bwd_golden as mk
    gold = np.load(GOLDEN)
    for name in mk.CASES:
        B, L, seed, mag = gold[f"{name}.meta"]
        B, L, seed = int(B), int(L), int(seed)
        enc = synthetic.build_align_encoder(seed=seed, device=DEV)
        mod = AlignTransformer(enc).train()
        tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0].to(DEV).requires_grad_(True)
        (mod(tok) * mk.upstream(B, L, seed, float(mag)).float().to(DEV)).sum().backward()
        assert _rel(tok.grad.cpu(), torch.from_numpy(gold[f"{name}.dtokens"])) <= 1e-2
        grads = {n: p.grad for n, p in enc.named_parameters()}
        for pname, p in enc.named_parameters():
            g = p.grad.double().cpu()
            if pname.endswith("key.bias"):
                assert _key_bias_ok(grads, pname), (name, pname)
            elif g.dim() == 1:
                r = _rel(g, torch.from_numpy(gold[f"{name}.{pname}"]))
                assert r <= 1e-2, (name, pname, r)
            else:
                assert _rel(g[:4], torch.from_numpy(gold[f"{name}.{pname}.rows"])) <= 1e-2, (name, pname)
                norm = float(gold[f"{name}.{pname}.norm"])
                assert abs(g.norm().item() - norm) <= 1e-2 * norm, (name, pname)
                proj = (g * mk.projection(g.shape, seed)).sum().item()
                assert abs(proj - float(gold[f"{name}.{pname}.proj"])) <= 4e-2 * norm, (name, pname)   # sum((g - w) P) ~ |g - w| for a unit-variance P


def test_training_step_never_calls_the_stock_modules(monkeypatch):
    enc = synthetic.build_align_encoder(seed=5, device=DEV)
    mod = AlignTransformer(enc).train()

    def boom(*a, **k):
        raise AssertionError("stock Dinov2Encoder.forward called")
    monkeypatch.setattr(enc, "forward", boom)
    tok = synthetic.make_inputs(1, 1, tokens_per_image=90, seed=5)[0].to(DEV)
    opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
    first = None
    for _ in range(2):
        opt.zero_grad()
        loss = mod(tok).square().mean()
        loss.backward()
        opt.step()                                         # bumps the version counters: weights re-packed
        first = loss.item() if first is None else first
    assert loss.item() < first
    mod.kernel_backward = False
    with pytest.raises(AssertionError):
        mod(tok)


def test_config4_training_step_through_align_transformer_and_loss():
    """The trainable chain of radzero.yaml (`module_to_update: [align_transformer, ...]`): ViT output ->
    AlignTransformer (train mode, kernels under autograd) -> RadZeroLoss.forward (contrastive step kernels)
    -> backward.  Loss and every gradient against the stock HF encoder + the CPU-side oracle of the loss
    (oracle/vlcabs.py) chained under torch autograd in fp64."""
    import copy

    import oracle
    from radzero_b200 import losses
    from tests.golden_util import split
    B, counts, L = 3, [4, 2, 5], 257
    vit, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts), tokens_per_image=L, seed=61)
    enc = synthetic.build_align_encoder(seed=61, device=DEV)
    ref_enc = copy.deepcopy(enc).double().train()
    mod = AlignTransformer(enc).train()
    fn = losses.RadZeroLoss(sim_op="cos").to(DEV)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma)
        fn.layer_norm.bias.copy_(beta)
        fn.loss_temperature.copy_(log_tau.reshape(fn.loss_temperature.shape))
    t = text.to(DEV).requires_grad_(True)
    x = vit.to(DEV).requires_grad_(True)
    feats = split(t, counts)
    out = fn(list(range(B)), mod(x), lambda i: {"text_features_wo_l2_norm": feats[i]})
    loss = out["losses"]["loss"]
    loss.backward()
    # the checker
    xd = vit.double().to(DEV).requires_grad_(True)
    td = text.double().to(DEV).requires_grad_(True)
    gd, bd = gamma.double().to(DEV).requires_grad_(True), beta.double().to(DEV).requires_grad_(True)
    ltd = log_tau.double().to(DEV).requires_grad_(True)
    tok = ref_enc(xd)["last_hidden_state"]
    tau = torch.exp(ltd)
    z, _ = oracle.similarity_logit(oracle.layer_norm_rows(td, gd, bd), oracle.layer_norm_rows(tok, gd, bd),
                                   temperature=tau, sim_op="cos", squeeze_quirk=False)
    want = oracle.multi_positive_nce_loss(z, oracle.build_group_map(counts).to(DEV), temperature=tau)
    want.backward()
    assert abs(loss.item() - want.item()) <= 1e-3 * abs(want.item())
    assert _rel(x.grad, xd.grad) <= 1e-2 and _rel(t.grad, td.grad) <= 1e-2
    assert _rel(fn.layer_norm.weight.grad, gd.grad) <= 1e-2 and _rel(fn.layer_norm.bias.grad, bd.grad) <= 1e-2
    got = {n: p.grad for n, p in enc.named_parameters()}
    want_g = {n: p.grad for n, p in ref_enc.named_parameters()}
    worst = max((_rel(got[n], want_g[n]), n) for n in want_g if not n.endswith("key.bias"))
    assert worst[0] <= 1.5e-2, worst
    assert all(_key_bias_ok(got, n) for n in want_g if n.endswith("key.bias"))


@pytest.mark.parametrize("B,L", [(8, 1370), (16, 257)])
def test_attention_backward_is_run_to_run_deterministic(B, L):
    """The two kernels double-buffer their tiles with cp.async and reuse the fixed operands' tiles as the second
    buffer: a missing barrier there would show up as run-to-run differences (this is how round 2 found the race
    in the small-N forward).  No atomics in these kernels, so every launch must be bit-identical."""
    torch.manual_seed(L)
    H = 12
    qkv = (torch.randn(B, L, 3 * H * 64, device=DEV) * 0.6).half()
    out = ops.attention(qkv, H)
    dout = torch.randn(B, L, H * 64, device=DEV).half()
    first = ops.attention_bwd(qkv, out, dout, H, 0.125)
    assert torch.isfinite(first.float()).all()
    for _ in range(40):
        assert torch.equal(ops.attention_bwd(qkv, out, dout, H, 0.125), first)


def test_attention_backward_in_image_slices(monkeypatch):
    """More than 65 535 (image, head) pairs go through several launches: forced here with a small limit."""
    torch.manual_seed(9)
    B, L, H = 5, 150, 12
    qkv = (torch.randn(B, L, 3 * H * 64, device=DEV) * 0.5).half()
    out = ops.attention(qkv, H)
    dout = torch.randn(B, L, H * 64, device=DEV).half()
    whole = ops.attention_bwd(qkv, out, dout, H, 0.125)
    monkeypatch.setattr(ops, "_ATTN_BWD_MAX_BH", 2 * H)          # slices of 2, 2 and 1 images
    assert torch.equal(ops.attention_bwd(qkv, out, dout, H, 0.125), whole)
