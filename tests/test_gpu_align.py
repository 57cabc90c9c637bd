"""Parity of the AlignTransformer kernels (rz_ln_rows, rz_linear, rz_attention) and of the
assembled ``radzero_b200.align.AlignTransformer`` against the CPU oracle (oracle/align.py) and the
golden vectors.  Tolerances: the kernels take fp16 operands with fp32 accumulation, so element
errors are ~1e-3 relative; the end-to-end bar is BASELINE.json's (cosine similarities <= 2e-3)."""
import os

import numpy as np
import pytest
import torch

from oracle import align as oalign
from oracle import vlcabs as orc
from radzero_b200 import ops, synthetic
from radzero_b200.align import AlignTransformer, pack_layer

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.npz")


def _h(t):   # fp16-rounded copy in fp64: what the kernel's operand actually holds
    return t.to(torch.float16).double()


@pytest.mark.parametrize("rows", [1, 37, 2740])
def test_ln_rows(rows):
    torch.manual_seed(rows)
    x = torch.randn(rows, 768, device=DEV) * 3 + 0.5
    g, b = torch.rand(768, device=DEV) + 0.5, torch.rand(768, device=DEV) - 0.5
    got = ops.ln_rows(x, g, b, 1e-6).float()
    want = torch.nn.functional.layer_norm(x.double(), (768,), g.double(), b.double(), 1e-6).float()
    assert (got - want).abs().max().item() <= 4e-3 * want.abs().max().item()
    assert (got - want).abs().mean().item() <= 4e-4


@pytest.mark.parametrize("m", [1, 128, 200, 257, 1370, 2 * 1370 + 5])
@pytest.mark.parametrize("epi,k,n", [("bias", 768, 2304), ("gelu", 768, 3072), ("residual", 768, 768),
                                     ("residual", 3072, 768), ("bias", 64, 256)])
def test_linear(m, epi, k, n):
    torch.manual_seed(m * 7 + k + n)
    a = torch.randn(m, k, device=DEV).half()
    w = (torch.randn(n, k, device=DEV) * 0.05).half()
    bias = torch.randn(n, device=DEV) * 0.1
    acc = a.double() @ w.double().T + bias.double()
    if epi == "bias":
        got = ops.linear(a, w, bias, "bias").double()
        want = acc
    elif epi == "gelu":
        got = ops.linear(a, w, bias, "gelu").double()
        want = oalign.gelu_erf(acc)
    else:
        res = torch.randn(m, n, device=DEV)
        scale = torch.rand(n, device=DEV) + 0.5
        want = res.double() + scale.double() * acc
        keep = res.clone()
        got = ops.linear(a, w, bias, "residual", scale=scale, residual=res).double()
        assert torch.equal(res, keep)                      # not in place unless asked
        inplace = ops.linear(a, w, bias, "residual", scale=scale, residual=res, out=res)
        assert inplace.data_ptr() == res.data_ptr() and torch.equal(inplace.double(), got)
    tol = 2e-3 if epi != "residual" else 1e-4              # fp16 output rounding vs fp32 output
    err = (got - want).abs().max().item()
    assert err <= tol * max(1.0, want.abs().max().item()), err


def test_linear_no_bias_no_scale():
    a = torch.randn(300, 768, device=DEV).half()
    w = (torch.randn(768, 768, device=DEV) * 0.05).half()
    res = torch.randn(300, 768, device=DEV)
    got = ops.linear(a, w, None, "residual", residual=res).double()
    want = res.double() + a.double() @ w.double().T
    assert (got - want).abs().max().item() <= 1e-4 * want.abs().max().item()


def _attn_ref(qkv, heads):
    B, L, W = qkv.shape
    d = W // 3
    q, k, v = [t.view(B, L, heads, 64).transpose(1, 2).double() for t in qkv.split(d, dim=-1)]
    p = torch.softmax(q @ k.transpose(2, 3), dim=-1)       # the scale is already folded into q
    return (p @ v).transpose(1, 2).reshape(B, L, d)


@pytest.mark.parametrize("B,L,heads", [(1, 1, 12), (2, 90, 12), (1, 128, 12), (1, 129, 3), (2, 1370, 12),
                                       (3, 257, 1)])
def test_attention(B, L, heads):
    torch.manual_seed(L + heads)
    qkv = torch.randn(B, L, 3 * heads * 64, device=DEV)
    qkv[..., : heads * 64] *= 0.6                          # scores ~ N(0, (0.6*8)^2): a peaked softmax
    qkv = qkv.half()
    got = ops.attention(qkv, heads).double()
    want = _attn_ref(qkv, heads)
    assert (got - want).abs().max().item() <= 3e-3, (got - want).abs().max().item()


def test_attention_growing_scores():
    """Keys ordered so that the running maximum rises in every tile (the online-softmax rescale of
    the register accumulator is exercised at every step), plus a query whose scores fall."""
    B, L, heads = 1, 1000, 2
    torch.manual_seed(3)
    qkv = torch.randn(B, L, 3 * heads * 64, device=DEV) * 0.05
    ramp = torch.linspace(0.0, 6.0, L, device=DEV)
    qkv[0, :, 0] = 4.0                                     # q[:, head 0, dim 0]
    qkv[0, : L // 2, 0] = -4.0                             # first half of the queries: falling scores
    qkv[0, :, heads * 64 + 0] = ramp                       # k[:, head 0, dim 0] grows with the key index
    qkv = qkv.half()
    got = ops.attention(qkv, heads).double()
    want = _attn_ref(qkv, heads)
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= 2e-3


def _gpu_weights(seed):
    enc = synthetic.build_align_encoder(seed=seed, device=DEV)
    return enc, synthetic.align_layer_weights(seed)


@pytest.mark.parametrize("name", ["small", "ragged"])
def test_align_transformer_golden(name):
    g = np.load(GOLDEN)
    B, L, seed = [int(v) for v in g[f"{name}.meta"]]
    enc, _ = _gpu_weights(seed)
    tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0].to(DEV)
    keep = tok.clone()
    got = AlignTransformer(enc)(tok)
    assert torch.equal(tok, keep)                          # the caller's tokens are not modified
    want = torch.from_numpy(g[f"{name}.out"]).to(DEV)
    err = (got - want).abs()
    assert err.max().item() <= 2e-2 and err.mean().item() <= 2e-3, (err.max().item(), err.mean().item())
    cos = torch.nn.functional.cosine_similarity(got, want, dim=-1)
    assert (1 - cos).max().item() <= 2e-5


def test_align_then_similarity_full_size():
    """AlignTransformer -> VL-CABS similarity at the real token count against the oracle chain:
    the cosine similarities must stay within BASELINE.json's 2e-3."""
    B, N, seed = 2, 14, 21
    enc, w = _gpu_weights(seed)
    tok, text, gamma, beta, log_tau = [t.to(DEV) for t in synthetic.make_inputs(B, N, seed=seed)]
    x_got = AlignTransformer(enc)(tok)
    x_ref = oalign.align_transformer(tok.double(), [{k: v.to(DEV) for k, v in l.items()} for l in w])
    assert (x_got.double() - x_ref).abs().max().item() <= 3e-2
    k_got = orc.l2_normalize_rows(orc.layer_norm_rows(x_got.double(), gamma.double(), beta.double()))
    k_ref = orc.l2_normalize_rows(orc.layer_norm_rows(x_ref, gamma.double(), beta.double()))
    q = orc.l2_normalize_rows(orc.layer_norm_rows(text.double(), gamma.double(), beta.double()))
    cos_err = (k_got @ q.T - k_ref @ q.T).abs().max().item()
    assert cos_err <= 2e-3, cos_err
    # and through the fused kernel of the path itself
    q16, _, _ = ops.prep_rows(text, gamma, beta)
    out = ops.sim_fwd_tokens(x_got, gamma, beta, q16, 1.0, want_scores=True, drop_cls=False)
    assert (out["scores"].double() - (k_ref @ q.T).transpose(1, 2)).abs().max().item() <= 2e-3


def test_final_layer_norm_and_refresh():
    enc, w = _gpu_weights(9)
    ln = torch.nn.LayerNorm(768).to(DEV)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.2, 0.2)
    tok = synthetic.make_inputs(1, 1, tokens_per_image=200, seed=9)[0].to(DEV)
    mod = AlignTransformer(enc, ln)
    got = mod(tok)
    want = oalign.align_transformer(tok.double(), [{k: v.to(DEV) for k, v in l.items()} for l in w],
                                    final_ln=(ln.weight.detach().double(), ln.bias.detach().double()))
    assert (got.double() - want).abs().max().item() <= 2e-2
    with torch.no_grad():
        enc.layer[0].layer_scale1.lambda1.zero_()
        enc.layer[0].layer_scale2.lambda1.zero_()
    # no refresh(): the in-place change is detected through the parameters' version counters
    w[0]["layer_scale1.lambda1"].zero_()
    w[0]["layer_scale2.lambda1"].zero_()
    want2 = oalign.align_transformer(tok.double(), [{k: v.to(DEV) for k, v in l.items()} for l in w],
                                     final_ln=(ln.weight.detach().double(), ln.bias.detach().double()))
    assert (mod(tok).double() - want2).abs().max().item() <= 2e-2


def test_cpu_tensors_raise():
    from radzero_b200._lib import RzError
    with pytest.raises(RzError):
        AlignTransformer(None)(torch.zeros(1, 4, 768))
    with pytest.raises(RzError):
        ops.attention(torch.zeros(1, 4, 2304, dtype=torch.float16), 12)


def test_inplace_and_grad_dispatch():
    enc, w = _gpu_weights(13)
    mod = AlignTransformer(enc).eval()
    tok = synthetic.make_inputs(1, 1, tokens_per_image=300, seed=13)[0].to(DEV)
    want = mod(tok)
    buf = tok.clone()
    got = mod(buf, inplace=True)
    assert got.data_ptr() == buf.data_ptr() and torch.equal(got, want)
    # bf16 tokens (the reference's training / bf16_full_eval dtype) are accepted and come back as bf16,
    # the dtype the reference's module would return
    got16 = mod(tok.to(torch.bfloat16))
    assert got16.dtype == torch.bfloat16 and (got16.float() - want).abs().max().item() <= 0.25
    # an input that itself requires grad (eval-mode saliency) is served under autograd, not detached
    with torch.enable_grad():
        assert mod(tok.clone().requires_grad_(True)).requires_grad
    # eval mode never leaves the kernels, with or without torch.no_grad()
    with torch.enable_grad():
        assert not mod(tok).requires_grad
    # train mode + gradients needed -> the same kernels under autograd (tests/test_gpu_align_bwd.py)
    mod.train()
    with torch.enable_grad():
        y = mod(tok)
    assert y.requires_grad
    assert (y.detach() - want).abs().max().item() <= 3e-2
    with torch.no_grad():
        assert not mod(tok).requires_grad


def test_model_vision_tokens_match_stock_modules():
    """The random-init RadZero architecture (BASELINE configs[0]): the tokens the kernels hand to the
    VL-CABS path against the stock HF AlignTransformer forward on the same ViT output, and the cosine
    similarities that follow."""
    from radzero_b200 import modeling
    model = modeling.build_random_init_model(device=DEV, vision_layers=2, text_layers=1)
    torch.manual_seed(0)
    vit_out = torch.randn(2, 1370, 768, device=DEV)
    with torch.no_grad():
        got = model.align_transformer(vit_out)
        want = model.align_transformer.stock_forward(vit_out.double().float())
    assert (got - want).abs().max().item() <= 2e-2
    fn = model.loss_fns["RadZeroLoss"]
    g, b = fn.layer_norm.weight.detach().double(), fn.layer_norm.bias.detach().double()
    text = torch.randn(14, 768, device=DEV, dtype=torch.float64)
    q = orc.l2_normalize_rows(orc.layer_norm_rows(text, g, b))
    kg = orc.l2_normalize_rows(orc.layer_norm_rows(got.double(), g, b))
    kw = orc.l2_normalize_rows(orc.layer_norm_rows(want.double(), g, b))
    assert (kg @ q.T - kw @ q.T).abs().max().item() <= 2e-3


def test_empty_inputs_and_validation():
    """Empty batches return empty outputs; malformed operands raise instead of launching."""
    from radzero_b200._lib import RzError
    g = torch.ones(768, device=DEV)
    assert ops.ln_rows(torch.empty(0, 768, device=DEV), g, g, 1e-6).shape == (0, 768)
    w = torch.zeros(768, 768, device=DEV, dtype=torch.float16)
    assert ops.linear(torch.empty(0, 768, device=DEV, dtype=torch.float16), w, None).shape == (0, 768)
    assert ops.attention(torch.empty(0, 5, 2304, device=DEV, dtype=torch.float16), 12).shape == (0, 5, 768)
    f, q = ops.text_pool(torch.empty(0, 4, 768, device=DEV), torch.empty(0, 4, device=DEV, dtype=torch.int64), g, g)
    assert f.shape == (0, 768) and q.shape == (0, 768)
    assert AlignTransformer(synthetic.build_align_encoder(seed=1, layers=1, device=DEV)).eval()(
        torch.empty(0, 7, 768, device=DEV)).shape == (0, 7, 768)
    with pytest.raises(RzError):
        ops.linear(torch.zeros(4, 100, device=DEV, dtype=torch.float16),
                   torch.zeros(256, 100, device=DEV, dtype=torch.float16), None)       # K % 64 != 0
    with pytest.raises(RzError):
        ops.linear(torch.zeros(4, 64, device=DEV, dtype=torch.float16),
                   torch.zeros(100, 64, device=DEV, dtype=torch.float16), None)        # N % 256 != 0
    with pytest.raises(RzError):
        ops.linear(torch.zeros(4, 64, device=DEV), torch.zeros(256, 64, device=DEV), None)   # fp32 operands
    with pytest.raises(RzError):
        ops.attention(torch.zeros(1, 4, 2304, device=DEV, dtype=torch.float16), 11)    # width != 3 * heads * 64
    with pytest.raises(RzError):
        ops.linear(torch.zeros(4, 64, device=DEV, dtype=torch.float16),
                   torch.zeros(256, 64, device=DEV, dtype=torch.float16), None, "residual")  # no residual given


def test_attention_many_items_odd_count():
    """An odd number of (image, head, query tile) items: the second pipeline of the last CTA runs a
    duplicate item whose output must not be stored; and more item pairs than SMs (persistent loop)."""
    for B, L, heads in [(1, 100, 1), (1, 300, 1), (3, 300, 1), (13, 300, 12)]:
        torch.manual_seed(B * L)
        qkv = (torch.randn(B, L, 3 * heads * 64, device=DEV) * 0.7).half()
        got = ops.attention(qkv, heads).double()
        want = _attn_ref(qkv, heads)
        assert (got - want).abs().max().item() <= 3e-3, (B, L, heads)


@pytest.mark.parametrize("scale", [0.05, 2.0, 5.0])
def test_attention_score_ranges(scale):
    """Near-uniform, peaked and almost one-hot softmax rows (|scores| up to ~150): the lazy maximum, the
    -126 clamp of the emulated exp2 and the fp16 probabilities must hold over the whole range."""
    torch.manual_seed(int(scale * 100))
    B, L, heads = 2, 700, 4
    qkv = torch.randn(B, L, 3 * heads * 64, device=DEV)
    qkv[..., : 2 * heads * 64] *= scale
    qkv = qkv.half()
    got = ops.attention(qkv, heads).double()
    want = _attn_ref(qkv, heads)
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= 4e-3, (got - want).abs().max().item()


def test_f16_handoff_of_the_last_layer():
    """the last fc2 epilogue hands the tokens over as fp16 (RZ_LIN_RESIDUAL_F16)
    and the small-N similarity kernel consumes them directly -- same values as the fp32 path rounded once,
    and the similarity within the north star's tolerances."""
    from radzero_b200 import losses
    torch.manual_seed(5)
    m, k, n = 2 * 1370 + 5, 3072, 768
    a = torch.randn(m, k, device=DEV).half()
    w = (torch.randn(n, k, device=DEV) * 0.05).half()
    bias = torch.randn(n, device=DEV) * 0.1
    scale = torch.rand(n, device=DEV) + 0.5
    res = torch.randn(m, n, device=DEV) * 3
    full = ops.linear(a, w, bias, "residual", scale=scale, residual=res)
    half = ops.linear(a, w, bias, "residual_f16", scale=scale, residual=res)
    assert half.dtype == torch.float16 and torch.equal(half, full.to(torch.float16))
    # the assembled module + the similarity behind it
    B, N, seed = 3, 14, 23
    enc, _ = _gpu_weights(seed)
    tok, text, gamma, beta, log_tau = [t.to(DEV) for t in synthetic.make_inputs(B, N, seed=seed)]
    mod = AlignTransformer(enc).eval()
    keep = tok.clone()
    x32 = mod(tok)
    x16 = mod(tok, handoff_f16=True)
    assert torch.equal(tok, keep) and x16.dtype == torch.float16 and x16.shape == x32.shape
    assert torch.equal(x16, x32.to(torch.float16))
    fn = losses.RadZeroLoss(sim_op="cos").to(DEV)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma)
        fn.layer_norm.bias.copy_(beta)
    p32 = fn.similarity_prob(text, x32)
    p16 = fn.similarity_prob(text, x16)
    assert ((p16 - p32) / p32).abs().max().item() < 1e-3
    _, s32, _ = fn.similarity(text, x32, want_scores=True)
    _, s16, _ = fn.similarity(text, x16, want_scores=True)
    assert (s16 - s32).abs().max().item() < 2e-3
