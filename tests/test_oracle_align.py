"""The AlignTransformer oracle (oracle/align.py) against transformers' Dinov2Encoder and the
committed golden vectors (tests/golden/align_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import align as oalign
from radzero_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.npz")


def _case(name):
    g = np.load(GOLDEN)
    B, L, seed = [int(v) for v in g[f"{name}.meta"]]
    tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0]
    w = synthetic.align_layer_weights(seed)
    got = np.array(synthetic.checksum(tok) + synthetic.checksum(w[1]["mlp.fc2.weight"]))
    assert np.allclose(got, g[f"{name}.checksum"], rtol=1e-9, atol=1e-6), "seeded inputs drifted"
    return tok, w, torch.from_numpy(g[f"{name}.out"])


@pytest.mark.parametrize("name", ["small", "ragged"])
def test_oracle_matches_golden(name):
    tok, w, want = _case(name)
    got = oalign.align_transformer(tok.double(), w)
    assert (got.float() - want).abs().max().item() <= 2e-5          # fixture is stored in fp32
    got32 = oalign.align_transformer(tok, w)
    assert (got32 - want).abs().max().item() <= 2e-4                 # fp32 evaluation order noise


def test_oracle_matches_transformers_module():
    enc = synthetic.build_align_encoder(seed=5).double()
    tok = synthetic.make_inputs(2, 1, tokens_per_image=50, seed=5)[0].double()
    with torch.no_grad():
        want = enc(tok)["last_hidden_state"]
    got = oalign.align_transformer(tok, synthetic.align_layer_weights(5))
    assert (got - want).abs().max().item() <= 1e-10


def test_attention_is_not_uniform():
    # the synthetic weights must exercise the softmax (a uniform average would hide layout bugs)
    tok, w, _ = _case("small")
    h = oalign._ln(tok, w[0]["norm1.weight"], w[0]["norm1.bias"], 1e-6)
    p = "attention.attention."
    q = torch.nn.functional.linear(h, w[0][p + "query.weight"], w[0][p + "query.bias"])
    k = torch.nn.functional.linear(h, w[0][p + "key.weight"], w[0][p + "key.bias"])
    s = (q[0, :, :64] @ k[0, :, :64].T) / 8.0
    assert s.std().item() > 0.5


def test_final_layer_norm_variant():
    tok, w, _ = _case("small")
    g, b = torch.rand(768) + 0.5, torch.rand(768) - 0.5
    y = oalign.align_transformer(tok, w)
    z = oalign.align_transformer(tok, w, final_ln=(g, b))
    assert torch.allclose(z, torch.nn.functional.layer_norm(y, (768,), g, b, 1e-5), atol=1e-5)


# ---- the backward: autograd through the oracle is the checker smoke() uses for the training step
BWD_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_bwd_golden.npz")


def _oracle_grads(B, L, seed, mag):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_align_bwd_golden as mk
    tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0].double().requires_grad_(True)
    w = [{k: v.double().requires_grad_(True) for k, v in layer.items()} for layer in synthetic.align_layer_weights(seed)]
    (oalign.align_transformer(tok, w) * mk.upstream(B, L, seed, mag)).sum().backward()
    return tok.grad, w, mk


@pytest.mark.parametrize("name", ["small", "ragged_tiny_grad"])
def test_oracle_autograd_matches_golden_gradients(name):
    """tests/golden/align_bwd_golden.npz was generated from transformers' Dinov2Encoder under autograd
    (make_align_bwd_golden.py); autograd through the oracle's restatement must give the same gradients."""
    g = np.load(BWD_GOLDEN)
    B, L, seed, mag = g[f"{name}.meta"]
    dtok, w, mk = _oracle_grads(int(B), int(L), int(seed), float(mag))
    scale = float(mag)
    assert np.abs(dtok.numpy() - g[f"{name}.dtokens"]).max() <= 1e-9 * scale
    for i, layer in enumerate(w):
        for key, t in layer.items():
            full = f"{name}.layer.{i}.{key}"
            if t.dim() == 1:
                assert np.abs(t.grad.numpy() - g[full]).max() <= 1e-8 * scale, full
            else:
                assert np.abs(t.grad[:4].numpy() - g[full + ".rows"]).max() <= 1e-8 * scale, full
                assert abs(t.grad.norm().item() - float(g[full + ".norm"])) <= 1e-9 * float(g[full + ".norm"]), full
                proj = (t.grad * mk.projection(t.grad.shape, int(seed))).sum().item()
                assert abs(proj - float(g[full + ".proj"])) <= 1e-8 * float(g[full + ".norm"]), full


def test_oracle_autograd_matches_transformers_autograd():
    enc = synthetic.build_align_encoder(seed=6).double().train()
    tok = synthetic.make_inputs(1, 1, tokens_per_image=40, seed=6)[0].double()
    up = torch.randn(tok.shape, generator=torch.Generator().manual_seed(6), dtype=torch.float64)
    t = tok.clone().requires_grad_(True)
    (enc(t)["last_hidden_state"] * up).sum().backward()
    x = tok.clone().requires_grad_(True)
    w = [{k: v.double().requires_grad_(True) for k, v in layer.items()} for layer in synthetic.align_layer_weights(6)]
    (oalign.align_transformer(x, w) * up).sum().backward()
    assert (x.grad - t.grad).abs().max().item() <= 1e-10
    for i, layer in enumerate(enc.layer):
        for key, p in layer.named_parameters():
            assert (w[i][key].grad - p.grad).abs().max().item() <= 1e-9, (i, key)
