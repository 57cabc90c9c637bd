"""Host-side logic of the AlignTransformer mirror that needs no GPU: weight packing (the 1/sqrt(64)
fold into the query projection, the fused q|k|v operand, state-dict compatibility with the reference's
``AlignTransformer``), staleness detection, and the no-CPU-fallback contract."""
import pytest
import torch

from radzero_b200 import synthetic
from radzero_b200._lib import RzError
from radzero_b200.align import AlignTransformer, pack_layer


@pytest.fixture(scope="module")
def enc():
    return synthetic.build_align_encoder(seed=4, layers=2)


def test_pack_layer_folds_the_attention_scale_exactly(enc):
    layer = enc.layer[0]
    w = pack_layer(layer, "cpu")
    att = layer.attention.attention
    assert w["heads"] == 12 and w["wqkv"].shape == (2304, 768) and w["wqkv"].dtype == torch.float16
    # 1/8 is a power of two: scaling commutes with the fp16 rounding (except for entries that the
    # scaling pushes into fp16's subnormal range, |w| < 2^-11: absolute error < 6e-8 there)
    wq = att.query.weight.detach()
    want = wq.half() * 0.125
    normal = wq.abs() >= 2.0 ** -11
    assert torch.equal(w["wqkv"][:768][normal], want[normal])
    assert (w["wqkv"][:768].float() - wq * 0.125).abs().max().item() <= 2.0 ** -11 * 0.125
    assert torch.equal(w["wqkv"][768:1536], att.key.weight.detach().half())
    assert torch.equal(w["wqkv"][1536:], att.value.weight.detach().half())
    assert torch.equal(w["bqkv"][:768], att.query.bias.detach() * 0.125)
    assert torch.equal(w["bqkv"][768:1536], att.key.bias.detach())
    assert w["eps1"] == 1e-6 and w["eps2"] == 1e-6
    assert torch.equal(w["ls1"], layer.layer_scale1.lambda1.detach())
    assert w["w1"].shape == (3072, 768) and w["w2"].shape == (768, 3072)


def test_state_dict_keys_match_the_reference_module(enc):
    """align_transformers.py:23-35: ``transformer_layers`` (Dinov2Encoder) and optional ``layer_norm``."""
    mod = AlignTransformer(enc, torch.nn.LayerNorm(768))
    keys = set(mod.state_dict().keys())
    assert "transformer_layers.layer.0.attention.attention.query.weight" in keys
    assert "transformer_layers.layer.1.mlp.fc2.bias" in keys
    assert "transformer_layers.layer.0.layer_scale1.lambda1" in keys
    assert "layer_norm.weight" in keys and "layer_norm.bias" in keys
    assert not any("_packed" in k for k in keys)


def test_packed_weights_follow_in_place_parameter_changes(enc):
    mod = AlignTransformer(enc)
    a = mod._weights(torch.device("cpu"))
    assert mod._weights(torch.device("cpu")) is a                    # cached
    with torch.no_grad():
        enc.layer[1].mlp.fc1.bias.add_(1.0)
    w1_before = a[1]["w1"].clone()
    with torch.no_grad():
        enc.layer[1].mlp.fc1.weight.mul_(2.0)
    b = mod._weights(torch.device("cpu"))
    assert b is not a and torch.equal(b[1]["bf1"], enc.layer[1].mlp.fc1.bias.detach())
    assert torch.allclose(b[1]["w1"].float(), w1_before.float() * 2, atol=1e-6)   # the fp16 operand was re-packed
    mod.load_state_dict(mod.state_dict())                            # copy_ bumps the version counters
    assert mod._weights(torch.device("cpu")) is not b


def test_cpu_tensors_raise_in_both_modes_and_the_stock_arm_is_opt_in(enc):
    mod = AlignTransformer(enc).eval()
    with pytest.raises(RzError):
        mod(torch.zeros(1, 4, 768))                                  # no CPU fallback for the kernels
    with pytest.raises(RzError):
        mod(torch.zeros(1, 4, 512, device="cpu"))
    mod.train()
    x = torch.randn(1, 5, 768)
    with pytest.raises(RzError):
        mod(x)                                                       # training runs on the kernels too: no CPU path
    mod.kernel_backward = False
    y = mod(x)                                                       # the comparison arm: reference's own forward
    assert y.requires_grad and y.shape == x.shape
    want = enc(x)["last_hidden_state"]
    assert torch.allclose(y, want)


def test_unsupported_layer_variants_are_rejected():
    from transformers import Dinov2Config
    from transformers.models.dinov2.modeling_dinov2 import Dinov2Encoder
    swiglu = Dinov2Encoder(Dinov2Config(hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                                        use_swiglu_ffn=True))
    with pytest.raises(RzError):
        pack_layer(swiglu.layer[0], "cpu")
    heads16 = Dinov2Encoder(Dinov2Config(hidden_size=768, num_hidden_layers=1, num_attention_heads=16))
    with pytest.raises(RzError):
        pack_layer(heads16.layer[0], "cpu")
