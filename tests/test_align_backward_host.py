"""Host logic of the AlignTransformer backward (radzero_b200/align.py: _AlignFn, layer_backward, pack_layer_bwd,
layer_params) on CPU: the kernels are replaced by the fp64 test double tests/cpu_align_ops.py, the checker is
torch autograd through the stock transformers Dinov2Encoder.  What this pins without a GPU: the saved
activations, the sequence of products, LayerScale folded into the transposed weights + rz_ls_weight_bwd's
identity, the folded 1/8 of the query projection, the power-of-two gradient scale and the ORDER in which the
eighteen gradients of a layer go back to autograd.  The kernels themselves: tests/test_gpu_align_bwd.py."""
import copy

import pytest
import torch

from radzero_b200 import align, synthetic
from tests import cpu_align_ops


@pytest.fixture()
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(align, "ops", cpu_align_ops)
    monkeypatch.setattr(align, "OVERLAP_WEIGHT_GRADS", False)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


@pytest.mark.parametrize("mag", [1.0, 1e-6])
def test_backward_orchestration_matches_autograd(cpu_kernels, mag):
    B, L, seed = 2, 19, 71
    enc = synthetic.build_align_encoder(seed=seed)
    with torch.no_grad():       # LayerScale away from its initial 1.0, so that a missing / doubled factor shows
        for i, layer in enumerate(enc.layer):
            layer.layer_scale1.lambda1.mul_(0.5 + 0.1 * i).add_(torch.linspace(-0.2, 0.2, 768))
            layer.layer_scale2.lambda1.mul_(1.5 - 0.2 * i).add_(torch.linspace(0.1, -0.1, 768))
    ref = copy.deepcopy(enc).double().train()
    mod = align.AlignTransformer(enc).train()
    tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0]
    up = torch.randn(B, L, 768, generator=torch.Generator().manual_seed(seed)) * mag
    x = tok.clone().requires_grad_(True)
    params = [p for l in enc.layer for p in align.layer_params(l)]
    y = align._AlignFn.apply(x, mod, *params)
    (y * up).sum().backward()
    xd = tok.double().requires_grad_(True)
    yd = ref(xd)["last_hidden_state"]
    (yd * up.double()).sum().backward()
    # the packed weights are fp16-rounded (5e-4 relative), everything else is fp64 / fp32 here
    assert _rel(y.detach(), yd.detach()) <= 2e-3
    assert _rel(x.grad, xd.grad) <= 5e-3
    want = dict(ref.named_parameters())
    for name, p in enc.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
        if name.endswith("key.bias"):                      # exactly zero in exact arithmetic
            assert p.grad.norm().item() <= 1e-3 * enc.get_parameter(name.replace("key", "query")).grad.norm().item()
            continue
        assert _rel(p.grad, want[name].grad) <= 5e-3, (name, _rel(p.grad, want[name].grad))


def test_layer_params_cover_every_parameter_once():
    enc = synthetic.build_align_encoder(seed=1)
    for layer in enc.layer:
        listed = [p for p in align.layer_params(layer) if p is not None]
        assert len(listed) == 18 and len({id(p) for p in listed}) == 18
        assert {id(p) for p in listed} == {id(p) for p in layer.parameters()}


def test_backward_weights_fold_layer_scale_and_keep_the_query_unscaled():
    enc = synthetic.build_align_encoder(seed=2)
    layer = enc.layer[0]
    with torch.no_grad():
        layer.layer_scale1.lambda1.copy_(torch.linspace(0.5, 1.5, 768))
    wb = align.pack_layer_bwd(layer, "cpu")
    att = layer.attention.attention
    assert tuple(wb["wqkv_t"].shape) == (768, 2304) and tuple(wb["w2_t"].shape) == (3072, 768)
    assert torch.equal(wb["wqkv_t"][:, :768], att.query.weight.detach().t().half())            # no 1/8 here
    want = (layer.attention.output.dense.weight.detach() * layer.layer_scale1.lambda1.detach()[:, None]).t().half()
    assert torch.equal(wb["wo_t"], want)
    assert wb["wo32"].dtype == torch.float32 and torch.equal(wb["wo32"], layer.attention.output.dense.weight.detach())
    w = align.pack_layer(layer, "cpu")
    assert w["q_scale"] == 0.125
    assert torch.equal(w["wqkv"][:768], (att.query.weight.detach() * 0.125).half())              # the forward's operand
