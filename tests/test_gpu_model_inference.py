"""BASELINE.json configs[0]: ``model_inference(image, text, tokenizer, image_processor, model)`` on the
random-init RadZero architecture (DINOv2-base @518 + 2 align layers + MPNet-base), one synthetic
chest X-ray at the image processor's default resolution, the 14 CheXpert finding prompts.

The encoders are stock HF modules (out of scope); what is checked is everything after them: the
CUDA path's (similarity_prob, similarity_map) against the CPU oracle fed with the SAME vision
tokens and sentence embeddings.  Tolerances are the north star's."""
import numpy as np
import pytest
import torch

import oracle
from radzero_b200 import inference, modeling, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    transformers = pytest.importorskip("transformers")
    model = modeling.build_random_init_model(device="cuda")
    proc = transformers.BlipImageProcessor(size={"height": 518, "width": 518})
    tok = inference.SyntheticTokenizer()
    image = synthetic.synthetic_cxr(600, 520, seed=7)            # (H, W) uint8, not square on purpose
    prompts = [f"There is {f}" for f in synthetic.CHEXPERT_FINDINGS]
    return model, proc, tok, image, prompts


def _oracle_outputs(model, proc, tok, image, prompts):
    from PIL import Image
    pil = Image.fromarray(image.numpy()).convert("RGB")
    pix = torch.as_tensor(np.array(proc(pil)["pixel_values"]), dtype=torch.float32).to(model.device)
    enc = tok(prompts, padding=True, truncation=True, return_tensors="pt").to(model.device)
    with torch.no_grad():
        tokens = model.forward_vision_model(pix)["vision_tokens"].float().cpu()
        text = model.forward_text_model(enc)["text_features_wo_l2_norm"].float().cpu()
    fn = model.loss_fns["RadZeroLoss"]
    ref = oracle.radzero_forward([text[i:i + 1] for i in range(text.shape[0])], tokens,
                                 fn.layer_norm.weight.detach().cpu(), fn.layer_norm.bias.detach().cpu(),
                                 fn.loss_temperature.detach().cpu(), need_attn_weights=True,
                                 compute_loss=False, squeeze_quirk=False)
    return oracle.compute_logits_glue(ref["t2i_logits"], ref["t2i_attn_weights"][0],
                                      fn.loss_temperature.detach().cpu())


def test_model_inference_matches_oracle(setup):
    model, proc, tok, image, prompts = setup
    assert len(prompts) == 14
    prob, smap = inference.model_inference(image, prompts, tok, proc, model)
    H, W = image.shape
    assert prob.shape == (14,) and smap.shape == (14, H, W)
    glue = _oracle_outputs(model, proc, tok, image, prompts)
    want_prob = torch.sigmoid(glue["logits"][0])
    assert ((prob.cpu() - want_prob) / want_prob).abs().max() < 1e-3
    assert int(prob.argmax()) == int(want_prob.argmax())                     # zero-shot label
    for n in (0, 5, 13):
        want = oracle.interpolate_similarity_scores(glue["similarity_scores"][0, n], (H, W))[0]
        assert (smap[n].cpu() - want).abs().max() < 2e-3
        # thresholded segmentation mask (visualization scripts use sigmoid > 0.7 / 0.4)
        m_got = torch.sigmoid(smap[n].cpu()) > 0.5
        m_want = torch.sigmoid(want) > 0.5
        near = (torch.sigmoid(want) - 0.5).abs() < 1e-3
        assert ((m_got != m_want) & ~near).sum() == 0


def test_single_prompt_shapes_and_compute_logits(setup):
    model, proc, tok, image, prompts = setup
    prob, smap = inference.model_inference(image, prompts[3], tok, proc, model)
    assert prob.dim() == 0 and smap.shape == tuple(image.shape)
    from PIL import Image
    pil = Image.fromarray(image.numpy()).convert("RGB")
    pix = torch.as_tensor(np.array(proc(pil)["pixel_values"]), dtype=torch.float32).to(model.device)
    enc = tok(prompts, padding=True, truncation=True, return_tensors="pt").to(model.device)
    out = model.compute_logits(pix, [enc])
    assert out["logits"].shape == (1, 14) and out["similarity_scores"].shape == (1, 14, 1369)
    glue = _oracle_outputs(model, proc, tok, image, prompts)
    assert (out["similarity_scores"].cpu() - glue["similarity_scores"]).abs().max() < 2e-3
    assert torch.equal(out["logits"].cpu().argmax(1), glue["logits"].argmax(1))
    p, m = model.similarity(pix, enc)
    assert p.shape == (1, 14) and m.shape == (1, 14, 37, 37)
