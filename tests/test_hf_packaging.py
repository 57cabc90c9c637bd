"""Surface 2 of SURVEY.md section 8b on CPU: ``CxrAlignModel`` is a transformers ``PreTrainedModel`` with
``config_class = CxrAlignConfig`` (exp/cxr_pt/model/modeling.py:23-25), ``save_pretrained`` /
``AutoModel.from_pretrained`` round-trip (README.md:77-82) and keeps the reference's state-dict keys."""
import torch
from transformers import AutoConfig, AutoModel, PreTrainedModel

from radzero_b200 import modeling


def _small():
    cfg = modeling.CxrAlignConfig(vision_config={"num_hidden_layers": 1}, text_config={"num_hidden_layers": 1})
    torch.manual_seed(0)
    return modeling.CxrAlignModel(cfg)


def test_auto_model_round_trip(tmp_path):
    m = _small()
    assert isinstance(m, PreTrainedModel) and modeling.CxrAlignModel.config_class is modeling.CxrAlignConfig
    with torch.no_grad():
        m.loss_fns["RadZeroLoss"].loss_temperature.fill_(-2.5)
        m.loss_fns["RadZeroLoss"].layer_norm.weight.uniform_(0.5, 1.5)
    m.save_pretrained(tmp_path)
    cfg = AutoConfig.from_pretrained(tmp_path)
    assert isinstance(cfg, modeling.CxrAlignConfig) and cfg.compute_logits_type == "radzero"
    assert cfg.align_transformer_config["num_hidden_layers"] == 2 and cfg.vision_config["img_size"] == 518
    m2 = AutoModel.from_pretrained(tmp_path)
    assert type(m2) is modeling.CxrAlignModel
    a, b = m.state_dict(), m2.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_reference_state_dict_keys():
    keys = set(_small().state_dict())
    for k in ("loss_fns.RadZeroLoss.loss_temperature", "loss_fns.RadZeroLoss.layer_norm.weight",
              "loss_fns.RadZeroLoss.layer_norm.bias",
              "align_transformer.transformer_layers.layer.1.mlp.fc2.weight",
              "align_transformer.transformer_layers.layer.0.layer_scale1.lambda1",
              "vision_model.embeddings.cls_token", "text_model.embeddings.word_embeddings.weight"):
        assert k in keys, k


def test_modules_can_be_passed_in_and_loss_defaults_follow_the_yaml():
    m = _small()
    m2 = modeling.CxrAlignModel(m.vision_model, m.align_transformer, m.text_model)
    fn = m2.loss_fns["RadZeroLoss"]
    assert fn.sim_op == "cos" and fn.use_vision_cls_token and m2.loss_ratio == {"RadZeroLoss": 1.0}
    assert m2.vision_model is m.vision_model
