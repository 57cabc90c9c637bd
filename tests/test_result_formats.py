"""Result formats of the zero-shot evaluators, host code only:
prompt construction (inference/utils.py:40-66), the similarities CSV (:213-215) and result.json
(common/utils.py:123-125)."""
import json
import os

import numpy as np
import torch

from radzero_b200 import inference


class _Model:
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def compute_logits(self, pixel_values, encoded_key_phrases, **kw):
        self.calls.append((pixel_values.shape[0], encoded_key_phrases[0]["input_ids"].shape[0]))
        b, n = pixel_values.shape[0], encoded_key_phrases[0]["input_ids"].shape[0]
        base = pixel_values.reshape(b, -1).sum(1, keepdim=True)
        return {"logits": base + torch.arange(n, dtype=torch.float32)[None, :]}


def test_process_class_prompts_builds_positive_and_negated_batches():
    tok = inference.SyntheticTokenizer()
    prompts = {"0": ["There is pneumonia", "unused"], "1": ["There is edema"], "2": ["There is no finding"]}
    out = inference.process_class_prompts(prompts, tok, _Model())
    assert out["encoded_key_phrases"]["input_ids"].shape[0] == 3
    assert out["encoded_negative_phrases"]["input_ids"].shape[0] == 3
    # "There is" -> "There is no" lengthens every prompt by one token under any whitespace tokenizer
    lp = out["encoded_key_phrases"]["attention_mask"].sum(1)
    ln = out["encoded_negative_phrases"]["attention_mask"].sum(1)
    assert torch.equal(ln, lp + 1)


def test_calculate_similarities_concatenates_batches_in_order():
    tok = inference.SyntheticTokenizer()
    text = inference.process_class_prompts({"0": ["There is a"], "1": ["There is b"]}, tok, _Model())
    model = _Model()
    batches = [torch.ones(3, 1, 2, 2), 2 * torch.ones(2, 1, 2, 2)]
    sims = inference.calculate_similarities(batches, text, model)
    assert sims.dtype == np.float32 and sims.shape == (5, 2) and model.calls == [(3, 2), (2, 2)]
    assert np.allclose(sims[:, 0], [4, 4, 4, 8, 8]) and np.allclose(sims[:, 1] - sims[:, 0], 1)


def test_csv_and_json_match_the_reference_writers(tmp_path):
    import pandas as pd
    sims = np.array([[0.125, -3.5, 1e-7], [2.0, 1 / 3, 14.285714]], dtype=np.float32)
    path = inference.save_similarities_csv(sims, str(tmp_path), "OpenI")
    assert path == os.path.join(str(tmp_path), "OpenI.csv")
    text = open(path).read()
    assert text.splitlines()[0] == "0,1,2" and len(text.splitlines()) == 3
    want = pd.DataFrame(sims).to_csv(index=False)                 # the reference's own call
    assert text == want
    back = pd.read_csv(path).to_numpy(dtype=np.float32)
    assert np.array_equal(back, sims)
    res = {"OpenI": {"AUC": 0.9, "per_class": [0.1, 0.2]}, "é": 1}
    jpath = inference.save_result_json(res, str(tmp_path / "out"))
    raw = open(jpath, encoding="utf-8").read()
    assert raw == json.dumps(res, indent=2) and json.loads(raw) == res
