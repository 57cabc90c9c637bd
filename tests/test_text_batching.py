"""Host logic of the batched text side: merging the per-image token
batches into one padded call must give the same features and group_map as the reference's
one-call-per-image loop (exp/cxr_pt/model/losses.py:126-153).  CPU only: ``collect_text_features``
touches no kernel."""
import torch

from radzero_b200.losses import RadZeroLoss


class _FakeText:
    """A row-independent 'text model': masked mean of an embedding table (+ a call counter)."""

    def __init__(self):
        torch.manual_seed(0)
        self.table = torch.randn(50, 768)
        self.calls = 0

    def __call__(self, enc):
        self.calls += 1
        emb = self.table[enc["input_ids"]]
        m = enc["attention_mask"].unsqueeze(-1).float()
        feats = (emb * m).sum(1) / m.sum(1).clamp(min=1e-9)
        return {"text_features_wo_l2_norm": feats, "text_features": torch.nn.functional.normalize(feats, dim=1)}


def _phrases(counts, lens):
    g = torch.Generator().manual_seed(1)
    out = []
    for n, t in zip(counts, lens):
        ids = torch.randint(2, 50, (n, t), generator=g)
        am = torch.ones(n, t, dtype=torch.int64)
        for r in range(n):
            k = int(torch.randint(1, t + 1, (1,), generator=g))
            am[r, k:] = 0
        out.append({"input_ids": ids, "attention_mask": am})
    return out


def test_batched_call_equals_per_image_loop():
    kp = _phrases([3, 1, 9, 4], [5, 8, 6, 8])
    fn = RadZeroLoss(sim_op="cos")
    tm = _FakeText()
    text_b, gm_b = fn.collect_text_features(kp, tm, rank=2)
    assert tm.calls == 1
    fn.batch_text_calls = False
    tm2 = _FakeText()
    text_l, gm_l = fn.collect_text_features(kp, tm2, rank=2)
    assert tm2.calls == 4
    assert torch.equal(gm_b, gm_l) and gm_b.tolist()[:4] == [8, 8, 8, 9]        # i + rank * b_local
    assert torch.allclose(text_b, text_l, atol=1e-6)


def test_non_tensor_inputs_keep_the_reference_call_pattern():
    fn = RadZeroLoss(sim_op="cos")
    calls = []

    def text_model(kp):
        calls.append(kp)
        return {"text_features_wo_l2_norm": torch.ones(2, 768), "text_features": torch.ones(2, 768)}

    text, gm = fn.collect_text_features(["a", "b", "c"], text_model)
    assert len(calls) == 3 and text.shape == (6, 768) and gm.tolist() == [0, 0, 1, 1, 2, 2]
    kp = _phrases([2, 2], [4, 4])
    kp[0]["token_type_ids"] = torch.zeros(2, 4, dtype=torch.int64)
    calls.clear()
    fn.collect_text_features(kp, text_model)
    assert len(calls) == 2


def test_wide_features_are_sliced():
    fn = RadZeroLoss(sim_op="cos")
    tm = lambda enc: {"text_features_wo_l2_norm": torch.arange(1536.).repeat(enc["input_ids"].shape[0], 1),
                      "text_features": None}
    text, _ = fn.collect_text_features(_phrases([2, 3], [4, 6]), tm)
    assert text.shape == (5, 768) and text[0, 0].item() == 768.0               # losses.py:144-145
