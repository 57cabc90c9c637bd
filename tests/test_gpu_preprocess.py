"""GPU image preprocessing (rz_preprocess_images) against the golden outputs of the reference's chain and
the oracle: BIT-exact in fp32 (integer resize, table-driven normalisation)."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as P
from radzero_b200 import ops, preprocess
from tests.preprocess_cases import CASES, make_raw

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_golden.npz")))


def _want(name):
    r = G[f"{name}.resized_u8"]
    lut = P.normalize_lut()
    return np.stack([lut[c][r if r.ndim == 2 else r[..., c]] for c in range(3)])


def _to_dev(raw):
    if raw.dtype == np.uint16:
        return torch.from_numpy(raw.view(np.int16)).to(DEV).view(torch.uint16)
    return torch.from_numpy(raw).to(DEV)


@pytest.mark.parametrize("name", list(CASES))
def test_preprocess_bit_exact_vs_golden(name):
    spec = CASES[name]
    raw = make_raw(spec)
    pv = ops.preprocess_images(_to_dev(raw)[None], spec["size"], mean=P.OPENAI_CLIP_MEAN, std=P.OPENAI_CLIP_STD)
    assert pv.shape == (1, 3) + tuple(spec["size"]) and pv.dtype == torch.float32
    want = _want(name)
    got = pv[0].cpu().numpy()
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {want.size} values differ"
    if spec.get("keep_pv"):
        assert np.array_equal(got, G[f"{name}.pixel_values"])


def test_batch_of_images_each_with_its_own_range():
    spec = CASES["dicom_u16"]
    raws = [make_raw(dict(spec, seed=s)) // d for s, d in ((11, 1), (12, 3), (13, 7), (14, 2), (15, 5))]
    batch = np.stack(raws)
    pv = ops.preprocess_images(_to_dev(batch), (518, 518), mean=P.OPENAI_CLIP_MEAN, std=P.OPENAI_CLIP_STD)
    for i, r in enumerate(raws):
        assert np.array_equal(pv[i].cpu().numpy(), P.preprocess_image(r, (518, 518))), i


def test_collate_fn_mirror_mixed_sizes_and_16bit_output():
    from PIL import Image
    from transformers import BlipImageProcessor
    proc = BlipImageProcessor(size={"height": 224, "width": 224}, image_mean=[0.5, 0.4, 0.3], image_std=[0.2, 0.25, 0.3])
    items = [Image.fromarray(make_raw(CASES["small_u8_upsample"])), make_raw(CASES["signed_i16"]),
             Image.fromarray(make_raw(CASES["rgb_u8"])), make_raw(dict(CASES["small_u8_upsample"], seed=33))]
    pv = preprocess.collate_fn(items, proc, device=DEV)
    assert pv.shape == (4, 3, 224, 224) and pv.is_cuda
    for i, it in enumerate(items):
        want = P.preprocess_image(np.array(it), (224, 224), mean=[0.5, 0.4, 0.3], std=[0.2, 0.25, 0.3])
        assert np.array_equal(pv[i].cpu().numpy(), want), i
    h = preprocess.collate_fn(items[:1], proc, device=DEV, out_dtype=torch.bfloat16)
    assert h.dtype == torch.bfloat16 and torch.equal(h, pv[:1].to(torch.bfloat16))


def test_unsupported_processors_refuse():
    class BitImageProcessor:          # the name is what processor_kind() dispatches on
        pass
    with pytest.raises(NotImplementedError):
        preprocess.collate_fn([np.zeros((8, 8), np.uint8)], BitImageProcessor(), device=DEV)
