"""The preprocessing oracle (oracle/preprocess.py) against the golden outputs of the reference's own chain
(cv2.normalize + Pillow bicubic + transformers BlipImageProcessor, tests/golden/make_preprocess_golden.py)
and, where those libraries are installed, against them run live.  Everything is BIT-exact."""
import os

import numpy as np
import pytest

from oracle import preprocess as P
from tests.preprocess_cases import CASES, make_raw

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_golden.npz")))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden(name):
    spec = CASES[name]
    raw = make_raw(spec)
    r = P.resize_bicubic_u8(P.minmax_to_uint8(raw), spec["size"])
    assert np.array_equal(r, G[f"{name}.resized_u8"])
    pv = P.preprocess_image(raw, spec["size"])
    assert pv.shape == (3,) + tuple(spec["size"]) and pv.dtype == np.float32
    s = np.array([pv.astype(np.float64).sum(), np.abs(pv.astype(np.float64)).sum()])
    assert np.array_equal(s, G[f"{name}.pv_sum"])
    if spec.get("keep_pv"):
        assert np.array_equal(pv, G[f"{name}.pixel_values"])


def test_constant_image_is_all_zero_level():
    pv = P.preprocess_image(make_raw(CASES["constant"]), CASES["constant"]["size"])
    lut = P.normalize_lut()
    for c in range(3):
        assert (pv[c] == lut[c][0]).all()          # max == min: cv2 maps everything to the lower bound


@pytest.mark.parametrize("name", ["cxr1024_u8", "dicom_u16", "rgb_u8", "float_identity_w"])
def test_oracle_matches_the_libraries_live(name):
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    try:
        from transformers import BlipImageProcessorPil as Proc
    except ImportError:                       # transformers 4.x: the PIL-backed class is BlipImageProcessor
        from transformers import BlipImageProcessor as Proc
    spec = CASES[name]
    raw = make_raw(spec)
    u8 = cv2.normalize(np.array(raw), None, 0, 255, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_8U)
    assert np.array_equal(P.minmax_to_uint8(raw), u8)
    proc = Proc(size={"height": spec["size"][0], "width": spec["size"][1]})
    want = np.array(proc([Image.fromarray(u8)])["pixel_values"])[0]
    assert np.array_equal(P.preprocess_image(raw, spec["size"]), want)


def test_tap_table_shapes():
    xmin, cnt, kk = P.bicubic_coeffs(1024, 518)
    assert kk.shape == (518, 9) and int(cnt.max()) <= 9 and int(xmin.min()) == 0
    # each row of taps sums to 2^22 up to rounding of the individual weights
    assert np.abs(kk.sum(1) - (1 << 22)).max() <= 9
