"""Generate tests/golden/preprocess_golden.npz: outputs of the reference's preprocessing chain
(exp/cxr_pt/inference/dataset.py:31-51 + the BlipImageProcessor of exp/cxr_pt/model/processing.py:85-101)
run with the reference's own dependencies -- cv2.normalize, Pillow's bicubic resize and transformers'
PIL-backed BlipImageProcessor -- on seeded synthetic images.  Run in the build container:

    python tests/golden/make_preprocess_golden.py

The raw images are regenerated from their seeds by tests (tests/preprocess_cases.py); stored here are the
uint8 planes after min-max + resize (exact integers) and, for the level-sweep case, the final float32
pixel_values (every entry of the rescale + normalise table)."""
import os
import sys

import cv2
import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.preprocess_cases import CASES, make_raw  # noqa: E402


def reference_chain(raw, size):
    from transformers import BlipImageProcessorPil
    proc = BlipImageProcessorPil(size={"height": size[0], "width": size[1]})
    img = Image.fromarray(cv2.normalize(np.array(raw), None, 0, 255, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_8U))
    pv = np.array(proc([img])["pixel_values"])[0]
    resized = np.array(img.convert("RGB").resize((size[1], size[0]), resample=Image.BICUBIC))
    return pv, resized


def main():
    out = {}
    for name, spec in CASES.items():
        raw = make_raw(spec)
        pv, resized = reference_chain(raw, spec["size"])
        assert pv.shape == (3,) + tuple(spec["size"]) and pv.dtype == np.float32
        if raw.ndim == 2:
            assert (resized[..., 0] == resized[..., 1]).all() and (resized[..., 0] == resized[..., 2]).all()
            resized = resized[..., 0]
        out[f"{name}.resized_u8"] = resized
        out[f"{name}.pv_sum"] = np.array([pv.astype(np.float64).sum(), np.abs(pv.astype(np.float64)).sum()])
        if spec.get("keep_pv"):
            out[f"{name}.pixel_values"] = pv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(CASES), "cases; cv2", cv2.__version__,
          "PIL", Image.__version__ if hasattr(Image, "__version__") else "")


if __name__ == "__main__":
    main()
