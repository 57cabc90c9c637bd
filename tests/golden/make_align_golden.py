"""Generate tests/golden/align_golden.npz from the installed transformers Dinov2Encoder.

    python tests/golden/make_align_golden.py

The AlignTransformer body is third-party code (transformers modeling_dinov2, see oracle/align.py);
the fixture holds its fp64 output on seeded inputs / weights (radzero_b200.synthetic) so that the
oracle and the CUDA path can be checked where transformers' module is not trusted or not present.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from radzero_b200 import synthetic  # noqa: E402

CASES = {"small": (2, 70, 11), "ragged": (1, 137, 12)}   # name -> (B, L, seed)


def main():
    out = {}
    for name, (B, L, seed) in CASES.items():
        enc = synthetic.build_align_encoder(seed=seed).double()
        tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0]
        with torch.no_grad():
            y = enc(tok.double())["last_hidden_state"]
        out[f"{name}.meta"] = np.array([B, L, seed])
        out[f"{name}.checksum"] = np.array(synthetic.checksum(tok) + synthetic.checksum(
            synthetic.align_layer_weights(seed)[1]["mlp.fc2.weight"]))
        out[f"{name}.out"] = y.float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "align_golden.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
