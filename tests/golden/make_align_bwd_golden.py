"""Generate tests/golden/align_bwd_golden.npz: autograd of the installed transformers Dinov2Encoder.

    python tests/golden/make_align_bwd_golden.py

The AlignTransformer body is third-party code (transformers modeling_dinov2, see oracle/align.py); the
fixture holds, in fp64 arithmetic on seeded inputs / weights (radzero_b200.synthetic), the gradients of
``loss = sum(encoder(tokens) * R)``: dL/dtokens in full, every vector parameter gradient in full and, for
the weight matrices (7 M elements a layer), their first four rows, their Frobenius norm and a fixed random
projection ``sum(dW * P)`` -- enough to pin every element's contribution without a 100 MB file.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from radzero_b200 import synthetic  # noqa: E402

CASES = {"small": (2, 70, 21, 1.0), "ragged_tiny_grad": (1, 137, 22, 1e-6)}   # name -> (B, L, seed, grad magnitude)


def upstream(B, L, seed, mag):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(B, L, 768, generator=g, dtype=torch.float64) * mag


def projection(shape, seed):
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64)


def main():
    out = {}
    for name, (B, L, seed, mag) in CASES.items():
        enc = synthetic.build_align_encoder(seed=seed).double().train()
        tok = synthetic.make_inputs(B, 1, tokens_per_image=L, seed=seed)[0].double().requires_grad_(True)
        y = enc(tok)["last_hidden_state"]
        (y * upstream(B, L, seed, mag)).sum().backward()
        out[f"{name}.meta"] = np.array([B, L, seed, mag])
        out[f"{name}.dtokens"] = tok.grad.numpy()
        for pname, p in enc.named_parameters():
            g = p.grad
            if g.dim() == 1:
                out[f"{name}.{pname}"] = g.numpy()
            else:
                out[f"{name}.{pname}.rows"] = g[:4].numpy()
                out[f"{name}.{pname}.norm"] = np.array(g.norm().item())
                out[f"{name}.{pname}.proj"] = np.array((g * projection(g.shape, seed)).sum().item())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "align_bwd_golden.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
