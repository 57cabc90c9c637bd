"""Helpers to read ``tests/golden/vlcabs_golden.npz`` and rebuild its seeded inputs."""
import os

import numpy as np
import torch

from radzero_b200 import synthetic

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vlcabs_golden.npz")
_G = {}


def golden():
    if "g" not in _G:
        _G["g"] = dict(np.load(GOLDEN_PATH))
    return _G["g"]


def T(name, dtype=None, device="cpu"):
    t = torch.from_numpy(np.asarray(golden()[name]))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device)


def case_inputs(name, dtype=torch.float32, device="cpu"):
    """Rebuild (tokens, text, gamma, beta, log_tau, counts) of a golden case from its seed.

    Fails loudly if the RNG stream differs from the one the fixture was generated with.
    """
    g = golden()
    seed, B, L, N = [int(v) for v in g[f"{name}.meta"]]
    counts = [int(c) for c in g[f"{name}.counts"]]
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, tokens_per_image=L, seed=seed)
    got = np.array(synthetic.checksum(tok) + synthetic.checksum(text))
    want = g[f"{name}.checksum"]
    if not np.allclose(got, want, rtol=1e-9, atol=1e-6):
        raise RuntimeError(f"golden case {name}: seeded inputs drifted ({got} vs {want}); "
                           "regenerate with tests/golden/make_golden.py")
    cast = lambda t: t.to(dtype).to(device)
    return cast(tok), cast(text), cast(gamma), cast(beta), log_tau.to(device), counts


def split(text, counts):
    out, o = [], 0
    for c in counts:
        out.append(text[o:o + c])
        o += c
    return out
