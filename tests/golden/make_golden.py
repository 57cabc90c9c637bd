"""Generate the golden fixtures from the REFERENCE itself (run in the build container).

    python tests/golden/make_golden.py

Executes the reference's own ``RadZeroLoss`` / ``SimilarityLogit`` /
``multi_positive_nce_loss`` (exp/cxr_pt/model/losses.py), ``interpolate_similarity_scores``
(exp/cxr_pt/inference/segmentation_utils.py:36-122) and ``get_grounding_point``
(exp/cxr_pt/inference/grounding_utils.py:166-261) in place, in fp64 on fp32-representable
seeded inputs, and freezes the outputs into ``tests/golden/vlcabs_golden.npz``.  The
reference has no tests or golden vectors of its own; these files are
what pins the oracle -- and through it the CUDA path -- on machines without the checkout.
Inputs of the large case are regenerated from the seed at test time and guarded by a
checksum, so the fixture stays small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from radzero_b200 import synthetic  # noqa: E402
from tests._refload import (PROCESSOR_KINDS, load_reference_function, load_reference_losses,  # noqa: E402
                            make_reference_loss, text_callback)

CASES = {
    # name: (seed, B, counts, L)
    "small": (7, 3, [2, 1, 3], 50),
    "full": (8, 2, [1, 2], 1370),
    "cls14": (9, 2, [1] * 14, 1370),
}
UPSAMPLE_SIZES = {"s64x80": ((64, 80), 1), "s518": ((518, 518), 5), "s1024": ((1024, 1024), 9),
                  "s300x417": ((300, 417), 3)}


def split(text, counts):
    out, o = [], 0
    for c in counts:
        out.append(text[o:o + c])
        o += c
    return out


def main():
    mod = load_reference_losses()
    assert mod is not None, "reference checkout required"
    out = {}
    for name, (seed, B, counts, L) in CASES.items():
        tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts), tokens_per_image=L, seed=seed)
        out[f"{name}.meta"] = np.array([seed, B, L, sum(counts)], dtype=np.int64)
        out[f"{name}.counts"] = np.array(counts, dtype=np.int64)
        out[f"{name}.checksum"] = np.array(synthetic.checksum(tok) + synthetic.checksum(text), dtype=np.float64)
        if name == "small":
            out["small.tokens"] = tok.numpy()
            out["small.text"] = text.numpy()
            out["small.gamma"] = gamma.numpy()
            out["small.beta"] = beta.numpy()
        t64 = text.double().clone().requires_grad_(True)
        x64 = tok.double().clone().requires_grad_(True)
        ref = make_reference_loss(mod, gamma.double(), beta.double(), log_tau.double())
        r = ref(list(range(len(counts))) if name != "cls14" else list(range(len(counts))),
                x64, text_callback(split(t64, counts)), ddp_gather=False,
                need_attn_weights=True, compute_loss=(name != "cls14"))
        z = r["t2i_logits"].detach()
        s = r["t2i_attn_weights"][0].detach()
        out[f"{name}.t2i_logits"] = z.numpy()
        # compute_logits glue run on the reference outputs (modeling.py:311-328)
        sim = torch.mean(torch.stack(r["t2i_attn_weights"]), dim=0)[:, :, 1:].detach()
        logits = (z.T / ref.loss_temperature.exp()).detach()
        out[f"{name}.logits"] = logits.numpy()
        if name == "small":
            out["small.scores"] = s.numpy()
            out["small.similarity_scores"] = sim.numpy()
        else:
            out[f"{name}.scores_stride7"] = s[:, :, ::7].numpy()
        if name != "cls14":
            loss = r["losses"]["loss"]
            out[f"{name}.loss"] = np.array(loss.item())
            loss.backward()
            if name == "small":
                out["small.grad_text"] = t64.grad.numpy()
                out["small.grad_tokens"] = x64.grad.numpy()
            else:
                out[f"{name}.grad_text"] = t64.grad.numpy()
                out[f"{name}.grad_tokens_stride13"] = x64.grad[:, ::13].numpy()
            out[f"{name}.grad_gamma"] = ref.layer_norm.weight.grad.numpy()
            out[f"{name}.grad_beta"] = ref.layer_norm.bias.grad.numpy()
            out[f"{name}.grad_log_tau"] = ref.loss_temperature.grad.numpy()

    # MP-NCE alone (+ dZ) in all four flag combinations
    g = torch.Generator().manual_seed(11)
    z = (torch.rand(13, 5, generator=g) * 2 - 1)
    gm = torch.tensor([0, 0, 1, 1, 1, 2, 3, 3, 3, 3, 4, 4, 4])
    out["mpnce.z"] = z.numpy()
    out["mpnce.group_map"] = gm.numpy()
    for rs in (False, True):
        for cs in (False, True):
            zz = z.clone().requires_grad_(True)
            l = mod.multi_positive_nce_loss(zz, gm, temperature=0.07, row_sum=rs, col_sum=cs)
            l.backward()
            out[f"mpnce.loss_r{int(rs)}c{int(cs)}"] = np.array(l.item(), dtype=np.float64)
            out[f"mpnce.dz_r{int(rs)}c{int(cs)}"] = zz.grad.numpy()

    # squeeze quirk shapes (losses.py:229-233)
    sl = mod.SimilarityLogit("cos")
    for (B, N) in [(1, 4), (3, 1), (1, 1), (2, 3)]:
        tok, text, *_ = synthetic.make_inputs(B, N, tokens_per_image=30, seed=1)
        zq, _ = sl(text, tok, temperature=torch.tensor(0.07))
        out[f"quirk.B{B}N{N}.shape"] = np.array(list(zq.shape), dtype=np.int64)

    # bilinear upsample + processor variants + grounding point
    interp, kinds = load_reference_function("exp/cxr_pt/inference/segmentation_utils.py",
                                            "interpolate_similarity_scores")
    gpoint, _ = load_reference_function("exp/cxr_pt/inference/grounding_utils.py", "get_grounding_point")
    g = torch.Generator().manual_seed(4)
    grid = torch.randn(1369, generator=g) * 3.0
    out["upsample.grid"] = grid.numpy()
    for key, (size, stride) in UPSAMPLE_SIZES.items():
        m = interp(grid, size, kinds["blip"])[0]
        out[f"upsample.{key}.size_stride"] = np.array([size[0], size[1], stride], dtype=np.int64)
        out[f"upsample.{key}.map"] = m[::stride, ::stride].numpy()
        out[f"upsample.{key}.sum"] = np.array(m.double().sum().item())
        out[f"upsample.{key}.point"] = np.array(gpoint(grid, size, kinds["blip"]), dtype=np.int64)
        out[f"upsample.{key}.mask_gt0p7_count"] = np.array(int((torch.sigmoid(m) > 0.7).sum()), dtype=np.int64)
    for kind in ("aspect_blip", "bit", "m3ae"):
        for size in ((300, 417), (417, 300)):
            m = interp(grid, size, kinds[kind])[0]
            out[f"upsample.{kind}.{size[0]}x{size[1]}.map"] = m[::3, ::3].numpy()
            out[f"upsample.{kind}.{size[0]}x{size[1]}.point"] = np.array(gpoint(grid, size, kinds[kind]), dtype=np.int64)

    # large float arrays are stored as fp32 (6e-8 relative) to keep the fixture small
    for k, v in list(out.items()):
        if v.dtype == np.float64 and v.size > 4096:
            out[k] = v.astype(np.float32)
    path = os.path.join(HERE, "vlcabs_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
