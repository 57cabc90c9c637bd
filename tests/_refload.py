"""Load the reference's ``exp/cxr_pt/model/losses.py`` by file path (test-only).

The file is pure torch except for ``from open_clip.loss import ClipLoss, SigLipLoss``
(losses.py:7), used only by two ablation wrappers that are off the hot path; a stub
module satisfies the import.  The package ``exp.cxr_pt.model`` itself cannot be imported
here (peft / transformers-4.39 internals missing), see SURVEY.md section 0.5.
Nothing from the reference is copied into this repo; this only executes it in place.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RZ_REFERENCE_ROOT", "/root/reference")
_CACHE = {}


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "exp/cxr_pt/model/losses.py"))


def load_reference_losses():
    if "mod" in _CACHE:
        return _CACHE["mod"]
    if not reference_present():
        return None
    import torch.nn as nn
    if "open_clip" not in sys.modules:
        pkg = types.ModuleType("open_clip")
        sub = types.ModuleType("open_clip.loss")

        class ClipLoss(nn.Module):  # placeholder base classes only
            def __init__(self, **kw):
                super().__init__()

        class SigLipLoss(nn.Module):
            def __init__(self, **kw):
                super().__init__()

        sub.ClipLoss, sub.SigLipLoss = ClipLoss, SigLipLoss
        pkg.loss = sub
        sys.modules["open_clip"] = pkg
        sys.modules["open_clip.loss"] = sub
    path = os.path.join(REFERENCE_ROOT, "exp/cxr_pt/model/losses.py")
    spec = importlib.util.spec_from_file_location("_radzero_reference_losses", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHE["mod"] = mod
    return mod


def make_reference_loss(mod, gamma, beta, log_tau, **cfg):
    """Instantiate the reference RadZeroLoss with the given LN / temperature parameters."""
    import torch
    kw = dict(hidden_dim=gamma.numel(), sim_op="cos", loss_temperature=0.07)
    kw.update(cfg)
    loss = mod.RadZeroLoss(**kw).to(gamma.dtype)
    with torch.no_grad():
        loss.layer_norm.weight.copy_(gamma)
        loss.layer_norm.bias.copy_(beta)
        loss.loss_temperature.copy_(log_tau.to(gamma.dtype))
    loss.compute_i2t_loss = False  # attribute the shipped code forgets
    return loss


def text_callback(feature_list):
    """A ``forward_text_model`` stand-in: key_phrases are indices into feature_list."""
    def fwd(kp):
        return {"text_features_wo_l2_norm": feature_list[kp]}
    return fwd


class BlipImageProcessor:  # stand-ins: only isinstance() dispatch is exercised
    pass


class AspectRatioBlipImageProcessor(BlipImageProcessor):
    pass


class BitImageProcessor:
    pass


class M3AEImageProcessor:
    pass


PROCESSOR_KINDS = {"blip": BlipImageProcessor(), "aspect_blip": AspectRatioBlipImageProcessor(),
                   "bit": BitImageProcessor(), "m3ae": M3AEImageProcessor()}


def load_reference_function(rel_path: str, func_name: str, extra_globals=None):
    """Execute ONE function definition of a reference file in place (test-only).

    Used for ``interpolate_similarity_scores`` / ``get_grounding_point``, whose modules
    import packages that are not installed here (pydicom, cv2, torchmetrics).  The
    function's source is read from the checkout with ``ast`` and compiled into a scratch
    namespace that provides stand-ins for the image-processor classes it dispatches on.
    """
    import ast
    import numpy as np
    import torch
    import torch.nn.functional as F
    path = os.path.join(REFERENCE_ROOT, rel_path)
    if not os.path.isfile(path):
        return None
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == func_name)
    code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")

    ns = {"torch": torch, "F": F, "np": np,
          "BlipImageProcessor": BlipImageProcessor,
          "AspectRatioBlipImageProcessor": AspectRatioBlipImageProcessor,
          "BitImageProcessor": BitImageProcessor, "M3AEImageProcessor": M3AEImageProcessor}
    ns.update(extra_globals or {})
    exec(code, ns)
    return ns[func_name], PROCESSOR_KINDS
