"""ctypes binding of librz_b200.so -- the C ABI declared in include/rz_b200.h.

There is no CPU fallback: if the shared library is missing (and cannot be built because
nvcc is absent) every op raises.  The library is built in-tree by ``radzero_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librz_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "rz_b200.h")

RZ_OK = 0
RZ_F32, RZ_BF16, RZ_F16 = 0, 1, 2
RZ_UP_RAW, RZ_UP_SIGMOID, RZ_UP_MASK, RZ_UP_ARGMAX, RZ_UP_MASK_BITS = 0, 1, 2, 3, 4
RZ_LIN_BIAS, RZ_LIN_GELU, RZ_LIN_RESIDUAL, RZ_LIN_RESIDUAL_F16 = 0, 1, 2, 3
RZ_IMG_U8, RZ_IMG_U16, RZ_IMG_I16, RZ_IMG_I32, RZ_IMG_F32 = 0, 1, 2, 3, 4

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_ull, _u = C.c_ulonglong, C.c_uint

# name -> (restype, argtypes); must list every function include/rz_b200.h declares
SIGNATURES: Dict[str, tuple] = {
    "rz_version": (_i, []),
    "rz_strerror": (C.c_char_p, [_i]),
    "rz_last_cuda_error": (C.c_char_p, []),
    "rz_launch_count": (_ll, []),
    "rz_device_sm_count": (_i, []),
    "rz_prep_rows": (_i, [_vp, _i, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "rz_sim_fwd": (_i, [_vp, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _ll, _ll, _i, _vp, _ll, _ll,
                        _f, _vp, _i, _vp, _vp, _vp, _vp]),
    "rz_sim_fwd_tokens_workspace_bytes": (C.c_size_t, [_i, _i]),
    "rz_sim_fwd_tokens": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _ll, _ll,
                               _i, _vp, _ll, _ll, _f, _vp, _i, _vp, _vp, C.c_size_t, _vp]),
    "rz_sim_fwd_large_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "rz_sim_fwd_large": (_i, [_vp, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _ll, _ll, _i, _vp, _ll, _ll,
                              _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, C.c_size_t, _vp]),
    "rz_sim_bwd_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "rz_sim_bwd": (_i, [_vp, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                        _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "rz_prep_rows_bwd_blocks": (_i, [_ll]),
    "rz_prep_rows_bwd": (_i, [_vp, _i, _vp, _vp, _ll, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _f, _vp]),
    "rz_upsample_maps": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _f, _vp, _vp]),
    "rz_map_threshold_stats": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "rz_mpnce_partials_scratch_floats": (C.c_size_t, [_i, _i]),
    "rz_mpnce_finish_scratch_floats": (C.c_size_t, [_i, _i, _i]),
    "rz_mpnce_partials": (_i, [_vp, _ll, _i, _i, _i, _vp, _i, _f, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "rz_mpnce_finish": (_i, [_vp, _ll, _i, _i, _i, _vp, _i, _f, _vp, _f, _i, _i, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp]),
    "rz_group_map": (_i, [_vp, _i, _ll, _vp, _vp]),
    "rz_preprocess_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i, _i, _i]),
    "rz_preprocess_images": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, C.c_double, _vp, _i, _vp,
                                  C.c_size_t, _vp]),
    "rz_text_pool": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "rz_ln_rows": (_i, [_vp, _i, _vp, _vp, _f, _ll, _vp, _vp]),
    "rz_linear": (_i, [_vp, _ll, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "rz_attention": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "rz_grad_scale_floats": (C.c_size_t, []),
    "rz_grad_scale": (_i, [_vp, _ll, _vp, _vp]),
    "rz_ls_cast_bwd": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp]),
    "rz_ls_weight_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "rz_transpose_pad": (_i, [_vp, _ll, _i, _ll, _vp, _vp, _vp, _vp]),
    "rz_gelu_bwd": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "rz_ln_rows_bwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp]),
    "rz_attention_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "rz_umma_probe": (_i, [_vp, _i, _vp, _i, _ull, _ull, _i, _i, _i, _u, _u, _i, _vp, _vp]),
}

_lib = None


class RzError(RuntimeError):
    pass


def declared_symbols() -> List[str]:
    """Function names declared in include/rz_b200.h (used by the symbol-export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rz_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True):
    """Load the shared library, building it first when it is absent and nvcc exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RzError(f"{LIB_PATH} is missing; run `python -m radzero_b200.build`")
        from . import build as _build
        _build.build()          # serialised across processes by a file lock (torchrun ranks)
    # RZ_B200_LIB: load another build of the same ABI (kernel-variant experiments under profiles/experiments/)
    lib = C.CDLL(os.environ.get("RZ_B200_LIB") or LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != RZ_OK:
        lib = load()
        msg = lib.rz_strerror(rc).decode()
        cu = lib.rz_last_cuda_error().decode()
        raise RzError(f"{what or 'rz call'} failed: {msg}" + (f" [{cu}]" if cu else ""))


def launch_count() -> int:
    return int(load().rz_launch_count())
