"""ctypes binding of include/mmcm.h -- the only way the Python host code reaches the GPU kernels.

This is the same stub a maintainer of the reference would add: plain pointers and sizes,
no torch types cross the boundary.  Loading never falls back to anything: a missing library or symbol raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import List, Optional

from . import build as _build

OK, EINVAL, ECUDA, ESTATE = 0, 1, 2, 3

EPI_BIAS_BF16, EPI_BIAS_ACT_BF16, EPI_BIAS_RESID_F32, EPI_PATCH_F32 = 0, 1, 2, 3


class MmcmConfig(C.Structure):
    """Mirror of `struct mmcm_config` (include/mmcm.h)."""
    _fields_ = [
        ("backend", C.c_int32), ("head", C.c_int32),
        ("text_hidden", C.c_int32), ("text_heads", C.c_int32), ("text_layers", C.c_int32),
        ("text_ffn", C.c_int32), ("text_act", C.c_int32),
        ("vis_hidden", C.c_int32), ("vis_heads", C.c_int32), ("vis_layers", C.c_int32),
        ("vis_ffn", C.c_int32), ("vis_act", C.c_int32),
        ("text_eps", C.c_float), ("vis_eps", C.c_float),
        ("vocab", C.c_int32), ("max_pos", C.c_int32), ("eos_id", C.c_int32),
        ("image", C.c_int32), ("patch", C.c_int32), ("proj_dim", C.c_int32),
        ("fusion_dim", C.c_int32), ("num_outputs", C.c_int32), ("head_hidden_dim", C.c_int32),
    ]


_P = C.c_void_p
_SIGS = {
    "mmcm_create": (C.c_int, [C.POINTER(MmcmConfig), C.c_int, C.POINTER(_P)]),
    "mmcm_destroy": (C.c_int, [_P]),
    "mmcm_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "mmcm_finalize_weights": (C.c_int, [_P]),
    "mmcm_save_packed": (C.c_int, [_P, C.c_char_p]),
    "mmcm_load_packed": (C.c_int, [_P, C.c_char_p]),
    "mmcm_packed_config": (C.c_int, [C.c_char_p, _P]),
    "mmcm_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmcm_forward_host": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmcm_forward_u8": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmcm_forward_host_u8": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmcm_prefetch_host": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32]),
    "mmcm_prefetch_host_u8": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32]),
    "mmcm_get_stage": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "mmcm_last_launch_count": (C.c_int64, [_P]),
    "mmcm_last_host_copy_share": (C.c_double, [_P]),
    "mmcm_last_chunks": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mmcm_gemm_time": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mmcm_gemm_time_epi": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_int64)]),
    "mmcm_gemm_time_shape": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mmcm_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "mmcm_last_error": (C.c_char_p, []),
    "mmcm_version": (C.c_char_p, []),
    "mmcm_gemm_bf16": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P,
                                 C.c_int32, C.c_int32, C.c_int32, _P]),
    "mmcm_fold_ln": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P, _P]),
    "mmcm_prep_rows": (C.c_int, [_P, _P, _P, C.c_float, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmcm_gemm_resid_stats": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "mmcm_gemm_lnfold": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P, _P]),
    "mmcm_tokenizer_create": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "mmcm_tokenizer_destroy": (C.c_int, [_P]),
    "mmcm_tokenizer_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]),
    "mmcm_tokenizer_encode": (C.c_int, [_P, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int32, C.c_int32, _P, _P,
                                        C.c_int32]),
    "mmcm_layernorm": (C.c_int, [_P, _P, _P, C.c_float, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmcm_attention": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mmcm_cast_bf16": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P]),
    "mmcm_debug_set_gemm_trace": (C.c_int, [_P]),
    "mmcm_resize_crop_u8": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "mmcm_preprocess_u8": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P]),
    "mmcm_postprocess": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def declared_symbols() -> List[str]:
    """Every function include/mmcm.h declares (parsed from the header, so the check cannot drift)."""
    with open(_build.HEADER) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmcm_[a-z0-9_]+)\s*\(", txt)))


def load() -> C.CDLL:
    """dlopen libmmcm.so (building it first if it is missing or stale) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MMCM_LIB_PATH")      # dev: A/B a differently built library
    if not path:
        path = _build.LIB_PATH
        if _build.is_stale():
            path = _build.build_library()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: the CUDA extension is mandatory (no CPU fallback)")
    lib = C.CDLL(path)
    for name in declared_symbols():
        if not hasattr(lib, name):
            raise RuntimeError(f"libmmcm.so does not export `{name}` declared in include/mmcm.h")
        if name not in _SIGS:
            raise RuntimeError(f"lib.py has no ctypes signature for `{name}` declared in include/mmcm.h")
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _SIGS[name]
    _lib = lib
    return lib


def last_error() -> str:
    return load().mmcm_last_error().decode("utf-8", "replace")


def check(code: int) -> None:
    """Map a C status to the exception type the reference raises for the same condition."""
    if code == OK:
        return
    msg = last_error()
    if code == EINVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)
