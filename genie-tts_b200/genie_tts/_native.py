"""ctypes binding of libgenie_b200.so (include/genie_b200.h).

The product path has no CPU fallback: if the shared library is missing or no
B200 is visible, every compute entry point raises ``GenieNativeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "csrc", "libgenie_b200.so"))

GRAPH_T2S_ENCODER, GRAPH_T2S, GRAPH_VITS, GRAPH_PROMPT_ENCODER = 0, 1, 2, 3
F32, F16 = 0, 1
CANCELLED = 2


class GenieNativeError(RuntimeError):
    pass


class Sampling(C.Structure):
    _fields_ = [("top_k", C.c_int), ("temperature", C.c_float), ("repetition_penalty", C.c_float),
                ("greedy", C.c_int), ("seed", C.c_ulonglong), ("max_steps", C.c_int), ("fixed_steps", C.c_int),
                ("top_p", C.c_float)]


_lib: Optional[C.CDLL] = None

# symbol -> (restype, argtypes); must list every function include/genie_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "genie_last_error": (C.c_char_p, []),
    "genie_version": (C.c_int, []),
    "genie_launch_count": (C.c_ulonglong, []),
    "genie_device_count": (C.c_int, []),
    "genie_host_alloc": (C.c_int, [C.c_size_t, _P]),
    "genie_host_free": (C.c_int, [_P]),
    "genie_model_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "genie_model_add_tensor": (C.c_int, [_P, C.c_int, C.c_char_p, _P, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "genie_model_set_constants": (C.c_int, [_P, _P, C.c_int, C.c_float, C.c_float, C.c_float]),
    "genie_model_finalize": (C.c_int, [_P]),
    "genie_model_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "genie_model_destroy": (None, [_P]),
    "genie_context_create": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "genie_set_stream": (C.c_int, [_P, _P]),
    "genie_prompt_create": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, C.POINTER(_P)]),
    "genie_prompt_create_with_ge": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, C.POINTER(_P)]),
    "genie_prompt_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "genie_prompt_read": (C.c_int, [_P, _P, _P, _P]),
    "genie_prompt_destroy": (None, [_P]),
    "genie_t2s_generate": (C.c_int, [_P, C.POINTER(_P), C.c_int, _P, _P, _P, C.POINTER(Sampling), _P, C.c_int,
                                     _P, C.c_int, _P, _P]),
    "genie_t2s_prefill": (C.c_int, [_P, C.POINTER(_P), C.c_int, _P, _P, _P, C.POINTER(Sampling), C.c_int]),
    "genie_t2s_decode_steps": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "genie_t2s_read": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "genie_t2s_pool_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "genie_t2s_pool_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "genie_t2s_admit": (C.c_int, [_P, C.c_int, _P, C.POINTER(_P), _P, _P, _P, _P]),
    "genie_t2s_pool_step": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "genie_t2s_pool_poll": (C.c_int, [_P, _P, _P, C.c_int]),
    "genie_t2s_pool_read": (C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "genie_t2s_release": (C.c_int, [_P, C.c_int]),
    "genie_vits_decode": (C.c_int, [_P, C.POINTER(_P), C.c_int, _P, _P, _P, _P, _P, C.c_ulonglong, C.c_float,
                                    C.c_int, _P, _P, _P]),
    "genie_debug_record_logits": (C.c_int, [_P, C.c_int]),
    "genie_debug_read_logits": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int)]),
    "genie_debug_read": (C.c_int, [_P, C.c_char_p, _P, C.c_longlong, C.POINTER(C.c_longlong)]),
    "genie_debug_keep": (C.c_int, [_P, C.c_int]),
    "genie_debug_sample": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, C.POINTER(Sampling), _P, C.c_int, _P, _P]),
    "genie_debug_tc_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "genie_profiler_range": (C.c_int, [C.c_int]),
    "genie_last_timing": (C.c_int, [_P, _P, C.c_int]),
    "genie_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GenieNativeError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(genie_b200 has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().genie_last_error()
        raise GenieNativeError((msg or b"unknown error").decode(errors="replace"))


def require_gpu() -> None:
    if lib().genie_device_count() <= 0:
        raise GenieNativeError("no CUDA device visible: genie_b200 runs on B200 (sm_100a) only, no CPU fallback")
