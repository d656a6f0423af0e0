"""ctypes binding of libkoa_b200.so (the C ABI declared in include/koa_b200.h).

The library is the product: there is no Python/PyTorch fallback for any compute entry point.
If the shared object is missing the import fails loudly; run ``python -c "import
__graft_entry__ as g; g.build()"`` (or ``make -C oaprogressionmmf_b200/csrc``) first.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkoa_b200.so")


class KoaError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """Mirror of ``koa_epilogue_t``."""

    _fields_ = [
        ("out", C.c_void_p),
        ("ldo", C.c_int),
        ("out_fp32", C.c_int),
        ("act", C.c_int),
        ("bias", C.c_void_p),
        ("pre_out_bf16", C.c_void_p),
        ("aux_bf16", C.c_void_p),
        ("residual_f32", C.c_void_p),
        ("add_bf16", C.c_void_p),
        ("mask_bf16", C.c_void_p),
        ("out_bf16_copy", C.c_void_p),
        ("col_sum", C.c_void_p),
        ("col_sumsq", C.c_void_p),
    ]


ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2, 3

_P = C.c_void_p
_I = C.c_int

# name -> (restype, argtypes). Every symbol declared in include/koa_b200.h must appear here;
# tests/test_abi.py cross-checks the two lists.
SIGNATURES = {
    "koa_last_error": (C.c_char_p, []),
    "koa_version": (_I, []),
    "koa_debug_flag": (_I, [C.POINTER(C.c_uint)]),
    "koa_debug_set_wgrad_desc": (_I, [C.c_uint, C.c_uint, C.c_uint]),
    "koa_gemm_bf16": (_I, [_P, _P, _I, _I, _I, C.POINTER(Epilogue), _P]),
    "koa_conv_fprop_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, C.POINTER(Epilogue), _P]),
    "koa_gemm_wgrad_bf16": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "koa_conv_wgrad_bf16": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KoaError(
            f"{LIB_PATH} not found: the CUDA extension is not built. There is no CPU or PyTorch "
            "fallback for this path; build it with __graft_entry__.build()."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().koa_last_error()
        raise KoaError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def debug_flag() -> int:
    v = C.c_uint(0)
    check(load().koa_debug_flag(C.byref(v)), "koa_debug_flag")
    return int(v.value)


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
