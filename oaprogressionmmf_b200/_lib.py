"""ctypes binding of libkoa_b200.so (the C ABI declared in include/koa_b200.h).

The library is the product: there is no Python/PyTorch fallback for any compute entry point.
If the shared object is missing the import fails loudly; run ``python -c "import
__graft_entry__ as g; g.build()"`` (or ``make -C oaprogressionmmf_b200/csrc``) first.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KOA_LIB") or os.path.join(_HERE, "libkoa_b200.so")  # KOA_LIB: A/B builds of the same ABI


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
        ("gate_bf16", C.c_void_p),
        ("out_bf16_copy", C.c_void_p),
        ("col_sum", C.c_void_p),
        ("col_sumsq", C.c_void_p),
        ("stat_y", C.c_void_p),
        ("stat_mean", C.c_void_p),
        ("stat_invstd", C.c_void_p),
        ("a_f16", C.c_int),
        ("b_f16", C.c_int),
        ("out_f16", C.c_int),
        ("act_f16", C.c_int),
        ("drop_p", C.c_float),
        ("drop_site", C.c_uint),
        ("drop_seed", C.c_ulonglong),
        ("bn_scale", C.c_void_p),
        ("bn_shift", C.c_void_p),
        ("res_scale", C.c_void_p),
        ("res_shift", C.c_void_p),
        ("col_bias", C.c_void_p),
    ]


class FeDesc(C.Structure):
    """Mirror of ``koa_fe_desc_t``."""

    _fields_ = [
        ("arch", C.c_int),
        ("n_img", C.c_int),
        ("h", C.c_int),
        ("w", C.c_int),
        ("slices", C.c_int),
        ("with_gap", C.c_int),
        ("training", C.c_int),
        ("need_backward", C.c_int),
        ("input_for_backward", C.c_void_p),
    ]


class FeatDesc(C.Structure):
    """Mirror of ``koa_feat_desc_t``."""

    _fields_ = [
        ("batch", C.c_int),
        ("n_patches", C.c_int),
        ("dim", C.c_int),
        ("depth", C.c_int),
        ("heads", C.c_int),
        ("mlp_dim", C.c_int),
        ("num_classes", C.c_int),
        ("with_cls", C.c_int),
        ("compute_head", C.c_int),
        ("training", C.c_int),
        ("need_backward", C.c_int),
        ("emb_dropout", C.c_float),
        ("mlp_dropout", C.c_float),
        ("seed", C.c_ulonglong),
    ]


class AdamTensor(C.Structure):
    """Mirror of ``koa_adam_tensor_t``."""

    _fields_ = [
        ("param", C.c_void_p),
        ("grad", C.c_void_p),
        ("exp_avg", C.c_void_p),
        ("exp_avg_sq", C.c_void_p),
        ("numel", C.c_longlong),
    ]


class AdamHyper(C.Structure):
    """Mirror of ``koa_adam_hyper_t``."""

    _fields_ = [
        ("lr", C.c_double),
        ("beta1", C.c_double),
        ("beta2", C.c_double),
        ("eps", C.c_double),
        ("weight_decay", C.c_double),
        ("grad_scale", C.c_double),
        ("step", C.c_int),
        ("decoupled_weight_decay", C.c_int),
    ]


class Augment(C.Structure):
    """Mirror of ``koa_augment_t``."""

    _fields_ = [
        ("off0", C.c_int),
        ("off1", C.c_int),
        ("off2", C.c_int),
        ("rotate", C.c_int),
        ("cos_t", C.c_float),
        ("sin_t", C.c_float),
        ("inv_gamma", C.c_float),
        ("lo", C.c_float),
        ("range", C.c_float),
        ("flip", C.c_int),
    ]


ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2, 3
DT_F32, DT_U8, DT_U16, DT_I16 = 0, 1, 2, 3
ARCH_IDS = {"resnet18": 0, "resnet34": 1, "resnet50": 2, "resnext50_32x4d": 3}

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_PP = C.POINTER(C.c_void_p)
_SZ = C.POINTER(C.c_size_t)

# name -> (restype, argtypes). Every symbol declared in include/koa_b200.h must appear here;
# tests/test_abi.py cross-checks the two lists.
SIGNATURES = {
    "koa_last_error": (C.c_char_p, []),
    "koa_version": (_I, []),
    "koa_debug_flag": (_I, [C.POINTER(C.c_uint)]),
    "koa_debug_flag_peek": (_I, [C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "koa_debug_set_wgrad_desc": (_I, [C.c_uint, C.c_uint, C.c_uint]),
    "koa_launch_count": (C.c_longlong, []),
    "koa_profile_enable": (_I, [_I]),
    "koa_profile_read": (_I, [C.POINTER(C.c_double)]),
    "koa_profile_dump": (_I, [C.c_char_p]),
    "koa_profile_pending": (_I, [C.c_char_p, _I]),
    "koa_gemm_bf16": (_I, [_P, _P, _I, _I, _I, C.POINTER(Epilogue), _P]),
    "koa_gemm_kcat_bf16": (_I, [_P, _I, _P, _I, _P, _I, _I, C.POINTER(Epilogue), _P]),
    "koa_conv_fprop_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, C.POINTER(Epilogue), _P]),
    "koa_gemm_wgrad_bf16": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "koa_conv_wgrad_bf16": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "koa_fe_workspace_bytes": (C.c_size_t, [C.POINTER(FeDesc)]),
    "koa_fe_out_shape": (_I, [C.POINTER(FeDesc), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "koa_fe_num_units": (_I, [C.POINTER(FeDesc)]),
    "koa_fe_forward": (_I, [C.POINTER(FeDesc), _PP, _P, _P, _P, _P]),
    "koa_fe_backward": (_I, [C.POINTER(FeDesc), _PP, _PP, _P, _P, _P]),
    "koa_fe_backward_range": (_I, [C.POINTER(FeDesc), _PP, _PP, _P, _P, _I, _I, _I, _P]),
    "koa_fe_num_blocks": (_I, [C.POINTER(FeDesc)]),
    "koa_fe_debug_offset": (_I, [C.POINTER(FeDesc), _I, _I, _SZ, _SZ]),
    "koa_feat_workspace_bytes": (C.c_size_t, [C.POINTER(FeatDesc)]),
    "koa_feat_num_params": (_I, [C.POINTER(FeatDesc)]),
    "koa_feat_probs_offset": (_I, [C.POINTER(FeatDesc), _I, _SZ, _SZ]),
    "koa_feat_forward": (_I, [C.POINTER(FeatDesc), _PP, _P, _P, _P, _P, _P]),
    "koa_feat_backward": (_I, [C.POINTER(FeatDesc), _PP, _PP, _P, _P, _P, _P, _P]),
    "koa_linear_small_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "koa_linear_small_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "koa_focal_loss": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "koa_layernorm_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "koa_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "koa_attention_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "koa_attention_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "koa_attention_fwd_fmt": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "koa_attention_bwd_fmt": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "koa_stem_pack": (_I, [_P, _P, _I, _I, _I, _P]),
    "koa_maxpool_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "koa_maxpool_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "koa_col_stats": (_I, [_P, _P, _P, C.c_longlong, _I, _P]),
    "koa_channel_dropout": (_I, [_P, _P, C.c_longlong, _I, _I, C.c_ulonglong, C.c_uint, _F, _P]),
    "koa_dropout_mask": (_I, [C.c_ulonglong, C.c_uint, C.c_longlong, _I, _F, _P, _P]),
    "koa_adam_step": (_I, [_P, _I, C.POINTER(AdamHyper), _P]),
    "koa_resample_linear": (_I, [_P, _I, _P, _I, C.POINTER(_I), C.POINTER(_I), _P, _P, _P]),
    "koa_unit_range_affine": (_I, [_P, _I, _I, C.c_longlong, _F, _F, _P, _P, _P, _P, _P]),
    "koa_augment_resample": (_I, [_P, _I, _P, _P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _F, _F, _P, _P]),
    "koa_predict": (_I, [_P, _P, _P, _I, _I, _P]),
    "koa_ensemble_proba": (_I, [_P, _P, _P, _I, _I, _I, _P]),
}


def ptr_table(tensors):
    """ctypes void* array of device pointers (None -> NULL)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr

def zeros_like_flat(tensors):
    """Zero-initialised gradient buffers for ``tensors`` (None entries stay None) carved out of ONE flat
    allocation: a single memset instead of one fill kernel per parameter, and one contiguous buffer per engine call
    for the data-parallel gradient all-reduce (dataparallel.sync_flat). Returns (views, flat)."""
    import torch

    live = [t for t in tensors if t is not None]
    if not live:
        return [None] * len(tensors), None
    sizes = [(t.numel() + 63) // 64 * 64 for t in live]  # 256-byte aligned slices
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=live[0].device)
    out, off, it = [], 0, iter(sizes)
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        n = next(it)
        out.append(flat[off:off + t.numel()].view(t.shape))
        off += n
    return out, flat


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


def require_cuda(t, what: str) -> None:
    """The product has no CPU path: a host tensor is an error, never a silent fallback."""
    if not t.is_cuda:
        raise KoaError(f"{what} needs a CUDA tensor: this path has no CPU fallback")


def require_same_device(what: str, *tensors):
    """Every tensor argument of a C call lives on ONE CUDA device (None entries are skipped); returns that device.
    A host tensor would hand the kernels a host pointer (a sticky illegal-address fault, not an exception); a tensor on
    another GPU would be read from the wrong device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        require_cuda(t, what)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise KoaError(f"{what}: tensors on different devices ({dev} and {t.device})")
    if dev is None:
        raise KoaError(f"{what}: no tensor argument")
    return dev


def on_device(device):
    """Context manager making ``device`` the current CUDA device for the launches of a C call."""
    import torch

    return torch.cuda.device(device)


def debug_flag() -> int:
    v = C.c_uint(0)
    check(load().koa_debug_flag(C.byref(v)), "koa_debug_flag")
    return int(v.value)


def debug_flag_peek() -> tuple[int, int]:
    """(latest, first) barrier time-out codes without waiting for the device and without clearing them."""
    v, f = C.c_uint(0), C.c_uint(0)
    check(load().koa_debug_flag_peek(C.byref(v), C.byref(f)), "koa_debug_flag_peek")
    return int(v.value), int(f.value)


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
