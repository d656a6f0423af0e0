"""The "last-chance preprocessing" in front of the models, on the CUDA path.

The reference resamples every modality on the GPU right before the forward pass
(``koafusion/run/train_prog_fus.py:111-116,143-146``: ``PTInterpolate(scale_factor=factor)(x)`` with the per-modality
factors of ``config.model.downscale``, e.g. ``[[0.5,0.5],[0.5,0.5,0.5],[0.5,0.5,1.0],[1.0]]``), on fp32 tensors that its
CPU workers have already mapped to the unit range and z-scored (``koafusion/preproc/_pt.py:75-124``;
``koafusion/datasets/_data_provider.py:297-334``). Here:

* ``PTInterpolate`` / ``downscale_x`` are the drop-ins (same arguments, same result to fp32 round-off) on one
  ``koa_resample_linear`` launch;
* ``unit_range_normalize_downscale`` is the fused form for loaders that ship the volumes in their on-disk integer type
  (uint8 / uint16: a quarter / half of the fp32 host->device bytes): per-volume minimum / maximum on the device
  (``koa_unit_range_affine``), then unit range + z-score + resampling in a single pass over the integers.

No CPU or PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Sequence

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.DT_F32, torch.uint8: _lib.DT_U8, torch.uint16: _lib.DT_U16, torch.int16: _lib.DT_I16}


def output_size(size_in: Sequence[int], scale_factor) -> list:
    """Spatial output size of ``F.interpolate(..., scale_factor=..., recompute_scale_factor=True)``:
    ``floor(size * factor)`` per dimension (a scalar factor applies to every dimension)."""
    if isinstance(scale_factor, (int, float)):
        scale_factor = [scale_factor] * len(size_in)
    scale_factor = list(scale_factor)
    if len(scale_factor) != len(size_in):
        raise ValueError(f"scale_factor {scale_factor} does not match the {len(size_in)} spatial dimensions")
    return [int(math.floor(float(s) * float(f))) for s, f in zip(size_in, scale_factor)]


def _dims3(spatial: Sequence[int]) -> list:
    return [1] * (3 - len(spatial)) + [int(s) for s in spatial]


def _resample(x: torch.Tensor, size_out: Sequence[int], scale=None, shift=None) -> torch.Tensor:
    if x.ndim not in (3, 4, 5):
        raise ValueError(f"expected a (B, CH, D0, ...) tensor with 3 to 5 dimensions, got {tuple(x.shape)}")
    _lib.require_cuda(x, "koa_resample_linear")
    if x.dtype not in _DTYPES:
        raise _lib.KoaError(f"koa_resample_linear takes float32 / uint8 / uint16 / int16 input, got {x.dtype}")
    if any(s < 1 for s in size_out):
        raise ValueError(f"output size {list(size_out)} is empty")
    x = x.contiguous()
    lead = x.shape[0] * x.shape[1]
    out = torch.empty(tuple(x.shape[:2]) + tuple(size_out), dtype=torch.float32, device=x.device)
    if lead == 0:
        return out
    din = (C.c_int * 3)(*_dims3(x.shape[2:]))
    dout = (C.c_int * 3)(*_dims3(size_out))
    with _lib.on_device(x.device):
        _lib.check(_lib.load().koa_resample_linear(x.data_ptr(), _DTYPES[x.dtype], out.data_ptr(), lead, din, dout,
                                                   _lib.ptr(scale), _lib.ptr(shift), _lib.current_stream()),
                   "koa_resample_linear")
    return out


class PTInterpolate:
    """``koafusion.preproc.PTInterpolate``: linear / bilinear / trilinear resampling of a (B, CH, D0, ...) batch."""

    def __init__(self, scale_factor):
        self.scale_factor = scale_factor

    def __call__(self, image, mask=None):
        if mask is not None:
            raise NotImplementedError("masks never reach the training / evaluation path (nearest-neighbour resampling "
                                      "of masks is not part of the CUDA path)")
        return _resample(image, output_size(image.shape[2:], self.scale_factor))


def downscale_x(x: torch.Tensor, factor) -> torch.Tensor:
    """``ProgressionPrediction._downscale_x`` (``train_prog_fus.py:111-116``): identity for an empty factor."""
    if factor:
        x = PTInterpolate(scale_factor=tuple(factor))(x)
    return x


def unit_range_affine(x: torch.Tensor, mean: float, std: float):
    """Per-volume ``(scale, shift)`` with ``scale * x + shift == PTNormalize(mean, std)(PTToUnitRange()(x))`` for every
    volume ``x[b, ch]`` (minimum / maximum over the whole volume). Returns ``(scale, shift, minmax)``, fp32 on the device."""
    _lib.require_cuda(x, "koa_unit_range_affine")
    if x.dtype not in _DTYPES:
        raise _lib.KoaError(f"koa_unit_range_affine takes float32 / uint8 / uint16 / int16 input, got {x.dtype}")
    x = x.contiguous()
    lead = x.shape[0] * x.shape[1]
    n_per = x[0, 0].numel()
    ws = torch.empty(2 * lead, dtype=torch.int32, device=x.device)
    scale = torch.empty(lead, dtype=torch.float32, device=x.device)
    shift = torch.empty_like(scale)
    minmax = torch.empty(lead, 2, dtype=torch.float32, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(_lib.load().koa_unit_range_affine(x.data_ptr(), _DTYPES[x.dtype], lead, n_per, float(mean), float(std),
                                                     ws.data_ptr(), scale.data_ptr(), shift.data_ptr(), minmax.data_ptr(),
                                                     _lib.current_stream()), "koa_unit_range_affine")
    return scale, shift, minmax


def unit_range_normalize_downscale(x: torch.Tensor, mean: float, std: float, factor=None) -> torch.Tensor:
    """``PTInterpolate(factor)(PTNormalize(mean, std)(PTToUnitRange()(x)))`` per volume of a (B, CH, D0, ...) batch in
    its storage type, in two passes over the integers (min / max, then the fused map); fp32 output."""
    scale, shift, _ = unit_range_affine(x, mean, std)
    size_out = output_size(x.shape[2:], tuple(factor)) if factor else list(x.shape[2:])
    return _resample(x, size_out, scale, shift)
