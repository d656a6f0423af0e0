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

* ``augment_normalize_downscale`` goes one step further up the loader: the whole per-sample transform chain of
  ``koafusion/datasets/_data_provider.py:297-334`` (crop, unit range, in-slice rotation, gamma, z-score) and the downscale
  in one ``koa_augment_resample`` call on the stored integers, so the 24 CPU workers of the reference only read files.

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


# ---------------------------------------------------------------------------------------------------------------------
# the per-sample transform chain of the loaders, fused with the downscale
# ---------------------------------------------------------------------------------------------------------------------
# axis mirrored for RIGHT knees, per sequence (koafusion/datasets/oai/_dataset.py:303-316; the value of koa_augment_t::flip)
FLIP_AXIS = {"sag_3d_dess": 2, "cor_iw_tse": 1, "sag_t2_map": 2, "xr_pa": 1}


def crop_offsets(size_in: Sequence[int], size_out: Sequence[int], ratios=None) -> list:
    """First voxel of the crop: ``RandomCrop`` (``floor(ratio * (in - out))`` per axis, ``ratios`` in [0, 1)) or, without
    ratios, ``CenterCrop`` (``(in - out) // 2``) — ``koafusion/preproc/_np_nd.py:62-140``."""
    for d_in, d_out in zip(size_in, size_out):
        if d_in < d_out:
            raise ValueError(f"Invalid crop size {list(size_out)!r} for input {list(size_in)!r}")
    if ratios is None:
        return [(i - o) // 2 for i, o in zip(size_in, size_out)]
    return [int(math.floor(r * (i - o))) for r, i, o in zip(ratios, size_in, size_out)]


def draw_train_state(rng, size_in, size_out, degree_range=(-15.0, 15.0), rotate_prob=0.5, gamma_range=(0.5, 2.0),
                     gamma_prob=0.5, with_gamma=True) -> dict:
    """One sample's random state, drawn as the reference transforms draw theirs (``random.random()`` per crop axis,
    ``p`` then ``theta`` for the rotation, ``p`` then ``gamma`` for the gamma correction; ``_np_nd.py:103-105``,
    ``_pt.py:229-232,305-307``). ``rng`` is a ``random.Random``. Sequences without gamma (T2 maps) pass
    ``with_gamma=False``."""
    ratios = [rng.random() for _ in size_out]
    p_rot, theta = rng.random(), rng.uniform(math.radians(degree_range[0]), math.radians(degree_range[1]))
    state = {"offsets": crop_offsets(size_in, size_out, ratios), "theta": theta if p_rot < rotate_prob else None,
             "gamma": None}
    if with_gamma:
        p_gam, gamma = rng.random(), rng.uniform(*gamma_range)
        state["gamma"] = gamma if p_gam < gamma_prob else None
    return state


def augment_normalize_downscale(x: torch.Tensor, crop_size: Sequence[int], states: Sequence[dict], mean: float, std: float,
                                factor=None) -> torch.Tensor:
    """``PTInterpolate(factor)(PTNormalize(mean, std)(PTGammaCorrection(PTRotate*(PTToUnitRange(crop(x))))))`` for every
    stored volume ``x[b, 0]`` of a (B, 1, R, C[, S]) batch in its storage type, with the per-sample state of
    ``draw_train_state`` (``offsets``, ``theta`` or None, ``gamma`` or None). The validation / test chain is the same call
    with ``{"offsets": crop_offsets(in, out), "theta": None, "gamma": None}``. An optional ``"flip"`` entry mirrors the
    stored volume first, as the dataset does for RIGHT knees (``koafusion/datasets/oai/_dataset.py:303-316``): ``FLIP_AXIS``
    of the sequence (columns for COR IW TSE and XR, slices for SAG 3D DESS and SAG T2 map) or None / 0; the offsets count
    in the mirrored volume. One ``koa_augment_resample`` call; fp32 output of shape (B, 1, *floor(crop * factor))."""
    _lib.require_cuda(x, "koa_augment_resample")
    if x.dtype not in _DTYPES:
        raise _lib.KoaError(f"koa_augment_resample takes float32 / uint8 / uint16 / int16 input, got {x.dtype}")
    if x.ndim not in (4, 5) or x.shape[1] != 1:
        raise ValueError(f"expected a (B, 1, R, C[, S]) batch, got {tuple(x.shape)}")
    spatial = list(x.shape[2:])
    crop_size = [int(c) for c in crop_size]
    if len(crop_size) != len(spatial) or len(states) != x.shape[0]:
        raise ValueError("crop size / state list do not match the batch")
    size_out = output_size(crop_size, tuple(factor)) if factor else list(crop_size)
    if any(s < 1 for s in size_out):
        raise ValueError(f"output size {size_out} is empty")
    pad = [1] * (3 - len(spatial))           # a 2-D image is a volume with one slice
    table = (_lib.Augment * len(states))()
    for t, st in zip(table, states):
        off = list(st["offsets"]) + [0] * len(pad)
        for o, c, s in zip(off, crop_size + pad, spatial + pad):
            if o < 0 or o + c > s:
                raise ValueError(f"crop {crop_size} at {st['offsets']} leaves the stored volume {spatial}")
        t.off0, t.off1, t.off2 = off
        theta = st.get("theta")
        t.rotate = int(theta is not None)
        t.cos_t, t.sin_t = (math.cos(theta), math.sin(theta)) if theta is not None else (1.0, 0.0)
        gamma = st.get("gamma")
        t.inv_gamma = 1.0 / gamma if gamma is not None else 0.0
        flip = st.get("flip") or 0
        if flip not in (0, 1, 2) or (flip == 2 and len(spatial) == 2):
            raise ValueError(f"flip must be None / 0, 1 (columns) or 2 (slices of a 3-D volume), got {flip!r}")
        t.flip = flip
    x = x.contiguous()
    b = x.shape[0]
    params = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(x.device)
    ws = torch.empty(2 * b, dtype=torch.int32, device=x.device)
    out = torch.empty((b, 1) + tuple(size_out), dtype=torch.float32, device=x.device)
    dims = [(C.c_int * 3)(*(list(d) + pad)) for d in (spatial, crop_size, size_out)]
    with _lib.on_device(x.device):
        _lib.check(_lib.load().koa_augment_resample(x.data_ptr(), _DTYPES[x.dtype], out.data_ptr(), params.data_ptr(), b,
                                                    dims[0], dims[1], dims[2], float(mean), float(std), ws.data_ptr(),
                                                    _lib.current_stream()), "koa_augment_resample")
    return out
