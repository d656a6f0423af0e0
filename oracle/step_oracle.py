"""CPU oracle of the rows next to the hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import this module; the product
(``oaprogressionmmf_b200``) never does (``tests/test_abi.py::test_product_never_imports_the_oracle``).

Plain numpy restatements, each citing the reference lines it follows, written as explicit index / scalar arithmetic so
that they are independent of the library kernels the reference happens to call:

* ``adam_step``                — ``torch.optim.Adam`` / ``AdamW`` as the reference constructs and steps them
                                 (``koafusion/various/_optimizers.py:49-54``; ``koafusion/run/train_prog_fus.py:88-91,166``).
                                 The algorithm lives in torch (pinned 2.5.1, ``env_base.yml:12``; installed 2.11):
                                 ``torch/optim/adam.py::_single_tensor_adam``. Pinned by ``tests/test_step_oracle.py``
                                 against ``torch.optim.Adam`` / ``AdamW`` themselves (torch travels to the GPU box).
* ``interpolate_linear``       — ``PTInterpolate`` (``koafusion/preproc/_pt.py:175-200``): ``F.interpolate(mode=(tri|bi)linear,
                                 align_corners=False, recompute_scale_factor=True)``. Pinned against outputs of the
                                 unmodified reference class (``tests/golden_step/step_rows.json``, ``oracle/make_golden_step.py``).
* ``unit_range_normalize``     — ``PTToUnitRange`` + ``PTNormalize`` (``koafusion/preproc/_pt.py:75-124``), same fixture.
* ``augment_chain``            — crop + ``PTToUnitRange`` + ``PTRotate3DInSlice`` / ``PTRotate2D`` + ``PTGammaCorrection`` +
                                 ``PTNormalize`` (``koafusion/preproc/_np_nd.py:62-140``, ``_pt.py:75-124,203-358``) in the
                                 order of ``koafusion/datasets/_data_provider.py:297-334``; same fixture (reference
                                 classes with their random state set explicitly).
* ``predict`` / ``ensemble``   — ``koafusion/run/eval_prog_fus.py:300-304,330-336``; the fixture holds the output of the
                                 reference's own ``ensemble_eval_foldw`` source, executed unmodified.
* ``lr_lambda_*``              — ``koafusion/various/_optimizers.py:4-46``, fixture from the reference functions.
"""
from __future__ import annotations

import math

import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# Adam
# ---------------------------------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
    """One update of fp32 arrays ``p, m, v`` (returned as new arrays) with gradient ``g``; ``step`` counts from 1.
    Element arithmetic in fp32, scalars formed in double and rounded to fp32 where torch hands them to a tensor op."""
    f = np.float32
    p, g, m, v = (np.asarray(a, dtype=f) for a in (p, g, m, v))
    b1, b2 = betas
    if decoupled:
        p = p * f(1.0 - lr * weight_decay)
    elif weight_decay != 0:
        g = g + f(weight_decay) * p
    m = m + (g - m) * f(1.0 - b1)
    v = v * f(b2) + f(1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


# ---------------------------------------------------------------------------------------------------------------------
# resampling
# ---------------------------------------------------------------------------------------------------------------------
def _taps(n_in: int, n_out: int):
    """Source indices / weights of linear interpolation with align_corners=False and the scale recomputed from the sizes
    (ATen ``area_pixel_compute_source_index``): src = (n_in / n_out) * (dst + 0.5) - 0.5, clamped at 0."""
    f = np.float32
    rscale = f(n_in) / f(n_out)
    dst = np.arange(n_out, dtype=f)
    src = np.maximum(rscale * (dst + f(0.5)) - f(0.5), f(0))
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (src - i0.astype(f)).astype(f)
    return i0, i1, (f(1) - l1).astype(f), l1


def output_size(size_in, scale_factor):
    if isinstance(scale_factor, (int, float)):
        scale_factor = [scale_factor] * len(size_in)
    return [int(math.floor(float(s) * float(k))) for s, k in zip(size_in, scale_factor)]


def interpolate_linear(x, scale_factor):
    """``PTInterpolate(scale_factor)(x)`` for a (B, CH, D0[, D1[, D2]]) array: separable two-tap interpolation along
    every spatial axis, fp32."""
    x = np.asarray(x, dtype=np.float32)
    spatial = x.shape[2:]
    size_out = output_size(spatial, scale_factor)
    for ax, (n_in, n_out) in enumerate(zip(spatial, size_out)):
        i0, i1, l0, l1 = _taps(n_in, n_out)
        shape = [1] * x.ndim
        shape[2 + ax] = n_out
        x = np.take(x, i0, axis=2 + ax) * l0.reshape(shape) + np.take(x, i1, axis=2 + ax) * l1.reshape(shape)
        x = x.astype(np.float32)
    return x


def unit_range_normalize(x, mean, std):
    """``PTNormalize(mean, std)(PTToUnitRange()(x[b, ch]))`` for every volume of a (B, CH, ...) array."""
    f = np.float32
    x = np.asarray(x).astype(f)
    out = np.empty_like(x)
    for b in range(x.shape[0]):
        for ch in range(x.shape[1]):
            vol = x[b, ch]
            lo, hi = vol.min(), vol.max()
            out[b, ch] = ((vol - lo) / (hi - lo) - f(mean)) / f(std)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# per-sample transform chain of the training loader
# ---------------------------------------------------------------------------------------------------------------------
def rotate_in_slice(vol, theta):
    """``PTRotate3DInSlice`` / ``PTRotate2D`` with the rotation applied (``koafusion/preproc/_pt.py:257-358``): every slice
    of a (R, C, S) volume resampled on the grid of ``F.affine_grid([[cos, -sin, 0], [sin, cos, 0]], align_corners=False)``
    with ``F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=False)``, written out as index arithmetic."""
    f = np.float32
    vol = np.asarray(vol, dtype=f)
    R, C, _ = vol.shape
    cs, sn = f(math.cos(theta)), f(math.sin(theta))
    x = ((f(2) * np.arange(C, dtype=f) + f(1)) / f(C) - f(1))[None, :]
    y = ((f(2) * np.arange(R, dtype=f) + f(1)) / f(R) - f(1))[:, None]
    gx, gy = cs * x - sn * y, sn * x + cs * y
    ix, iy = ((gx + f(1)) * f(C) - f(1)) * f(0.5), ((gy + f(1)) * f(R) - f(1)) * f(0.5)
    x0, y0 = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)
    wx1, wy1 = (ix - x0).astype(f), (iy - y0).astype(f)
    out = np.zeros_like(vol)
    for dy, wy in ((0, f(1) - wy1), (1, wy1)):
        for dx, wx in ((0, f(1) - wx1), (1, wx1)):
            yy, xx = y0 + dy, x0 + dx
            ok = (yy >= 0) & (yy < R) & (xx >= 0) & (xx < C)
            tap = vol[np.clip(yy, 0, R - 1), np.clip(xx, 0, C - 1), :] * ok[..., None]
            out += tap * (wx * wy)[..., None].astype(f)
    return out.astype(f)


def augment_chain(raw, offsets, crop_size, theta, gamma, mean, std, factor, flip=0):
    """One sample of the training loader (``koafusion/datasets/_data_provider.py:297-334``) and the on-GPU downscale:
    crop -> ``PTToUnitRange`` -> rotation (``theta`` or None) -> ``PTGammaCorrection`` (``gamma`` or None) ->
    ``PTNormalize`` -> ``PTInterpolate``. ``raw``: stored (R, C[, S]) array; ``flip``: 0, 1 (columns) or 2 (slices), the
    mirroring of RIGHT knees in front of the transforms; returns (1, R', C'[, S']) fp32."""
    f = np.float32
    raw = np.asarray(raw)
    two_d = raw.ndim == 2
    if flip:                # RIGHT knee: mirrored before the transforms (koafusion/datasets/oai/_dataset.py:303-316)
        raw = np.flip(raw, axis=flip)
    sel = tuple(slice(o, o + c) for o, c in zip(offsets, crop_size))
    vol = raw[sel].astype(f)
    vol = (vol - vol.min()) / (vol.max() - vol.min())
    if theta is not None:
        vol = rotate_in_slice(vol[..., None], theta)[..., 0] if two_d else rotate_in_slice(vol, theta)
    if gamma is not None:
        vol = np.power(vol, f(1.0 / gamma)).astype(f)
    vol = ((vol - f(mean)) / f(std)).astype(f)[None, None]
    if factor:
        vol = interpolate_linear(vol, factor)
    return vol[0]


# ---------------------------------------------------------------------------------------------------------------------
# predictions
# ---------------------------------------------------------------------------------------------------------------------
def _softmax(t):
    t = np.asarray(t, dtype=np.float64)
    e = np.exp(t - t.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def predict(logits):
    """``(softmax(logits, dim=1), argmax(logits, dim=1))`` (``eval_prog_fus.py:300-304``)."""
    logits = np.asarray(logits, dtype=np.float32)
    return _softmax(logits), logits.argmax(axis=1)


def ensemble(proba_foldw):
    """``(folds, B, classes)`` -> ``softmax(mean over folds)`` and its argmax (``eval_prog_fus.py:330-336``)."""
    t = _softmax(np.mean(np.asarray(proba_foldw, dtype=np.float64), axis=0))
    return t, t.argmax(axis=-1)


def ensemble_eval_foldw(raw_foldw):
    """The whole of ``ensemble_eval_foldw`` (``eval_prog_fus.py:314-339``) on plain dicts: inner 1:1 join on
    ``exam_knee_id`` in the order of the first fold, then ``ensemble``."""
    folds = list(raw_foldw)
    index = {f: {k: i for i, k in enumerate(raw_foldw[f]["exam_knee_id"])} for f in folds}
    keep = [k for k in raw_foldw[folds[0]]["exam_knee_id"] if all(k in index[f] for f in folds)]
    out = {"exam_knee_id": keep, "target": [raw_foldw[folds[0]]["target"][index[folds[0]][k]] for k in keep]}
    for f in folds:
        out[f"predict__{f}"] = [raw_foldw[f]["predict"][index[f][k]] for k in keep]
        out[f"predict_proba__{f}"] = [raw_foldw[f]["predict_proba"][index[f][k]] for k in keep]
    proba, pred = ensemble([out[f"predict_proba__{f}"] for f in folds])
    out["predict_proba"] = proba.tolist()
    out["predict"] = pred.tolist()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# learning-rate lambdas
# ---------------------------------------------------------------------------------------------------------------------
def lr_lambda_warmup_static_decay(epoch, epochs_warmup, epochs_static, warmup_factor=0.1, decay_factor=0.9):
    if epoch <= epochs_warmup:
        return warmup_factor + (1.0 - warmup_factor) * epoch / float(epochs_warmup)
    if epoch <= epochs_warmup + epochs_static:
        return 1.0
    return decay_factor ** (epoch - epochs_warmup - epochs_static)


def lr_lambda_warmup_multistep(epoch, epochs_warmup, mstep_milestones, warmup_factor=0.1, mstep_factor=0.1):
    if epoch <= epochs_warmup:
        return warmup_factor + (1.0 - warmup_factor) * epoch / float(epochs_warmup)
    return mstep_factor ** sum(epoch >= epochs_warmup + e for e in mstep_milestones)
