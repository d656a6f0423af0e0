"""The two epoch loops of the reference trainer around the CUDA path.

``koafusion/run/train_prog_fus.py:119-170`` (``train_epoch``) and ``:172-240`` (``val_epoch``) with the same order of
operations per batch — ``zero_grad``, modalities in ``modals`` order to the device, "last-chance" downscale, model,
loss on ``pred.squeeze(1)`` / ``target.long().squeeze(1)``, ``backward``, optimiser step — and the same return value
(``{"batch-w": {"loss_prog": [...]}, "epoch-w": {...}}``; validation losses rounded to three decimals as the reference
logs them, ``:212``). What differs is where the host waits: the reference reads ``loss.item()`` twice per batch (and the
probabilities once more in validation), which serialises host and device every step; here the per-batch losses and the
validation predictions stay on the device and cross to the host once, at the end of the epoch. TensorBoard / tqdm /
checkpointing (``fit``, ``:242-340``) are the reference's control plane and stay there: ``fit`` keeps calling these two
functions.

``transforms``: optional callable ``batch -> positional inputs`` replacing the extract + downscale steps (for batches of
stored volumes: ``functools.partial(synthetic.device_transforms, modals=..., sizes=..., downscale=..., device=...)``).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, Iterable, Optional, Sequence

import numpy as np
import torch

from .evalpath import predict
from .preproc import downscale_x


def _inputs(batch, modals, downscale, device, transforms):
    if transforms is not None:
        return tuple(transforms(batch))
    xs = tuple(batch[f"image__{m}"] for m in modals)
    if device is not None:
        xs = tuple(x.to(device, non_blocking=True) for x in xs)
    if downscale:
        xs = tuple(downscale_x(x, tuple(f)) for x, f in zip(xs, downscale))
    return xs


def _loss(model, loss_fn, xs, target):
    out = model(*xs)
    pred = out["main"] if isinstance(out, dict) else out
    return pred, loss_fn(input=pred.squeeze(1), target=target.long().squeeze(1))


def train_epoch(model, loader: Iterable[dict], modals: Sequence[str], loss_fn, optimizer, downscale=None, device=None,
                transforms: Optional[Callable] = None) -> dict:
    """One training epoch (``train_prog_fus.py:119-170``). The model must be in train mode (``fit`` sets it, ``:254``)."""
    metrics = {"batch-w": defaultdict(list), "epoch-w": dict()}
    losses = []
    for batch in loader:
        optimizer.zero_grad()
        xs = _inputs(batch, modals, downscale, device, transforms)
        target = batch["target"]
        if device is not None:
            target = target.to(device, non_blocking=True)
        _, loss = _loss(model, loss_fn, xs, target)
        losses.append(loss.detach())
        loss.backward()
        optimizer.step()
    if losses:
        metrics["batch-w"]["loss_prog"] = torch.stack(losses).float().cpu().tolist()      # the epoch's only read-back
    return metrics


@torch.no_grad()
def val_epoch(model, loader: Iterable[dict], modals: Sequence[str], loss_fn, downscale=None, device=None,
              transforms: Optional[Callable] = None, metrics_fn: Optional[Callable] = None) -> dict:
    """One validation epoch (``train_prog_fus.py:172-240``); the model must be in eval mode (``fit``, ``:264``).
    ``metrics_fn(prog_target=..., prog_pred_proba=...)`` is the reference's ``calc_metrics_v2`` (host-side scikit-learn
    statistics) when given; without it ``"epoch-w"`` carries the two arrays it would receive."""
    metrics = {"batch-w": defaultdict(list), "epoch-w": dict()}
    losses, targets, probas = [], [], []
    for batch in loader:
        xs = _inputs(batch, modals, downscale, device, transforms)
        target = batch["target"]
        if device is not None:
            target = target.to(device, non_blocking=True)
        pred, loss = _loss(model, loss_fn, xs, target)
        losses.append(loss.detach())
        targets.append(target)
        probas.append(predict(pred.reshape(pred.shape[0], -1))[0])
    if not losses:
        return metrics
    metrics["batch-w"]["loss_prog"] = [float(np.round(v, 3)) for v in torch.stack(losses).float().cpu().tolist()]
    t_target = torch.cat(targets, dim=0).cpu().numpy()
    t_proba = torch.cat(probas, dim=0).cpu().numpy()
    if metrics_fn is not None:
        metrics["epoch-w"] = metrics_fn(prog_target=t_target, prog_pred_proba=t_proba)
    else:
        metrics["epoch-w"] = {"target": t_target, "predict_proba": t_proba}
    return metrics
