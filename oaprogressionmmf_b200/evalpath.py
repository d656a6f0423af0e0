"""Evaluation / fold-ensemble arithmetic of the reference on the CUDA path.

``koafusion/run/eval_prog_fus.py:250-343``: every fold model is run over the test loader under ``no_grad``; per batch the
logits go to the host where ``argmax`` and ``softmax`` produce ``predict`` / ``predict_proba``; the per-fold tables are
then merged on ``exam_knee_id`` and the ensemble prediction is ``softmax(mean over folds of predict_proba)`` (the softmax
over already-normalised probabilities is the reference's, kept for parity) and its ``argmax``.

Here both steps stay on the device (``koa_predict``, ``koa_ensemble_proba``) and only the final lists cross to the host,
once per epoch instead of once per batch. The accumulator keys (``exam_knee_id``, ``target``, ``predict``,
``predict_proba``) are the reference's. No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, Sequence

import torch

from . import _lib


def predict(logits: torch.Tensor):
    """``(softmax(logits, dim=1), argmax(logits, dim=1))`` for ``(B, classes)`` fp32 logits, one launch."""
    _lib.require_cuda(logits, "koa_predict")
    if logits.ndim != 2:
        raise ValueError(f"expected (B, classes) logits, got {tuple(logits.shape)}")
    logits = logits.detach().contiguous().float()
    b, c = logits.shape
    proba = torch.empty_like(logits)
    pred = torch.empty(b, dtype=torch.int64, device=logits.device)
    if b:
        with _lib.on_device(logits.device):
            _lib.check(_lib.load().koa_predict(logits.data_ptr(), proba.data_ptr(), pred.data_ptr(), b, c,
                                               _lib.current_stream()), "koa_predict")
    return proba, pred


def ensemble_proba(proba_foldw: torch.Tensor):
    """``proba_foldw``: ``(folds, B, classes)`` per-fold probabilities -> ``(softmax(mean over folds), argmax)``."""
    _lib.require_cuda(proba_foldw, "koa_ensemble_proba")
    if proba_foldw.ndim != 3:
        raise ValueError(f"expected (folds, B, classes) probabilities, got {tuple(proba_foldw.shape)}")
    p = proba_foldw.detach().contiguous().float()
    f, b, c = p.shape
    out = torch.empty(b, c, dtype=torch.float32, device=p.device)
    pred = torch.empty(b, dtype=torch.int64, device=p.device)
    if b:
        with _lib.on_device(p.device):
            _lib.check(_lib.load().koa_ensemble_proba(p.data_ptr(), out.data_ptr(), pred.data_ptr(), f, b, c,
                                                      _lib.current_stream()), "koa_ensemble_proba")
    return out, pred


@torch.no_grad()
def predict_batched(model, inputs_host: Sequence[torch.Tensor], device, micro_batch: int = 32):
    """Class predictions and probabilities of an eval-mode model for a batch that lives in (pinned) host memory, in
    micro-batches of ``micro_batch`` knees (eval mode has no coupling between knees: BatchNorm uses its running
    statistics): the host -> device copy of micro-batch i + 1 runs on a copy stream while micro-batch i computes, and
    the predictions of the whole batch cross to the host once (the reference reads them per batch,
    ``eval_prog_fus.py:286-304``). Returns ``(pred int64 (B,), proba fp32 (B, classes))`` on the host."""
    if model.training:
        raise ValueError("predict_batched runs an eval-mode model (call model.eval() first)")
    device = torch.device(device)
    n = inputs_host[0].shape[0]
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)
    # two sets of device input buffers, allocated once on the compute stream (no tensor ever changes streams: the model
    # runs its modality branches on several streams, and blocks that migrate between per-stream allocator pools stall it):
    # the copy stream fills set k while the compute streams read set 1 - k; events order the two
    mb = min(micro_batch, n)
    bufs = [[torch.empty((mb,) + tuple(t.shape[1:]), dtype=t.dtype, device=device) for t in inputs_host] for _ in range(2)]
    filled = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    starts = list(range(0, n, mb))

    def fetch(j):
        k, i = j & 1, starts[j]
        m = min(mb, n - i)
        with torch.cuda.stream(copy):
            if j >= 2:
                copy.wait_event(consumed[k])
            else:
                copy.wait_stream(main)  # the buffers exist from here
            for dst, src in zip(bufs[k], inputs_host):
                dst[:m].copy_(src[i:i + m], non_blocking=True)
            filled[k].record(copy)
        return m

    sizes = {0: fetch(0)}
    probas, preds = [], []
    for j in range(len(starts)):
        if j + 1 < len(starts):
            sizes[j + 1] = fetch(j + 1)
        k = j & 1
        main.wait_event(filled[k])
        out = model(*[t[:sizes[j]] for t in bufs[k]])
        logits = out["main"] if isinstance(out, dict) else out
        proba, pred = predict(logits.reshape(logits.shape[0], -1))
        consumed[k].record(main)
        probas.append(proba)
        preds.append(pred)
    return torch.cat(preds).cpu(), torch.cat(probas).cpu()


@torch.no_grad()
def eval_epoch(model, loader: Iterable[dict], modals: Sequence[str], downscale=None, device=None) -> Dict[str, list]:
    """One pass of a fold model over a loader with the reference's batch contract (``image__{modal}``, ``target``,
    ``("-", "exam_knee_id")``; ``eval_prog_fus.py:250-312``). Returns the reference's accumulator: lists under
    ``exam_knee_id``, ``target``, ``predict``, ``predict_proba``. Predictions are gathered on the device and read back
    once at the end."""
    from .preproc import downscale_x

    was_training = model.training
    model.eval()
    ids, targets, probas, preds = [], [], [], []
    for batch in loader:
        xs = [batch[f"image__{m}"] for m in modals]
        if device is not None:
            xs = [x.to(device, non_blocking=True) for x in xs]
        if downscale:
            xs = [downscale_x(x, f) for x, f in zip(xs, downscale)]
        out = model(*xs)
        logits = out["main"] if isinstance(out, dict) else out
        proba, pred = predict(logits.reshape(logits.shape[0], -1))
        ids.extend(batch[("-", "exam_knee_id")])
        targets.append(batch["target"])
        probas.append(proba)
        preds.append(pred)
    model.train(was_training)
    acc = defaultdict(list)
    acc["exam_knee_id"] = list(ids)
    if probas:
        acc["target"] = torch.cat([t.reshape(t.shape[0], -1) for t in targets]).cpu().tolist()
        acc["predict"] = torch.cat(preds).cpu().tolist()
        acc["predict_proba"] = torch.cat(probas).cpu().tolist()
    return acc


def align_folds(raw_foldw: Dict[object, Dict[str, list]]):
    """Host half of ``ensemble_eval_foldw`` (``eval_prog_fus.py:314-329``): inner 1:1 join of the per-fold tables on
    ``exam_knee_id`` in the row order of the first fold. Returns ``(ids, target, {fold: [row index per id]})``."""
    folds = list(raw_foldw)
    if not folds:
        raise ValueError("no folds to merge")
    index = {}
    for f in folds:
        ids = raw_foldw[f]["exam_knee_id"]
        idx = {k: i for i, k in enumerate(ids)}
        if len(idx) != len(ids):
            raise ValueError(f"fold {f}: exam_knee_id is not unique (the reference validates the merge as 1:1)")
        index[f] = idx
    first = raw_foldw[folds[0]]
    keep = [k for k in first["exam_knee_id"] if all(k in index[f] for f in folds[1:])]
    rows = {f: [index[f][k] for k in keep] for f in folds}
    target = [first["target"][i] for i in rows[folds[0]]]
    return keep, target, rows


def ensemble_eval_foldw(raw_foldw: Dict[object, Dict[str, list]], device="cuda") -> Dict[str, list]:
    """``ProgressionPrediction.ensemble_eval_foldw``: merged table with ``predict__{fold}`` / ``predict_proba__{fold}``
    columns and the ensemble ``predict_proba`` / ``predict`` computed by ``koa_ensemble_proba``."""
    ids, target, rows = align_folds(raw_foldw)
    out = {"exam_knee_id": ids, "target": target}
    stack = []
    for f, idx in rows.items():
        d = raw_foldw[f]
        out[f"predict__{f}"] = [d["predict"][i] for i in idx]
        out[f"predict_proba__{f}"] = [d["predict_proba"][i] for i in idx]
        stack.append(out[f"predict_proba__{f}"])
    if ids:
        proba, pred = ensemble_proba(torch.tensor(stack, dtype=torch.float32, device=device))
        out["predict_proba"] = proba.cpu().tolist()
        out["predict"] = pred.cpu().tolist()
    else:
        out["predict_proba"], out["predict"] = [], []
    return out
