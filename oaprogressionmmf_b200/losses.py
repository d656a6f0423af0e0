"""``FocalLoss`` of the reference (``koafusion/various/_losses.py:53-108``; gamma 2, mean reduction, no class
weights) on the CUDA path: loss and d(loss)/d(logits) come from one ``koa_focal_loss`` launch."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


class _FocalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, gamma):
        lib = _lib.load()
        dev = _lib.require_same_device("FocalLoss", logits, target)
        logits = logits.contiguous().float()
        target = target.reshape(-1).contiguous().long()
        b, c = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        with _lib.on_device(dev):
            _lib.check(lib.koa_focal_loss(logits.data_ptr(), target.data_ptr(), loss.data_ptr(), dlogits.data_ptr(), b, c,
                                          float(gamma), _lib.current_stream()), "koa_focal_loss")
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None


class FocalLoss(nn.Module):
    """Same constructor arguments as the reference (``_losses.py:54-87``). ``class_weight`` is the one it acts on that
    this path does not implement: asking for it raises instead of silently training an unweighted loss. A target outside
    ``[0, classes)`` gives a NaN loss and sets the library's diagnostic word (``_lib.debug_flag()``), where the
    reference's ``F.cross_entropy`` trips a device assert."""

    def __init__(self, num_classes=2, batch_avg=True, batch_weight=None, class_avg=True, class_weight=None, gamma=2,
                 reduction="mean", **kwargs):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise ValueError("Unknown `reduction` value")
        if class_weight is not None:
            raise ValueError("class_weight is not implemented on the CUDA path (koafusion trains without it: "
                             "run/conf/prog_fus.yaml); refusing to ignore it")
        self.num_classes = num_classes
        self.gamma = gamma
        self.reduction = reduction

    def forward(self, input, target, **kwargs):
        loss = _FocalLossFn.apply(input, target, self.gamma)  # mean over the batch
        if self.reduction == "sum":
            loss = loss * input.shape[0]
        return loss
