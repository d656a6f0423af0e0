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
        logits = logits.contiguous().float()
        target = target.reshape(-1).contiguous().long()
        b, c = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        _lib.check(lib.koa_focal_loss(logits.data_ptr(), target.data_ptr(), loss.data_ptr(), dlogits.data_ptr(), b, c,
                                      float(gamma), _lib.current_stream()), "koa_focal_loss")
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None


class FocalLoss(nn.Module):
    def __init__(self, num_classes=2, gamma=2, reduction="mean", **kwargs):
        super().__init__()
        if reduction != "mean":
            raise ValueError("the CUDA path implements reduction='mean' (the configuration koafusion trains with)")
        self.num_classes = num_classes
        self.gamma = gamma

    def forward(self, input, target, **kwargs):
        return _FocalLossFn.apply(input, target, self.gamma)
