"""Optimiser step and learning-rate schedules of the reference on the CUDA path.

The reference builds ``dict_optimizers[name](model.parameters(), lr=..., weight_decay=...)`` and
``dict_schedulers[name](optimizer=..., **params)`` (``koafusion/run/train_prog_fus.py:88-96``; registries in
``koafusion/various/_optimizers.py:49-71``) and calls ``optimizer.step()`` after every backward (``:166``). ``Adam`` /
``AdamW`` here are ``torch.optim.Optimizer`` subclasses with the constructor, ``param_groups`` and ``state_dict`` layout
of their torch namesakes (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so LR schedulers and optimiser
checkpoints interoperate; ``step()`` is one ``koa_adam_step`` call per parameter group: every tensor of the group is
updated by a handful of multi-tensor launches (28 B of HBM traffic per element) instead of torch's per-tensor /
foreach kernels. No CPU or PyTorch fallback: CPU parameters raise.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import optim

from . import _lib


class Adam(optim.Optimizer):
    """``torch.optim.Adam`` (L2 weight decay added to the gradient, ``amsgrad=False``) through ``koa_adam_step``."""

    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad:
            raise ValueError("amsgrad is not implemented on the CUDA path (the reference trains with the default)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)  # host tensor, as torch keeps it (capturable=False)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            # validate the whole group first: a parameter that fails a check must not leave the step counters of the
            # parameters before it advanced without an update
            live = []
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                _lib.require_cuda(p, "koa_adam_step")
                _lib.require_cuda(g, "koa_adam_step")
                if g.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                if p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise _lib.KoaError("koa_adam_step updates fp32 master parameters with fp32 gradients")
                if not p.is_contiguous():
                    raise _lib.KoaError("koa_adam_step needs contiguous parameters")
                live.append((p, g if g.is_contiguous() else g.contiguous()))
            if live and any(p.device != live[0][0].device or g.device != live[0][0].device for p, g in live):
                raise _lib.KoaError("one parameter group must live on one device")
            by_step = {}
            for p, g in live:
                st = self._init_state(p)
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append((p, g, st))
            lr = group["lr"]
            if isinstance(lr, torch.Tensor):
                lr = float(lr)
            for step, entries in by_step.items():
                table = (_lib.AdamTensor * len(entries))()
                for i, (p, g, st) in enumerate(entries):
                    t = table[i]
                    t.param, t.grad, t.numel = p.data_ptr(), g.data_ptr(), p.numel()
                    t.exp_avg, t.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                hyper = _lib.AdamHyper(lr=lr, beta1=group["betas"][0], beta2=group["betas"][1], eps=group["eps"],
                                       weight_decay=group["weight_decay"], grad_scale=grad_scale, step=step,
                                       decoupled_weight_decay=int(self._decoupled))
                with _lib.on_device(entries[0][0].device):
                    _lib.check(lib.koa_adam_step(C.cast(table, C.c_void_p), len(entries), C.byref(hyper),
                                                 _lib.current_stream()), "koa_adam_step")
        return loss


class AdamW(Adam):
    """``torch.optim.AdamW``: the parameter is multiplied by ``1 - lr * weight_decay`` before the Adam update."""

    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)


# ---------------------------------------------------------------------------------------------------------------------
# learning-rate lambdas of the reference (koafusion/various/_optimizers.py:4-46): pure host arithmetic, once per epoch
# ---------------------------------------------------------------------------------------------------------------------
def _linear_warmup(epoch, epochs_warmup, warmup_factor):
    return warmup_factor + (1.0 - warmup_factor) * epoch / float(epochs_warmup)


def CustomWarmupStaticDecayLR(optimizer, epochs_warmup, epochs_static, epochs_decay, warmup_factor=0.1,
                              decay_factor=0.9, **kwargs):
    """Linear warm-up from ``warmup_factor`` to 1 over ``epochs_warmup`` epochs, 1 for ``epochs_static`` epochs, then
    ``decay_factor ** (epochs past the plateau)`` (``epochs_decay`` is accepted and unused, as in the reference)."""
    plateau_end = epochs_warmup + epochs_static

    def factor(epoch):
        if epoch <= epochs_warmup:
            return _linear_warmup(epoch, epochs_warmup, warmup_factor)
        if epoch <= plateau_end:
            return 1.0
        return decay_factor ** (epoch - plateau_end)

    return optim.lr_scheduler.LambdaLR(optimizer=optimizer, lr_lambda=factor)


def CustomWarmupMultiStepLR(optimizer, epochs_warmup, mstep_milestones, warmup_factor=0.1, mstep_factor=0.1, **kwargs):
    """Linear warm-up, then ``mstep_factor`` to the power of the number of milestones (counted from the end of the
    warm-up) that the epoch has reached."""
    milestones = [epochs_warmup + m for m in mstep_milestones]

    def factor(epoch):
        if epoch <= epochs_warmup:
            return _linear_warmup(epoch, epochs_warmup, warmup_factor)
        return mstep_factor ** sum(1 for m in milestones if epoch >= m)

    return optim.lr_scheduler.LambdaLR(optimizer=optimizer, lr_lambda=factor)


# Same names as the reference registries; the optimisers without a CUDA-path implementation stay torch's.
dict_optimizers = {
    "SGD": optim.SGD,
    "Adam": Adam,
    "AdamW": AdamW,
    "RMSprop": optim.RMSprop,
}

dict_schedulers = {
    "LambdaLR": optim.lr_scheduler.LambdaLR,
    "MultiplicativeLR": optim.lr_scheduler.MultiplicativeLR,
    "StepLR": optim.lr_scheduler.StepLR,
    "MultiStepLR": optim.lr_scheduler.MultiStepLR,
    "ExponentialLR": optim.lr_scheduler.ExponentialLR,
    "CosineAnnealingLR": optim.lr_scheduler.CosineAnnealingLR,
    "ReduceLROnPlateau": optim.lr_scheduler.ReduceLROnPlateau,
    "CyclicLR": optim.lr_scheduler.CyclicLR,
    "OneCycleLR": optim.lr_scheduler.OneCycleLR,
    "CosineAnnealingWarmRestarts": optim.lr_scheduler.CosineAnnealingWarmRestarts,
    "CustomWarmupStaticDecayLR": CustomWarmupStaticDecayLR,
    "CustomWarmupMultiStepLR": CustomWarmupMultiStepLR,
}
