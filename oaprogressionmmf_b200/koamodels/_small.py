"""Small fp32 dense layers that sit next to the engines: the clinical-variable embedding ``FeatC1``
(``koafusion/models/_xrNmrMcP.py:11-29``), the ``XR1Cnn`` head (``_xr1_cnn.py:31-39``). They run on the
CUDA-core kernels ``koa_linear_small_{fwd,bwd}`` (N or K too small for tensor-core tiles)."""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib


class _SmallLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        lib = _lib.load()
        dev = _lib.require_same_device("small_linear", x, weight, bias)
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous().float()
        m, k = x2.shape
        n = weight.shape[0]
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        pre = torch.empty_like(y) if act != _lib.ACT_NONE else None
        with _lib.on_device(dev):
            _lib.check(lib.koa_linear_small_fwd(x2.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
                                                y.data_ptr(), None if pre is None else pre.data_ptr(), m, n, k, act,
                                                _lib.current_stream()), "koa_linear_small_fwd")
        ctx.save_for_backward(x2, weight, pre)
        ctx.act, ctx.shape, ctx.has_bias = act, shape, bias is not None
        return y.reshape(*shape[:-1], n)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x2, weight, pre = ctx.saved_tensors
        m, k = x2.shape
        n = weight.shape[0]
        dev = _lib.require_same_device("small_linear backward", dy, x2, weight)
        dy2 = dy.reshape(m, n).contiguous().float()
        scratch = torch.empty_like(dy2)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(weight) if ctx.needs_input_grad[1] else None
        db = torch.zeros(n, dtype=torch.float32, device=dy.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        with _lib.on_device(dev):
            _lib.check(lib.koa_linear_small_bwd(dy2.data_ptr(), None if pre is None else pre.data_ptr(), x2.data_ptr(),
                                                weight.data_ptr(), scratch.data_ptr(), None if dx is None else dx.data_ptr(),
                                                None if dw is None else dw.data_ptr(), None if db is None else db.data_ptr(),
                                                m, n, k, ctx.act, _lib.current_stream()), "koa_linear_small_bwd")
        return (None if dx is None else dx.reshape(ctx.shape), dw, db, None)


def small_linear(x, weight, bias, act=_lib.ACT_NONE):
    if not (weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()):
        raise _lib.KoaError("small_linear needs contiguous fp32 CUDA parameters")
    return _SmallLinear.apply(x, weight, bias, act)


class FeatC1(nn.Module):
    """Clinical token: Linear(dim_in -> dim_out) + GELU + Dropout (``_xrNmrMcP.py:11-29``)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self._fe = nn.Sequential(nn.Linear(config["dim_in"], config["dim_out"]), nn.GELU(), nn.Dropout(config["dropout"]))

    def forward(self, input_):
        out = small_linear(input_, self._fe[0].weight, self._fe[0].bias, _lib.ACT_GELU)
        return self._fe[2](out)
