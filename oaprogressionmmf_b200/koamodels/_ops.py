"""Per-operator autograd Functions over the C ABI (``koa_gemm_bf16``, ``koa_gemm_wgrad_bf16``, ``koa_layernorm_*``,
``koa_attention_*_fmt``): the stand-alone ``forward`` of the classes ``koafusion.models`` exports next to ``FeaT``
(``Transformer``, ``Attention``, ``FeedForward``; reference ``_core_trf.py:141-205``).

The model classes never go through these: ``FeaT`` runs the whole transformer as one engine call with every bias / GELU /
dropout / residual fused into GEMM epilogues (``koa_feat_forward``). This module gives a caller who instantiates one of
the building blocks on its own the same arithmetic (fp16 operands, fp32 accumulation and residual stream, bf16 gradients),
one launch per operator; activations, dropout and residual adds between the operators are plain element-wise torch ops.
CUDA only, like everything else: CPU tensors raise ``KoaError``.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def _rows(x):
    return x.reshape(-1, x.shape[-1])


class _Linear16(torch.autograd.Function):
    """y = x @ W^T + b on tcgen05: fp16 operands, fp32 accumulate / output; backward in bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib.load()
        dev = _lib.require_same_device("koa linear", x, weight, bias)
        x2 = _rows(x).contiguous()
        m, k = x2.shape
        n = weight.shape[0]
        if k % 8 or n % 32:
            raise _lib.KoaError(f"koa linear needs in_features % 8 == 0 and out_features % 32 == 0 (got {k}, {n})")
        xh = x2.to(torch.float16)
        wh = weight.detach().to(torch.float16).contiguous()
        y = torch.empty((m, n), dtype=torch.float32, device=dev)
        b32 = None if bias is None else bias.detach().float().contiguous()  # (kept alive until the launch is queued)
        ep = _lib.Epilogue(out=y.data_ptr(), ldo=n, out_fp32=1, a_f16=1, b_f16=1, bias=None if b32 is None else b32.data_ptr())
        with _lib.on_device(dev):
            _lib.check(lib.koa_gemm_bf16(xh.data_ptr(), wh.data_ptr(), m, n, k, C.byref(ep), _lib.current_stream()), "koa_gemm_bf16")
        ctx.save_for_backward(xh, weight)
        ctx.has_bias = bias is not None
        ctx.shape = x.shape
        return y.reshape(*x.shape[:-1], n)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        xh, weight = ctx.saved_tensors
        m, k = xh.shape
        n = weight.shape[0]
        dev = xh.device
        g = gy.reshape(m, n).to(torch.bfloat16).contiguous()
        dx = dw = db = None
        with _lib.on_device(dev):
            if ctx.needs_input_grad[0]:
                wt = weight.detach().t().contiguous().to(torch.bfloat16)     # [K][N]: the B operand of dX = dY . W
                dx = torch.empty((m, k), dtype=torch.float32, device=dev)
                ep = _lib.Epilogue(out=dx.data_ptr(), ldo=k, out_fp32=1)
                if n % 8 or k % 32:
                    raise _lib.KoaError("koa linear backward needs out_features % 8 == 0 and in_features % 32 == 0")
                _lib.check(lib.koa_gemm_bf16(g.data_ptr(), wt.data_ptr(), m, k, n, C.byref(ep), _lib.current_stream()), "koa_gemm_bf16")
                dx = dx.reshape(ctx.shape)
            if ctx.needs_input_grad[1]:
                dw = torch.zeros((n, k), dtype=torch.float32, device=dev)
                if k % 64 == 0 and n % 8 == 0:
                    # fp16 activation converted to bf16 inside the kernel (x_f16 = 2)
                    _lib.check(lib.koa_gemm_wgrad_bf16(g.data_ptr(), xh.data_ptr(), dw.data_ptr(), m, n, k, 2,
                                                       _lib.current_stream()), "koa_gemm_wgrad_bf16")
                else:
                    raise _lib.KoaError("koa linear weight gradient needs in_features % 64 == 0")
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = g.float().sum(0)
        return dx, dw, db


def linear(x, weight, bias=None):
    return _Linear16.apply(x, weight, bias)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib.load()
        dev = _lib.require_same_device("koa layer norm", x, weight, bias)
        x2 = _rows(x).contiguous().float()
        rows, d = x2.shape
        out = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        with _lib.on_device(dev):
            _lib.check(lib.koa_layernorm_fwd(x2.data_ptr(), weight.data_ptr(), bias.data_ptr(), None, out.data_ptr(),
                                             mean.data_ptr(), rstd.data_ptr(), rows, d, _lib.current_stream()), "koa_layernorm_fwd")
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.shape = x.shape
        return out.reshape(x.shape)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, weight, mean, rstd = ctx.saved_tensors
        rows, d = x2.shape
        g = gy.reshape(rows, d).contiguous().float()
        dx = torch.empty_like(x2)
        dg = torch.zeros(d, dtype=torch.float32, device=x2.device)
        db = torch.zeros_like(dg)
        with _lib.on_device(x2.device):
            _lib.check(lib.koa_layernorm_bwd(g.data_ptr(), x2.data_ptr(), weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, d, _lib.current_stream()),
                       "koa_layernorm_bwd")
        return dx.reshape(ctx.shape), dg, db


def layer_norm(x, ln: torch.nn.LayerNorm):
    if ln.eps != 1e-5 or not ln.elementwise_affine:
        raise _lib.KoaError("koa layer norm implements nn.LayerNorm(dim) with its default eps and affine parameters")
    return _LayerNorm.apply(x, ln.weight, ln.bias)


class _Attention(torch.autograd.Function):
    """softmax(Q K^T * scale) V on a packed fp16 qkv tensor (B, n, 3 * D) with the (qkv, head, d) feature split."""

    @staticmethod
    def forward(ctx, qkv, heads, scale):
        lib = _lib.load()
        _lib.require_cuda(qkv, "koa attention")
        b, n, d3 = qkv.shape
        d = d3 // 3
        hd = d // heads
        q16 = qkv.to(torch.float16).contiguous()
        out = torch.empty((b, n, d), dtype=torch.float16, device=qkv.device)
        probs = torch.empty((b, heads, n, n), dtype=torch.float32, device=qkv.device)
        with _lib.on_device(qkv.device):
            _lib.check(lib.koa_attention_fwd_fmt(q16.data_ptr(), out.data_ptr(), probs.data_ptr(), b, n, heads, hd, float(scale), 1,
                                                 _lib.current_stream()), "koa_attention_fwd_fmt")
        ctx.save_for_backward(q16, probs)
        ctx.args = (b, n, heads, hd, float(scale))
        ctx.mark_non_differentiable(probs)
        return out.float(), probs

    @staticmethod
    def backward(ctx, gout, _gprobs):
        lib = _lib.load()
        q16, probs = ctx.saved_tensors
        b, n, heads, hd, scale = ctx.args
        g = gout.to(torch.bfloat16).contiguous()
        dqkv = torch.empty(q16.shape, dtype=torch.bfloat16, device=q16.device)
        with _lib.on_device(q16.device):
            _lib.check(lib.koa_attention_bwd_fmt(q16.data_ptr(), probs.data_ptr(), g.data_ptr(), dqkv.data_ptr(), b, n, heads, hd,
                                                 scale, 1, _lib.current_stream()), "koa_attention_bwd_fmt")
        return dqkv.float(), None, None


def attention(qkv, heads, scale):
    return _Attention.apply(qkv, heads, scale)
