"""Parity of the CUDA path (through the C ABI of libkoa_b200.so) against the oracle, on a real B200.

Three layers of evidence, all ``-m gpu``:

1. **Operators** (``koa_gemm_bf16``, ``koa_conv_fprop_bf16``, ``koa_*_wgrad_bf16``, LayerNorm, attention, max-pool,
   focal loss, small linears, stem pack): called directly through ctypes on seeded inputs and compared with the
   same op in fp32 PyTorch on the same (bf16-representable) operands. Tolerances are one bf16 rounding of the output
   (2^-8 relative per element, 4e-3 relative L2) for bf16 outputs and 1e-4 for fp32 outputs.
2. **Engines** (``koa_fe_*`` = whole per-slice CNN, ``koa_feat_*`` = whole token transformer): compared with
   ``oracle.koa_oracle.fe_forward`` / ``feat_forward``. Eval mode is held to BASELINE.json's 1e-2 relative
   tolerance. Train-mode BatchNorm at random initialisation amplifies *any* perturbation from block to block
   (the fp32 oracle and the same oracle with activations rounded to bf16 differ by 1e-2 .. 3e-1 at the feature
   level, depending on the residual gain), so the train-mode bar is "at the bf16 precision floor": the error of
   the CUDA path against the fp32 oracle must not exceed the error of the bf16-emulating oracle against the fp32
   oracle by more than a stated factor. The backward pass is additionally checked with the forward state forced
   to be identical (teacher forcing), which removes the ReLU-mask avalanche from the gradient comparison.
3. **Models**: the six ``koafusion.models`` classes (+ the two 3-MRI extensions) replayed on the golden fixtures
   generated from the unmodified reference (``oracle/make_golden.py``): eval logits within 1e-2 relative and
   identical class predictions, state_dict round trip, gradients present exactly where the reference has them.

Nothing here reads /root/reference; the oracle is the checker, never the thing measured.
"""
import ctypes as C
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oaprogressionmmf_b200 import _lib
from oracle import koa_oracle as ko
from tests.util import rel, to_attr

pytestmark = pytest.mark.gpu

BF16_REL_L2 = 4e-3      # one bf16 rounding of the output: 2^-9 max, ~1.1e-3 rms per element
BF16_ELEM = 2.0 ** -7   # per-element bound used with an absolute floor
LOGIT_TOL = 1e-2        # BASELINE.json: logits within 1e-2 relative error under bf16
TINY_LOGIT_TOL = 2e-2   # tiny golden fixtures (four small logits): see test_model_eval_logits_match_reference


def _stream():
    return _lib.current_stream()


def _bf(t):
    return t.to(torch.bfloat16).contiguous()


def _randn(*shape, seed=0, scale=1.0, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def _epi(**kw):
    e = _lib.Epilogue()
    for k, v in kw.items():
        setattr(e, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    return e


def _close_bf16(got, ref, what):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all(), what
    r = rel(got, ref)
    assert r < BF16_REL_L2, f"{what}: rel L2 {r:.3e}"
    bound = BF16_ELEM * ref.abs() + 2e-3 * float(ref.abs().mean() + 1e-12) + 1e-6
    worst = float(((got - ref).abs() - bound).max())
    assert worst <= 0, f"{what}: element outside one bf16 rounding by {worst:.3e}"


# =====================================================================================================
# 1. operators
# =====================================================================================================
@pytest.mark.parametrize("m,n,k", [(300, 256, 64), (1000, 2048, 2048), (257, 96, 72), (128, 64, 576), (50, 6144, 256),
                                   (19000, 128, 64)])
def test_gemm_plain(cuda, m, n, k):
    """koa_gemm_bf16 == F.linear (reference: _core_trf.py:104,145,148,161,163; 1x1 convs _torchvision.py:29-31)."""
    lib = _lib.load()
    a, b = _bf(_randn(m, k, seed=1)), _bf(_randn(n, k, seed=2, scale=k ** -0.5))
    out = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    ep = _epi(out=out, ldo=n)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm")
    _close_bf16(out, a.float() @ b.float().t(), f"gemm {m}x{n}x{k}")


@pytest.mark.parametrize("a_f16,b_f16,out_f16", [(1, 1, 1), (1, 1, 0), (0, 0, 1)])
def test_gemm_operand_formats(cuda, a_f16, b_f16, out_f16):
    """fp16 x fp16 -> fp16 is the CNN forward (11-bit significands), bf16 x bf16 -> bf16 its backward. tcgen05
    kind::f16 wants both operands in ONE format (mixing them raises an illegal-instruction fault on B200), so the
    API rejects a mixed request instead of launching it."""
    lib = _lib.load()
    m, n, k = 700, 256, 320
    a32, b32 = _randn(m, k, seed=1), _randn(n, k, seed=2, scale=k ** -0.5)
    a = a32.half() if a_f16 else a32.bfloat16()
    b = b32.half() if b_f16 else b32.bfloat16()
    out = torch.empty(m, n, dtype=torch.float16 if out_f16 else torch.bfloat16, device=cuda)
    ep = _epi(out=out, ldo=n, a_f16=a_f16, b_f16=b_f16, out_f16=out_f16)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm formats")
    ref = a.float() @ b.float().t()
    assert rel(out, ref) < (6e-4 if out_f16 else BF16_REL_L2), rel(out, ref)
    bad = _epi(out=out, ldo=n, a_f16=1, b_f16=0)
    assert lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(bad), _stream()) == -1
    assert b"same 16-bit format" in lib.koa_last_error()


def test_gemm_empty_and_bad_shapes_are_rejected(cuda):
    lib = _lib.load()
    a = torch.zeros(8, 64, dtype=torch.bfloat16, device=cuda)
    out = torch.zeros(8, 64, dtype=torch.bfloat16, device=cuda)
    ep = _epi(out=out, ldo=64)
    assert lib.koa_gemm_bf16(a.data_ptr(), a.data_ptr(), 0, 64, 64, C.byref(ep), _stream()) == -1   # empty
    assert lib.koa_gemm_bf16(a.data_ptr(), a.data_ptr(), 8, 40, 64, C.byref(ep), _stream()) == -1   # N % 32
    assert lib.koa_gemm_bf16(a.data_ptr(), a.data_ptr(), 8, 64, 60, C.byref(ep), _stream()) == -1   # K % 8
    assert b"multiple" in lib.koa_last_error()
    ep2 = _epi(out=None, ldo=64)
    assert lib.koa_gemm_bf16(a.data_ptr(), a.data_ptr(), 8, 64, 64, C.byref(ep2), _stream()) == -1  # NULL out


def test_gemm_fused_epilogues(cuda):
    """bias + GELU with the pre-activation copy (ff.net.0), fp32 output + fp32 residual + bf16 copy (to_out / ff.net.3
    on the residual stream), GELU' gating (backward of ff.net.0), bf16 addend + ReLU gate (identity path of a residual
    block in backward) with the fused BatchNorm-backward reductions, BatchNorm column statistics (conv forward)."""
    lib = _lib.load()
    m, n, k = 333, 256, 192
    a, b = _bf(_randn(m, k, seed=3)), _bf(_randn(n, k, seed=4, scale=k ** -0.5))
    acc = a.float() @ b.float().t()
    bias = _randn(n, seed=5, scale=0.3)
    # bias + GELU, pre-activation copy
    out = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    pre = torch.empty_like(out)
    ep = _epi(out=out, ldo=n, act=_lib.ACT_GELU, bias=bias, pre_out_bf16=pre)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm gelu")
    _close_bf16(pre, acc + bias, "pre-activation")
    _close_bf16(out, F.gelu(acc + bias), "gelu")
    # fp32 out + fp32 residual + bf16 copy
    res = _randn(m, n, seed=6)
    out32 = torch.empty(m, n, dtype=torch.float32, device=cuda)
    cp = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    ep = _epi(out=out32, ldo=n, out_fp32=1, bias=bias, residual_f32=res, out_bf16_copy=cp)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm res")
    ref = acc + bias + res
    assert rel(out32, ref) < 1e-5
    _close_bf16(cp, ref, "bf16 copy")
    # GELU' gating: out = acc * gelu'(h)
    h = _bf(_randn(m, n, seed=7))
    ep = _epi(out=out, ldo=n, act=_lib.ACT_GELU_GRAD, aux_bf16=h)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm gelu'")
    hf = h.float().requires_grad_(True)
    (gp,) = torch.autograd.grad(F.gelu(hf).sum(), hf)
    _close_bf16(out, acc * gp, "gelu grad")
    # addend + ReLU-backward gate: out = (acc + add) * (gate > 0)
    add, gate = _bf(_randn(m, n, seed=8)), _bf(_randn(m, n, seed=9).relu())
    ep = _epi(out=out, ldo=n, add_bf16=add, gate_bf16=gate)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm add")
    _close_bf16(out, (acc + add.float()) * (gate.float() > 0), "gated add")
    # the same with the BatchNorm-backward reductions of the stored values: sum(dz), sum(dz * xhat)
    y = _bf(_randn(m, n, seed=10) * 2 + 0.5)
    mean, invstd = _randn(n, seed=11, scale=0.3), torch.rand(n, generator=torch.Generator().manual_seed(12)).to(cuda) + 0.5
    s = torch.zeros(n, device=cuda)
    q = torch.zeros(n, device=cuda)
    ep = _epi(out=out, ldo=n, add_bf16=add, gate_bf16=gate, col_sum=s, col_sumsq=q, stat_y=y, stat_mean=mean,
              stat_invstd=invstd)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm bn-bwd stats")
    _close_bf16(out, (acc + add.float()) * (gate.float() > 0), "gated add (stats)")
    dz = out.float()
    xhat = (y.float() - mean) * invstd
    assert torch.allclose(s, dz.sum(0), rtol=1e-4, atol=1e-3)
    assert torch.allclose(q, (dz * xhat).sum(0), rtol=1e-4, atol=2e-3)
    # ReLU + column statistics of the stored (rounded) values
    s = torch.zeros(n, device=cuda)
    q = torch.zeros(n, device=cuda)
    ep = _epi(out=out, ldo=n, col_sum=s, col_sumsq=q)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm stats")
    _close_bf16(out, acc, "stats out")
    assert torch.allclose(s, out.float().sum(0), rtol=1e-4, atol=1e-3)
    assert torch.allclose(q, (out.float() ** 2).sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("m,n,k", [(40000, 256, 64), (5000, 64, 576), (777, 2048, 128), (128 * 148 * 3 + 5, 128, 64)])
def test_gemm_column_statistics_many_tiles(cuda, m, n, k):
    """Forward BatchNorm statistics fused in the epilogue over many tiles per CTA (persistent kernel, per-CTA running
    totals in shared memory, one global atomic per column and CTA), ragged last row tile."""
    lib = _lib.load()
    a, b = _bf(_randn(m, k, seed=3)), _bf(_randn(n, k, seed=4, scale=k ** -0.5))
    out = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    s, q = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ep = _epi(out=out, ldo=n, col_sum=s, col_sumsq=q)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "gemm stats")
    _close_bf16(out, a.float() @ b.float().t(), "stats out")
    of = out.double()
    assert torch.allclose(s.double(), of.sum(0), rtol=1e-4, atol=1e-2 * (m ** 0.5) * 1e-1)
    assert torch.allclose(q.double(), (of ** 2).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("m,n,k", [(333, 256, 64), (40000, 64, 256), (6, 1024, 256), (128 * 148 * 2 + 77, 128, 64),
                                   (5000, 512, 128), (96, 96, 64), (2500, 2048, 512), (3000, 128, 1024), (1000, 768, 512),
                                   (257, 256, 576), (128 * 74 * 4 + 300, 256, 512)])
@pytest.mark.parametrize("act_f16", [0, 1])
def test_gemm_conv_backward_epilogue(cuda, m, n, k, act_f16):
    """Data-gradient epilogue of a residual block (fe_engine.cu: G = (dgrad + G_b) * (x > 0) with the fused
    BatchNorm-backward sums; reference autograd of _torchvision.py:118-138): bf16 dy / weights / output, forward
    activations (gate, y) in fp16 or bf16, ragged row tiles, one n-tile per CTA with 1..16 n-tiles, in-place addend."""
    lib = _lib.load()
    a, b = _bf(_randn(m, k, seed=3)), _bf(_randn(n, k, seed=4, scale=k ** -0.5))
    acc = a.float() @ b.float().t()
    adt = torch.float16 if act_f16 else torch.bfloat16
    add = _bf(_randn(m, n, seed=8))
    gate = _randn(m, n, seed=9).relu().to(adt).contiguous()
    y = (_randn(m, n, seed=10) * 2 + 0.5).to(adt).contiguous()
    mean = _randn(n, seed=11, scale=0.3)
    invstd = torch.rand(n, generator=torch.Generator().manual_seed(12)).to(cuda) + 0.5
    ref = (acc + add.float()) * (gate.float() > 0)
    xhat = (y.float() - mean) * invstd
    # gate + addend + BatchNorm-backward sums, written in place over the addend
    out = add.clone()
    s, q = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ep = _epi(out=out, ldo=n, add_bf16=out, gate_bf16=gate, col_sum=s, col_sumsq=q, stat_y=y, stat_mean=mean,
              stat_invstd=invstd, act_f16=act_f16)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "conv bwd epilogue")
    _close_bf16(out, ref, "in-place gated add")
    dz = out.double()
    tol = 2e-4 * float(dz.abs().sum(0).max()) + 1e-3
    assert float((s.double() - dz.sum(0)).abs().max()) < tol
    assert float((q.double() - (dz * xhat.double()).sum(0)).abs().max()) < 4 * tol
    # gate only + sums (dgrad of conv3 / conv2), and the plain epilogue without operands
    out2 = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    s.zero_(); q.zero_()
    ep = _epi(out=out2, ldo=n, gate_bf16=gate, col_sum=s, col_sumsq=q, stat_y=y, stat_mean=mean, stat_invstd=invstd,
              act_f16=act_f16)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "conv bwd epilogue 2")
    _close_bf16(out2, acc * (gate.float() > 0), "gated")
    dz = out2.double()
    assert float((s.double() - dz.sum(0)).abs().max()) < tol
    assert float((q.double() - (dz * xhat.double()).sum(0)).abs().max()) < 4 * tol
    ep = _epi(out=out2, ldo=n, add_bf16=add)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "conv bwd epilogue 3")
    _close_bf16(out2, acc + add.float(), "add only")
    assert _lib.debug_flag() == 0


@pytest.mark.parametrize("m,n,k", [(333, 256, 64), (40000, 64, 64), (6, 1024, 256), (128 * 148 * 2 + 77, 128, 64), (96, 96, 128),
                                   (3000, 128, 1024), (1000, 768, 512), (257, 256, 576), (128 * 74 * 4 + 300, 512, 512)])
def test_gemm_conv_forward_epilogue_fp16(cuda, m, n, k):
    """Forward convolution flavour as the extractor runs it (fe_engine.cu conv_forward): fp16 operands and output with
    the BatchNorm batch statistics (sum, sum of squares) of the stored values."""
    lib = _lib.load()
    a = _randn(m, k, seed=3).half().contiguous()
    b = _randn(n, k, seed=4, scale=k ** -0.5).half().contiguous()
    out = torch.empty(m, n, dtype=torch.float16, device=cuda)
    s, q = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ep = _epi(out=out, ldo=n, col_sum=s, col_sumsq=q, a_f16=1, b_f16=1, out_f16=1)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "conv fwd epilogue")
    ref = a.float() @ b.float().t()
    assert rel(out.float(), ref) < 6e-4
    of = out.double()
    assert torch.allclose(s.double(), of.sum(0), rtol=1e-4, atol=2e-4 * float(of.abs().sum(0).max()) + 1e-3)
    assert torch.allclose(q.double(), (of ** 2).sum(0), rtol=1e-4, atol=1e-2)
    assert _lib.debug_flag() == 0



@pytest.mark.parametrize("m,n,k", [(333, 256, 64), (40000, 64, 64), (6, 1024, 256), (128 * 148 * 2 + 77, 128, 64), (96, 96, 128),
                                   (3000, 128, 1024), (1000, 768, 512), (257, 2048, 512), (128 * 74 * 4 + 300, 512, 512)])
def test_gemm_conv_bn_apply_epilogue(cuda, m, n, k):
    """conv -> bn -> (+ identity | + bn_d(downsample)) -> relu as ONE kernel (reference _torchvision.py:118-138; here
    gemm_conv.cuh MODE 2, used by inference and by the y-free tail of a train-mode bottleneck): fp16 operands, per-column
    scale / shift, fp16 residual with its own optional per-column affine map, fp16 output + bf16 copy; single CTAs and
    CTA pairs, 1..16 n-tiles, ragged row tiles."""
    lib = _lib.load()
    a = _randn(m, k, seed=3).half().contiguous()
    b = _randn(n, k, seed=4, scale=k ** -0.5).half().contiguous()
    acc = a.float() @ b.float().t()
    scale = torch.rand(n, generator=torch.Generator().manual_seed(5)).to(cuda) + 0.5
    shift = _randn(n, seed=6, scale=0.3)
    res = _randn(m, n, seed=7).half().contiguous()
    rs = torch.rand(n, generator=torch.Generator().manual_seed(8)).to(cuda) + 0.5
    rt = _randn(n, seed=9, scale=0.3)
    out = torch.empty(m, n, dtype=torch.float16, device=cuda)
    cp = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    fmt = dict(a_f16=1, b_f16=1, out_f16=1, act_f16=1)
    # bn + relu (conv1 / conv2 in inference)
    ep = _epi(out=out, ldo=n, bn_scale=scale, bn_shift=shift, act=_lib.ACT_RELU, **fmt)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "bn relu")
    ref = (acc * scale + shift).relu()
    assert rel(out.float(), ref) < 6e-4, rel(out.float(), ref)
    # bn, no activation (downsample branch in inference)
    ep = _epi(out=out, ldo=n, bn_scale=scale, bn_shift=shift, **fmt)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "bn")
    assert rel(out.float(), acc * scale + shift) < 6e-4
    # bn + identity + relu, with the bf16 copy (last convolution of an identity block)
    ep = _epi(out=out, ldo=n, bn_scale=scale, bn_shift=shift, add_bf16=res, act=_lib.ACT_RELU, out_bf16_copy=cp, **fmt)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "bn res relu")
    ref = (acc * scale + shift + res.float()).relu()
    assert rel(out.float(), ref) < 6e-4
    _close_bf16(cp, ref, "bf16 copy of the block output")
    # bn + bn_d(raw downsample output) + relu (first block of a stage, train mode)
    ep = _epi(out=out, ldo=n, bn_scale=scale, bn_shift=shift, add_bf16=res, res_scale=rs, res_shift=rt, act=_lib.ACT_RELU, **fmt)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "bn res-affine relu")
    ref = (acc * scale + shift + res.float() * rs + rt).relu()
    assert rel(out.float(), ref) < 6e-4
    # in place over the residual (not used by the engine, but nothing forbids it)
    io = res.clone()
    ep = _epi(out=io, ldo=n, bn_scale=scale, bn_shift=shift, add_bf16=io, act=_lib.ACT_RELU, **fmt)
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), _stream()), "bn res in place")
    assert rel(io.float(), (acc * scale + shift + res.float()).relu()) < 6e-4
    # a request the flavour cannot serve is rejected, not silently served by another kernel
    bad = _epi(out=cp, ldo=n, bn_scale=scale, bn_shift=shift)
    assert lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(bad), _stream()) == -1
    assert _lib.debug_flag() == 0


@pytest.mark.parametrize("n_img,h,w,cin,cout,k,stride", [(3, 20, 20, 64, 64, 3, 1), (2, 16, 16, 128, 128, 3, 2),
                                                         (2, 22, 22, 256, 512, 1, 2)])
def test_conv_bn_apply_epilogue(cuda, n_img, h, w, cin, cout, k, stride):
    """The same epilogue behind the implicit-GEMM convolutions (3x3 stride 1 / 2, strided 1x1 downsample)."""
    lib = _lib.load()
    pad = k // 2
    x = _randn(n_img, cin, h, w, seed=1).half()
    wt = _randn(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5).half()
    scale = torch.rand(cout, generator=torch.Generator().manual_seed(5)).to(cuda) + 0.5
    shift = _randn(cout, seed=6, scale=0.3)
    y = F.conv2d(x.float(), wt.float(), stride=stride, padding=pad)
    ref = (y * scale[None, :, None, None] + shift[None, :, None, None]).relu()
    xn = _nhwc(x).half().contiguous()
    wp = wt.permute(0, 2, 3, 1).contiguous()
    out = torch.empty(ref.shape[0] * ref.shape[2] * ref.shape[3], cout, dtype=torch.float16, device=cuda)
    ep = _epi(out=out, ldo=cout, bn_scale=scale, bn_shift=shift, act=_lib.ACT_RELU, a_f16=1, b_f16=1, out_f16=1, act_f16=1)
    _lib.check(lib.koa_conv_fprop_bf16(xn.data_ptr(), wp.data_ptr(), n_img, h, w, cin, cout, k, k, stride, pad, C.byref(ep),
                                       _stream()), "conv bn relu")
    assert rel(out.float(), ref.permute(0, 2, 3, 1).reshape(-1, cout)) < 6e-4
    assert _lib.debug_flag() == 0


@pytest.mark.parametrize("m,k1,k2,n", [(333, 256, 64, 64), (40000, 256, 64, 64), (5000, 512, 128, 128), (2500, 2048, 512, 512),
                                       (128 * 74 * 4 + 300, 1024, 256, 256), (700, 2048, 1024, 1024)])
def test_gemm_kcat_bias_gate_stats(cuda, m, k1, k2, n):
    """out = (([A1 | A2] . B^T + bias) * (gate > 0)) with the BatchNorm-backward sums: the data gradient of the y-free
    bottleneck tail (fe_engine.cu; autograd of _torchvision.py:130-133), K split over two tensors."""
    lib = _lib.load()
    a1, a2 = _bf(_randn(m, k1, seed=1)), _bf(_randn(m, k2, seed=2))
    b = _bf(_randn(n, k1 + k2, seed=3, scale=(k1 + k2) ** -0.5))
    bias = _randn(n, seed=4, scale=0.2)
    gate = _randn(m, n, seed=5).relu().half().contiguous()
    y = (_randn(m, n, seed=6) * 2 + 0.5).half().contiguous()
    mean = _randn(n, seed=7, scale=0.3)
    invstd = torch.rand(n, generator=torch.Generator().manual_seed(8)).to(cuda) + 0.5
    acc = torch.cat([a1, a2], 1).float() @ b.float().t() + bias
    out = torch.empty(m, n, dtype=torch.bfloat16, device=cuda)
    s, q = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ep = _epi(out=out, ldo=n, col_bias=bias, gate_bf16=gate, col_sum=s, col_sumsq=q, stat_y=y, stat_mean=mean,
              stat_invstd=invstd, act_f16=1)
    _lib.check(lib.koa_gemm_kcat_bf16(a1.data_ptr(), k1, a2.data_ptr(), k2, b.data_ptr(), m, n, C.byref(ep), _stream()), "kcat")
    _close_bf16(out, acc * (gate.float() > 0), "kcat gated")
    dz = out.double()
    xhat = ((y.float() - mean) * invstd).double()
    tol = 2e-4 * float(dz.abs().sum(0).max()) + 1e-3
    assert float((s.double() - dz.sum(0)).abs().max()) < tol
    assert float((q.double() - (dz * xhat).sum(0)).abs().max()) < 4 * tol
    # plain concatenation, no epilogue operands
    ep = _epi(out=out, ldo=n)
    _lib.check(lib.koa_gemm_kcat_bf16(a1.data_ptr(), k1, a2.data_ptr(), k2, b.data_ptr(), m, n, C.byref(ep), _stream()), "kcat plain")
    _close_bf16(out, acc - bias, "kcat plain")
    assert lib.koa_gemm_kcat_bf16(a1.data_ptr(), 72, a2.data_ptr(), k2, b.data_ptr(), m, n, C.byref(ep), _stream()) == -1
    assert _lib.debug_flag() == 0


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("n_img,h,w,cin,cout,k,stride,pad", [
    (3, 20, 20, 64, 64, 3, 1, 1), (2, 20, 20, 128, 128, 3, 2, 1), (5, 10, 10, 256, 512, 1, 2, 0),
    (2, 40, 40, 64, 256, 3, 1, 1), (7, 5, 5, 512, 512, 3, 1, 1), (1, 6, 10, 64, 96, 3, 1, 1)])
def test_conv_fprop_and_wgrad(cuda, n_img, h, w, cin, cout, k, stride, pad):
    """Implicit-GEMM convolution and its weight gradient == nn.Conv2d / autograd (_torchvision.py:23-31,110,211-213)."""
    lib = _lib.load()
    x = _bf(_randn(n_img, cin, h, w, seed=11))
    wt = _bf(_randn(cout, cin, k, k, seed=12, scale=(cin * k * k) ** -0.5))
    xf, wf = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    ref = F.conv2d(xf, wf, stride=stride, padding=pad)
    ho, wo = ref.shape[2], ref.shape[3]
    x_nhwc = _nhwc(x)
    w_pack = wt.permute(0, 2, 3, 1).contiguous()  # [Cout][R][S][Cin]
    out = torch.empty(n_img, ho, wo, cout, dtype=torch.bfloat16, device=cuda)
    ep = _epi(out=out, ldo=cout)
    _lib.check(lib.koa_conv_fprop_bf16(x_nhwc.data_ptr(), w_pack.data_ptr(), n_img, h, w, cin, cout, k, k, stride, pad,
                                       C.byref(ep), _stream()), "conv fprop")
    _close_bf16(out, _nhwc(ref.detach()), "conv fprop")
    # weight gradient
    dy = _bf(_randn(n_img, cout, ho, wo, seed=13))
    (gw,) = torch.autograd.grad(ref, wf, dy.float())
    dw = torch.zeros(cout, k, k, cin, dtype=torch.float32, device=cuda)
    dy_nhwc = _nhwc(dy)
    _lib.check(lib.koa_conv_wgrad_bf16(dy_nhwc.data_ptr(), x_nhwc.data_ptr(), dw.data_ptr(), n_img, h, w, cin, cout, k, k,
                                       stride, pad, 0, _stream()), "conv wgrad")
    assert rel(dw, gw.permute(0, 2, 3, 1)) < 1e-4
    # the same from the fp16 activation, converted to bf16 inside the kernel (x_f16 = 2; im2col operand)
    dw.zero_()
    x16 = (x_nhwc.float() * 1.37).half().contiguous()  # generic fp16 values: the kernel rounds them to bf16
    (gw16,) = torch.autograd.grad(F.conv2d(x16.bfloat16().float().permute(0, 3, 1, 2), wf, stride=stride, padding=pad), wf,
                                  dy.float())
    _lib.check(lib.koa_conv_wgrad_bf16(dy_nhwc.data_ptr(), x16.data_ptr(), dw.data_ptr(), n_img, h, w, cin, cout, k, k,
                                       stride, pad, 2, _stream()), "conv wgrad xcvt")
    assert rel(dw, gw16.permute(0, 2, 3, 1)) < 1e-4
    # the CNN's forward formats: fp16 activations / weights in, fp16 out
    xh, wh = x.float().half(), wt.float().half()
    ref_h = F.conv2d(xh.float(), wh.float(), stride=stride, padding=pad)
    out_h = torch.empty(n_img, ho, wo, cout, dtype=torch.float16, device=cuda)
    ep = _epi(out=out_h, ldo=cout, a_f16=1, b_f16=1, out_f16=1)
    xh_nhwc, wh_pack = _nhwc(xh), wh.permute(0, 2, 3, 1).contiguous()  # keep the operands alive across the call
    _lib.check(lib.koa_conv_fprop_bf16(xh_nhwc.data_ptr(), wh_pack.data_ptr(), n_img, h, w, cin, cout, k, k, stride, pad,
                                       C.byref(ep), _stream()), "conv fprop fp16")
    assert rel(out_h, _nhwc(ref_h)) < 6e-4  # one fp16 rounding of the output


@pytest.mark.parametrize("pixels,cout,cin", [(1000, 256, 64), (64, 64, 64), (4133, 512, 128), (100, 2048, 512)])
def test_gemm_wgrad(cuda, pixels, cout, cin):
    lib = _lib.load()
    dy, x = _bf(_randn(pixels, cout, seed=21)), _bf(_randn(pixels, cin, seed=22))
    dw = torch.zeros(cout, cin, dtype=torch.float32, device=cuda)
    _lib.check(lib.koa_gemm_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), pixels, cout, cin, 0, _stream()), "wgrad")
    assert rel(dw, dy.float().t() @ x.float()) < 1e-4
    # accumulates (split-K partial sums land with atomics on top of what is there)
    _lib.check(lib.koa_gemm_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), pixels, cout, cin, 0, _stream()), "wgrad")
    assert rel(dw, 2 * (dy.float().t() @ x.float())) < 1e-4
    # fp16 x fp16 also works (x_f16 covers both operands); mixing formats is not offered
    xh, dyh = x.float().half(), dy.float().half()
    dw.zero_()
    _lib.check(lib.koa_gemm_wgrad_bf16(dyh.data_ptr(), xh.data_ptr(), dw.data_ptr(), pixels, cout, cin, 1, _stream()), "wgrad")
    assert rel(dw, dyh.float().t() @ xh.float()) < 1e-4
    # x_f16 = 2: bf16 dy with the fp16 forward activation, converted to bf16 in shared memory inside the kernel (what the
    # extractor backward runs): equals the product with the bf16-rounded activation
    xh = (_randn(pixels, cin, seed=23) * 1.7).half().contiguous()
    dw.zero_()
    _lib.check(lib.koa_gemm_wgrad_bf16(dy.data_ptr(), xh.data_ptr(), dw.data_ptr(), pixels, cout, cin, 2, _stream()), "wgrad")
    assert rel(dw, dy.float().t() @ xh.bfloat16().float()) < 1e-4
    assert _lib.debug_flag() == 0


@pytest.mark.parametrize("rows,d", [(7, 2048), (1472, 2048), (33, 256)])
def test_layernorm(cuda, rows, d):
    """nn.LayerNorm(eps 1e-5) forward / backward (_core_trf.py:111,190,192)."""
    lib = _lib.load()
    x = _randn(rows, d, seed=31, scale=2.0) + 0.5
    gamma, beta = _randn(d, seed=32) * 0.2 + 1.0, _randn(d, seed=33) * 0.1
    o16 = torch.empty(rows, d, dtype=torch.bfloat16, device=cuda)
    o32 = torch.empty(rows, d, device=cuda)
    mean, rstd = torch.empty(rows, device=cuda), torch.empty(rows, device=cuda)
    _lib.check(lib.koa_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), o16.data_ptr(), o32.data_ptr(),
                                     mean.data_ptr(), rstd.data_ptr(), rows, d, _stream()), "ln fwd")
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (d,), gr, br, 1e-5)
    assert rel(o32, ref) < 1e-5
    _close_bf16(o16, ref.detach(), "layernorm bf16 copy")
    dy = _randn(rows, d, seed=34)
    ref.backward(dy)
    dx = torch.empty_like(x)
    dg, db = torch.zeros(d, device=cuda), torch.zeros(d, device=cuda)
    _lib.check(lib.koa_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, d, _stream()), "ln bwd")
    assert rel(dx, xr.grad) < 1e-4 and rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("f16", [0, 1])
@pytest.mark.parametrize("b,n,heads,hd", [(2, 65, 8, 256), (3, 25, 8, 256), (1, 124, 8, 256), (16, 92, 8, 256), (2, 128, 8, 256),
                                          (3, 64, 2, 128), (2, 7, 4, 64), (1, 1, 8, 256), (1, 1, 8, 32), (2, 33, 4, 96)])
def test_attention(cuda, b, n, heads, hd, f16):
    """softmax(Q K^T * scale) V with the reference's (qkv, head, d) feature split and model-dim scale
    (_core_trf.py:160,167-182), forward + backward: the tcgen05 kernels (head_dim a multiple of 64: one UMMA tile per
    contraction, probabilities / dS rounded to 16 bit as MMA operands) and the CUDA-core kernels (other head sizes), with
    fp16 or bf16 forward tensors; single-token, 25 / 64 / 92 / 124 / 128-token sequences (the last rows of a 128-row TMA box
    belong to the next sequence or lie past the end of the matrix)."""
    lib = _lib.load()
    d = heads * hd
    scale = float(d) ** -0.5
    fdt = torch.float16 if f16 else torch.bfloat16
    qkv = _randn(b * n, 3 * d, seed=41).to(fdt).contiguous()
    out = torch.empty(b * n, d, dtype=fdt, device=cuda)
    probs = torch.empty(b, heads, n, n, device=cuda)
    _lib.check(lib.koa_attention_fwd_fmt(qkv.data_ptr(), out.data_ptr(), probs.data_ptr(), b, n, heads, hd, scale, f16,
                                         _stream()), "attn fwd")
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.view(b, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    pr = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    ref = (pr @ v).permute(0, 2, 1, 3).reshape(b * n, d)
    assert torch.isfinite(probs).all() and rel(probs, pr) < 1e-4, rel(probs, pr)
    assert rel(out.float(), ref.detach()) < (1.5e-3 if f16 else BF16_REL_L2), rel(out.float(), ref.detach())
    dout = _bf(_randn(b * n, d, seed=42))
    ref.backward(dout.float())
    dqkv = torch.full_like(qkv, float("nan"), dtype=torch.bfloat16)
    _lib.check(lib.koa_attention_bwd_fmt(qkv.data_ptr(), probs.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), b, n, heads, hd,
                                         scale, f16, _stream()), "attn bwd")
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        e = rel(dqkv[:, sl].float(), qf.grad[:, sl])
        assert e < 6e-3, (name, e)
    if f16 == 0:   # the original entry points are the bf16 form of the same call
        out2 = torch.empty_like(out)
        _lib.check(lib.koa_attention_fwd(qkv.data_ptr(), out2.data_ptr(), probs.data_ptr(), b, n, heads, hd, scale, _stream()),
                   "attn fwd (bf16 entry point)")
        assert torch.equal(out2, out)
    assert _lib.debug_flag() == 0


@pytest.mark.parametrize("n,h,w", [(3, 32, 32), (2, 175, 175), (1, 80, 80)])
def test_maxpool(cuda, n, h, w):
    """nn.MaxPool2d(3, 2, 1) (_torchvision.py:174) incl. the odd 175 -> 88 XR size; bit exact (pure selection)."""
    lib = _lib.load()
    c = 64
    x = _randn(n, c, h, w, seed=51).half()  # forward activations of the CNN are fp16, gradients bf16
    ref, _ = F.max_pool2d(x.float(), 3, 2, 1, return_indices=True)
    ho, wo = ref.shape[2], ref.shape[3]
    out = torch.empty(n, ho, wo, c, dtype=torch.float16, device=cuda)
    idx = torch.empty(n, ho, wo, c, dtype=torch.uint8, device=cuda)
    x_nhwc = _nhwc(x)
    _lib.check(lib.koa_maxpool_fwd(x_nhwc.data_ptr(), out.data_ptr(), idx.data_ptr(), n, h, w, c, _stream()), "pool")
    assert torch.equal(out.float(), _nhwc(ref))
    dout = _bf(_randn(n, c, ho, wo, seed=52))
    xf = x.float().requires_grad_(True)
    F.max_pool2d(xf, 3, 2, 1).backward(dout.float())
    dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=cuda)
    dout_nhwc = _nhwc(dout)
    _lib.check(lib.koa_maxpool_bwd(dout_nhwc.data_ptr(), idx.data_ptr(), dx.data_ptr(), n, h, w, c, _stream()), "pool bwd")
    _close_bf16(dx, _nhwc(xf.grad), "maxpool bwd")  # windows overlap: up to 4 bf16 addends per input pixel


@pytest.mark.parametrize("batch", [1, 16, 257])
def test_focal_loss(cuda, batch):
    """FocalLoss(gamma 2, mean) and d loss / d logits (various/_losses.py:89-108)."""
    lib = _lib.load()
    logits = _randn(batch, 2, seed=61, scale=2.0)
    target = (torch.rand(batch, generator=torch.Generator().manual_seed(62)) < 0.3).long().to(cuda)
    loss = torch.empty((), device=cuda)
    dl = torch.empty_like(logits)
    _lib.check(lib.koa_focal_loss(logits.data_ptr(), target.data_ptr(), loss.data_ptr(), dl.data_ptr(), batch, 2, 2.0,
                                  _stream()), "focal")
    lr = logits.clone().requires_grad_(True)
    ref = ko.focal_loss(lr, target)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-6 + 1e-5 * abs(float(ref))
    assert rel(dl, lr.grad) < 1e-5


def test_channel_dropout_is_dropout2d(cuda):
    """nn.Dropout2d on the extractor output (reference _xrNmrMcP.py:62-74,226-229: applied to (B*S, C, h, w)) on the
    engine's token layout: ONE draw per (slice, channel) shared by all spatial positions of the slice (not one per
    position), scale 1 / (1 - p), the mask replayed by koa_dropout_mask, backward = the same mask on the gradient,
    identity in eval mode; and the model classes call it that way (with and without the global average pool)."""
    from oaprogressionmmf_b200.koamodels import _models as km

    lib = _lib.load()
    n_img, pos, c, p = 12, 4, 256, 0.3
    x = (_randn(3, 4 * pos, c, seed=1).abs() + 0.1).requires_grad_(True)   # (B = 3, S * positions = 16, C): 4 slices per knee
    drop = torch.nn.Dropout2d(p)
    torch.manual_seed(5)
    y = km._apply_drop2d(drop, x, n_img, site=2)
    mask = torch.empty(n_img, c, device=cuda)
    _lib.check(lib.koa_dropout_mask(drop.last_seed, drop.last_site, n_img, c, p, mask.data_ptr(), _stream()), "mask")
    assert all(v == 0.0 or abs(v - 1 / (1 - p)) < 1e-6 for v in mask.unique().tolist())
    keep = float((mask > 0).float().mean())
    assert abs(keep - (1 - p)) < 0.05, keep
    want = x.detach().reshape(n_img, pos, c) * mask[:, None, :]
    assert torch.equal(y.detach().reshape(n_img, pos, c), want)          # every position of a slice shares the draw
    gy = _randn(*y.shape, seed=2)
    y.backward(gy)
    assert torch.equal(x.grad.reshape(n_img, pos, c), gy.reshape(n_img, pos, c) * mask[:, None, :])
    torch.manual_seed(5)
    assert torch.equal(km._apply_drop2d(drop, x.detach(), n_img, site=2), y.detach())   # torch.manual_seed replays it
    drop.eval()
    assert km._apply_drop2d(drop, x, n_img) is x
    # distribution matches nn.Dropout2d on the reference's (B*S, C, h, w) layout: zero fraction per (slice, channel)
    ref = torch.nn.Dropout2d(p)(torch.ones(4000, 64, 2, 2, device=cuda))
    assert abs(float((ref[:, :, 0, 0] == 0).float().mean()) - p) < 0.02
    assert bool((ref == ref[:, :, :1, :1]).all())


def test_focal_loss_edge_cases(cuda):
    """What F.cross_entropy-based FocalLoss of the reference does at the edges (various/_losses.py:89-108): a target
    outside [0, classes) is an error there (device assert); here it is a NaN loss plus the diagnostic word, never a
    silent out-of-bounds read. gamma < 1 with a saturated prediction (pt -> 1) has a finite (zero) gradient.
    reduction='sum' and the rejected class_weight."""
    from oaprogressionmmf_b200.losses import FocalLoss

    logits = torch.tensor([[0.3, -0.2], [40.0, -40.0], [-1.0, 2.0]], device=cuda, requires_grad=True)
    target = torch.tensor([[0], [0], [1]], device=cuda)
    for gamma in (0.5, 2.0):
        ref_in = logits.detach().clone().requires_grad_(True)
        logpt = -F.cross_entropy(ref_in, target.reshape(-1), reduction="none")
        ref = (-((1 - logpt.exp()) ** gamma) * logpt)
        for red in ("mean", "sum"):
            lg = logits.detach().clone().requires_grad_(True)
            loss = FocalLoss(gamma=gamma, reduction=red)(lg, target)
            loss.backward()
            want = ref.mean() if red == "mean" else ref.sum()
            assert torch.isfinite(lg.grad).all(), (gamma, red)
            assert abs(float(loss) - float(want)) < 1e-6 + 1e-5 * abs(float(want))
            assert float(lg.grad[1].abs().max()) < 1e-6   # saturated sample: no gradient, no NaN
    # the loss is bit-reproducible (fixed-order reduction)
    big = _randn(4096, 2, seed=3)
    tg = (torch.arange(4096, device=cuda) % 2).reshape(-1, 1)
    l0 = float(FocalLoss()(big, tg))
    assert all(float(FocalLoss()(big, tg)) == l0 for _ in range(5))
    assert _lib.debug_flag() == 0
    bad = torch.tensor([[0], [2], [1]], device=cuda)
    loss = FocalLoss()(logits, bad)
    assert torch.isnan(loss)
    assert _lib.debug_flag() == 0xF0CA1
    assert _lib.debug_flag() == 0   # read clears
    with pytest.raises(ValueError):
        FocalLoss(class_weight=torch.ones(2))
    with pytest.raises(ValueError):
        FocalLoss(reduction="max")
    with pytest.raises(_lib.KoaError):
        FocalLoss()(logits.detach().cpu(), target.cpu())


@pytest.mark.parametrize("m,n,k,act", [(16, 2048, 9, _lib.ACT_GELU), (8, 512, 2048, _lib.ACT_RELU), (5, 2, 512, _lib.ACT_NONE)])
def test_linear_small(cuda, m, n, k, act):
    """FeatC1 Linear(9 -> 2048)+GELU (_xrNmrMcP.py:15-19), XR1Cnn head (_xr1_cnn.py:31-39)."""
    lib = _lib.load()
    x, w, b = _randn(m, k, seed=71), _randn(n, k, seed=72, scale=k ** -0.5), _randn(n, seed=73, scale=0.1)
    y, pre = torch.empty(m, n, device=cuda), torch.empty(m, n, device=cuda)
    _lib.check(lib.koa_linear_small_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), pre.data_ptr(), m, n, k, act,
                                        _stream()), "small fwd")
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = F.linear(xr, wr, br)
    ref = {_lib.ACT_GELU: F.gelu, _lib.ACT_RELU: F.relu, _lib.ACT_NONE: lambda t: t}[act](z)
    assert rel(y, ref) < 1e-5
    dy = _randn(m, n, seed=74)
    ref.backward(dy)
    dx, dw, db = torch.empty_like(x), torch.zeros_like(w), torch.zeros_like(b)
    scratch = torch.empty(m, n, device=cuda)
    _lib.check(lib.koa_linear_small_bwd(dy.data_ptr(), pre.data_ptr(), x.data_ptr(), w.data_ptr(), scratch.data_ptr(),
                                        dx.data_ptr(), dw.data_ptr(), db.data_ptr(), m, n, k, act, _stream()), "small bwd")
    assert rel(dx, xr.grad) < 1e-4 and rel(dw, wr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


def test_stem_pack_is_the_einops_rearrange(cuda):
    """'b ch r c s -> (b s) ch r c' (_xrNmrMcP.py:209-210): bit exact (pure data movement), ragged slice counts."""
    lib = _lib.load()
    for b, r, c, s in [(2, 16, 16, 5), (1, 32, 32, 25), (3, 8, 12, 1)]:
        vol = _randn(b, 1, r, c, s, seed=81)
        img = torch.empty(b * s, r, c, device=cuda)
        _lib.check(lib.koa_stem_pack(vol.data_ptr(), img.data_ptr(), b, r * c, s, _stream()), "stem pack")
        assert torch.equal(img, ko._slices_to_images(vol)[:, 0])


def test_col_stats(cuda):
    lib = _lib.load()
    y = _bf(_randn(5000, 256, seed=91) + 0.3)
    s, q = torch.zeros(256, device=cuda), torch.zeros(256, device=cuda)
    _lib.check(lib.koa_col_stats(y.data_ptr(), s.data_ptr(), q.data_ptr(), 5000, 256, _stream()), "col stats")
    assert torch.allclose(s, y.float().sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(q, (y.float() ** 2).sum(0), rtol=1e-4, atol=1e-2)


# =====================================================================================================
# 2. engines
# =====================================================================================================
def _fe_pair(arch, dev, seed=11, res_gain=1.0, with_gap=True):
    from oaprogressionmmf_b200.koamodels import SliceEncoder, dict_fes

    spec = ko.fe_param_spec(arch, "_fe")
    sd = ko.make_state_dict(spec, seed, device=dev, res_gain=res_gain)
    enc = SliceEncoder(dict_fes[arch](pretrained=False), with_gap=with_gap).to(dev)
    enc.load_state_dict({k[len("_fe."):]: v.clone() for k, v in sd.items()})
    return sd, enc


def _leafify(sd):
    params = {k: v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    for v in params.values():
        v.requires_grad_(True)
    return params


@pytest.mark.parametrize("arch,xr,size,b,s", [("resnet50", False, 64, 2, 3), ("resnet50", False, 160, 1, 2),
                                              ("resnet18", False, 64, 2, 3), ("resnet34", False, 64, 1, 2),
                                              ("resnext50_32x4d", True, 64, 3, 0), ("resnext50_32x4d", True, 350, 1, 0)])
def test_fe_eval_features(cuda, arch, xr, size, b, s):
    """Whole extractor, eval mode (BatchNorm running statistics): features within 1e-2 of the fp32 oracle, incl. the
    real 160x160 MRI and the odd 350x350 XR geometry (350 -> 175 -> 88 -> 44 -> 22 -> 11)."""
    sd, enc = _fe_pair(arch, cuda)
    enc.eval()
    if xr:
        x = _randn(b, 1, size, size, seed=5)
        imgs = x.expand(-1, 3, -1, -1)
        with torch.no_grad():
            tok = enc.encode_image(x)
    else:
        x = _randn(b, 1, size, size, s, seed=5)
        imgs = ko._slices_to_images(x)
        with torch.no_grad():
            tok = enc.encode_volume(x)
    with torch.no_grad():
        ref = ko.fe_forward(sd, "_fe", arch, imgs, False, True).flatten(1)
    got = tok.reshape(-1, tok.shape[-1])
    assert torch.isfinite(got).all()
    assert rel(got, ref) < LOGIT_TOL, rel(got, ref)


def test_fe_reference_call_convention(cuda):
    """SliceEncoder.forward takes what the reference feeds its nn.Sequential: (N, 3, H, W) with three identical
    channels, returns (N, C, 1, 1) with GAP and (N, C, h, w) without (_xrNmrMcP.py:47-59,218-220)."""
    sd, enc = _fe_pair("resnet18", cuda)
    enc.eval()
    x = _randn(4, 1, 64, 64, seed=6).expand(-1, 3, -1, -1)
    with torch.no_grad():
        got = enc(x)
        ref = ko.fe_forward(sd, "_fe", "resnet18", x, False, True)
    assert got.shape == ref.shape == (4, 512, 1, 1)
    assert rel(got, ref) < LOGIT_TOL
    sd2, enc2 = _fe_pair("resnet18", cuda, with_gap=False)
    enc2.eval()
    with torch.no_grad():
        got = enc2(x)
        ref = ko.fe_forward(sd2, "_fe", "resnet18", x, False, False)
    assert got.shape == ref.shape == (4, 512, 2, 2)
    assert rel(got, ref) < LOGIT_TOL


@pytest.mark.parametrize("arch,res_gain,b,s,size", [("resnet50", 0.1, 4, 4, 64), ("resnet50", 1.0, 4, 4, 64),
                                                    ("resnet18", 1.0, 2, 3, 64)])
def test_fe_train_at_bf16_floor(cuda, arch, res_gain, b, s, size):
    """Train mode (batch statistics, running-stat update, full backward) against the fp32 oracle. The bar is the
    bf16 precision floor measured in the same test with no CUDA-path code involved (fp32 oracle vs the oracle with
    every stored activation / GEMM operand rounded to bf16): CUDA-path error <= 1.5 x floor + 2e-3."""
    x = _randn(b, 1, size, size, s, seed=5)
    imgs = ko._slices_to_images(x)
    gy = None
    runs = {}
    for tag in ("fp32", "bf16emu"):
        sd, _ = _fe_pair(arch, cuda, res_gain=res_gain)
        params = _leafify(sd)
        f = ko.fe_forward(sd, "_fe", arch, imgs, True, True, None, emulate_16bit=(tag == "bf16emu")).flatten(1)
        if gy is None:
            gy = _randn(*f.shape, seed=7)
        (f * gy).sum().backward()
        runs[tag] = (f.detach(), {k: v.grad for k, v in params.items()}, sd)
    sd, enc = _fe_pair(arch, cuda, res_gain=res_gain)
    enc.train()
    tok = enc.encode_volume(x)
    got = tok.reshape(-1, tok.shape[-1])
    (got * gy).sum().backward()
    f0, g0, sd0 = runs["fp32"]
    f1, g1, _ = runs["bf16emu"]
    floor_f, mine_f = rel(f1, f0), rel(got, f0)
    assert mine_f <= 1.5 * floor_f + 2e-3, (mine_f, floor_f)
    mine = dict(enc.named_parameters())
    floor_g = torch.tensor([rel(g1[k], g0[k]) for k in g0])
    mine_g = torch.tensor([rel(mine[k[len("_fe."):]].grad, g0[k]) for k in g0])
    assert torch.isfinite(mine_g).all()
    assert float(mine_g.median()) <= 1.5 * float(floor_g.median()) + 2e-3, (float(mine_g.median()), float(floor_g.median()))
    assert float(mine_g.max()) <= 2.0 * float(floor_g.max()) + 2e-3, (float(mine_g.max()), float(floor_g.max()))
    # running statistics / num_batches_tracked are updated like nn.BatchNorm2d(momentum 0.1)
    msd = enc.state_dict()
    for k, v in sd0.items():
        if k.endswith(("running_mean", "running_var")):
            assert rel(msd[k[len("_fe."):]], v) < max(2e-2, 2 * floor_f), k
        if k.endswith("num_batches_tracked"):
            assert int(msd[k[len("_fe."):]]) == int(v) == 1


@pytest.mark.parametrize("kind", ["loader_range", "raw_0_255"])
def test_fe_input_ranges(cuda, kind):
    """Inputs in the ranges a real loader produces, not N(0, 1): (a) what the reference's transform chain delivers,
    uniform [0, 1] voxels -> (x - mean) / std with the DESS statistics (`_data_provider.py:297-334`);
    (b) what a loader that skipped PTToUnitRange / PTNormalize would deliver, 0 .. 255. The forward activations are fp16
    (max 65504); every tensor but the stem's raw convolution output sits behind a BatchNorm, and the stem sees
    |x| * |w| * 147 taps, far below the fp16 range for 8-bit data. Train mode (batch statistics): features finite and at the
    same 16-bit floor as with N(0, 1) inputs."""
    g = torch.Generator().manual_seed(21)
    u = torch.rand(4, 1, 64, 64, 4, generator=g)
    x = ((u - 0.257) / 0.235 if kind == "loader_range" else torch.floor(u * 256.0).clamp(max=255.0)).to(cuda)
    imgs = ko._slices_to_images(x)
    runs = {}
    with torch.no_grad():
        for tag in ("fp32", "emu"):
            sd, _ = _fe_pair("resnet50", cuda)
            runs[tag] = ko.fe_forward(sd, "_fe", "resnet50", imgs, True, True, None, emulate_16bit=(tag == "emu")).flatten(1)
    sd, enc = _fe_pair("resnet50", cuda)
    enc.train()
    tok = enc.encode_volume(x)
    got = tok.detach().reshape(-1, tok.shape[-1])
    assert torch.isfinite(got).all()
    floor_f, mine_f = rel(runs["emu"], runs["fp32"]), rel(got, runs["fp32"])
    assert mine_f <= 1.5 * floor_f + 2e-3, (kind, mine_f, floor_f)
    assert _lib.debug_flag() == 0


def test_fe_fp16_overflow_is_loud(cuda):
    """The stem's raw convolution output is stored in fp16: inputs of magnitude 1e7 (nothing a loader produces) exceed its
    range. Without a guard the features came back FINITE and wrong (0.49 relative; the ReLU kernels turn NaN into 0); the
    encoder therefore replaces them by NaN wherever max|x| * max_c sum|w_c| reaches the fp16 range (`_fe.py`), and leaves
    ordinary inputs bit for bit alone."""
    from oaprogressionmmf_b200.koamodels._fe import poison_if_out_of_fp16_range

    x = _randn(2, 1, 64, 64, 3, seed=9)
    sd, enc = _fe_pair("resnet50", cuda)
    enc.train()
    tok = enc.encode_volume(x * 1e7)
    assert bool(torch.isnan(tok).all())
    enc.eval()
    with torch.no_grad():
        good = enc.encode_volume(x)
        assert torch.isfinite(good).all()
        assert torch.equal(poison_if_out_of_fp16_range(x, enc[0].weight, good), good)
        assert bool(torch.isnan(enc.encode_volume(torch.full_like(x, float("nan")))).all())
    _lib.debug_flag()


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("arch,size,s", [("resnet50", 64, 3), ("resnext50_32x4d", 64, 0), ("resnet18", 64, 2)])
def test_fe_backward_in_stages_equals_one_call(cuda, arch, size, s, train):
    """koa_fe_backward_range over adjacent block ranges (how the data-parallel wrapper runs the backward pass, one
    gradient all-reduce per finished stage) computes what koa_fe_backward computes, for any split of the block list.
    With BatchNorm in eval mode the backward pass is well conditioned: equal up to the summation order of the split-K
    atomics (1e-5). In train mode on this tiny batch (2 x 2 pixels in layer4) the BatchNorm backward amplifies the
    summation-order noise of its statistics from layer to layer, and that noise depends on the launch timing: there the
    yardstick is the scatter between two identical runs, with a 2e-2 allowance."""
    from oaprogressionmmf_b200 import dataparallel

    lib = _lib.load()
    sd, enc = _fe_pair(arch, cuda)
    enc.train(train)
    x = _randn(2, 1, size, size, max(s, 1), seed=5) if s else _randn(3, 1, size, size, seed=5)
    tok = enc.encode_volume(x) if s else enc.encode_image(x)
    gy = _randn(*tok.shape, seed=7)
    fn = tok.grad_fn
    while fn is not None and not hasattr(fn, "ws"):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    ws, desc, table = fn.ws, fn.desc, fn.table
    params = enc._trainable()
    n_blocks = lib.koa_fe_num_blocks(C.byref(desc))
    assert n_blocks == sum(len(enc[li]) for li in range(4, 8))

    def run(ranges):
        grads, flat = _lib.zeros_like_flat(params)
        gt = _lib.ptr_table(grads)
        d = gy.reshape(-1, gy.shape[-1]).contiguous()
        for begin, end, stem in ranges:
            _lib.check(lib.koa_fe_backward_range(C.byref(desc), table, gt, ws.data_ptr(), d.data_ptr(), begin, end, stem,
                                                 _stream()), "range")
        torch.cuda.synchronize()
        return flat.clone()

    whole = run([(0, -1, 1)])
    # two runs of the same backward pass differ by the summation order of the statistics atomics, amplified through the
    # BatchNorm backward of ~50 layers on this tiny batch: that run-to-run scatter is the yardstick
    scatter = max(rel(run([(0, -1, 1)]), whole) for _ in range(2))
    stages = [(b, e, st) for b, e, st, *_ in enc._stage_bounds(params, *_lib.zeros_like_flat(params))]
    assert stages[0][1] == n_blocks and stages[-1][0] == 0 and stages[-1][2] == 1
    for ranges in (stages, [(bi, bi + 1, 1 if bi == 0 else 0) for bi in range(n_blocks - 1, -1, -1)]):
        got = run(ranges)
        assert rel(got, whole) < (max(3 * scatter, 2e-2) if train else 1e-5), (rel(got, whole), scatter)
    assert scatter < (3e-2 if train else 1e-5), scatter
    # the slices the wrapper hands to the all-reduce cover every gradient exactly once
    grads, flat = _lib.zeros_like_flat(params)
    covered = torch.zeros_like(flat, dtype=torch.int32)
    seen = []
    for _, _, _, lo, hi, ps in enc._stage_bounds(params, grads, flat):
        covered[lo:hi] += 1
        seen += ps
    assert int(covered.max()) == 1 and len(seen) == len(params) and len({id(p) for p in seen}) == len(params)
    for g in grads:
        off = (g.data_ptr() - flat.data_ptr()) // 4
        assert bool((covered[off:off + g.numel()] == 1).all())
    assert not dataparallel.active()
    assert _lib.debug_flag() == 0


def _ws_view(lib, desc, ws, what, index, dtype):
    off, nb = C.c_size_t(), C.c_size_t()
    _lib.check(lib.koa_fe_debug_offset(C.byref(desc), what, index, C.byref(off), C.byref(nb)), "koa_fe_debug_offset")
    return ws[off.value:off.value + nb.value].view(dtype)


def _set_bf16_copy(lib, desc, ws, what, index, values):
    """bf16 copies of the activations exist only with KOA_WGRAD_XCVT=0 (otherwise the weight-gradient kernels convert
    the fp16 activations themselves and the selector reports 0 bytes)."""
    v = _ws_view(lib, desc, ws, what, index, torch.bfloat16)
    if v.numel():
        v.copy_(values)


def _nhwc_f16(t):
    return t.detach().permute(0, 2, 3, 1).contiguous().half().reshape(-1)


@pytest.mark.parametrize("arch,xr,b,s,size", [("resnet50", False, 2, 3, 64), ("resnet50", False, 2, 8, 96),
                                              ("resnext50_32x4d", True, 3, 0, 64), ("resnet18", False, 2, 3, 64)])
def test_fe_backward_teacher_forced(cuda, arch, xr, b, s, size):
    """koa_fe_backward on a forward state forced to equal the bf16-emulating oracle's (the saved activations in the
    workspace are overwritten through koa_fe_debug_offset): every parameter gradient within a few bf16 roundings
    of autograd (median < 2e-2, max < 1e-1 relative L2 over 60-159 tensors; gradients are stored in bf16 between
    layers, ~150 roundings deep)."""
    lib = _lib.load()
    sd, enc = _fe_pair(arch, cuda)
    enc.train()
    if xr:
        x = _randn(b, 1, size, size, seed=5)
        imgs = x.expand(-1, 3, -1, -1)
        tok = enc.encode_image(x)
    else:
        x = _randn(b, 1, size, size, s, seed=5)
        imgs = ko._slices_to_images(x)
        tok = enc.encode_volume(x)
    fn = tok.grad_fn
    while fn is not None and not hasattr(fn, "ws"):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    ws, desc = fn.ws, fn.desc
    params = _leafify(sd)
    taps = {}
    ref = ko.fe_forward(sd, "_fe", arch, imgs, True, True, taps, emulate_16bit=True).flatten(1)
    ykeys = ["_fe.0.y"]
    plan = ko.fe_block_plan(arch)
    for blk in plan:
        p = f"_fe.{blk['layer']}.{blk['index']}"
        ykeys += [f"{p}.conv1.y", f"{p}.conv2.y"]
        if blk["kind"] == "bottleneck":
            ykeys.append(f"{p}.conv3.y")
        if blk["downsample"]:
            ykeys.append(f"{p}.downsample.0.y")
    for ui, key in enumerate(ykeys):
        y = taps[key].detach()
        yv = _ws_view(lib, desc, ws, 0, ui, torch.float16)
        if yv.numel():  # (the last convolution of a train-mode bottleneck stores no y: bn_gram.cu)
            yv.copy_(_nhwc_f16(y))
        coef = _ws_view(lib, desc, ws, 6, ui, torch.float32).view(7, -1)
        coef[2].copy_(y.mean(dim=(0, 2, 3)))
        coef[3].copy_(torch.rsqrt(y.var(dim=(0, 2, 3), unbiased=False) + 1e-5))
    _ws_view(lib, desc, ws, 4, 0, torch.float16).copy_(_nhwc_f16(taps["_fe.stem"]))
    for bi, blk in enumerate(plan):
        p = f"_fe.{blk['layer']}.{blk['index']}"
        # fp16 activations (gates / next-layer operands) and their bf16 copies (weight-gradient operands)
        _ws_view(lib, desc, ws, 1, bi, torch.float16).copy_(_nhwc_f16(taps[p]))
        _set_bf16_copy(lib, desc, ws, 8, bi, _nhwc_f16(taps[p]).bfloat16())
        _ws_view(lib, desc, ws, 2, bi, torch.float16).copy_(_nhwc_f16(taps[f"{p}.a1"]))
        _set_bf16_copy(lib, desc, ws, 9, bi, _nhwc_f16(taps[f"{p}.a1"]).bfloat16())
        if blk["kind"] == "bottleneck":
            _ws_view(lib, desc, ws, 3, bi, torch.float16).copy_(_nhwc_f16(taps[f"{p}.a2"]))
            _set_bf16_copy(lib, desc, ws, 10, bi, _nhwc_f16(taps[f"{p}.a2"]).bfloat16())
            sa2 = _ws_view(lib, desc, ws, 12, bi, torch.float32)
            if sa2.numel():  # y-free tail: its forward state is s = colsum(a2) and Q = W3 . (a2^T a2) / N
                a2 = _nhwc_f16(taps[f"{p}.a2"]).double().view(-1, sa2.numel())
                w3 = sd[f"{p}.conv3.weight"].detach().half().double().flatten(1)
                sa2.copy_(a2.sum(0))
                _ws_view(lib, desc, ws, 13, bi, torch.float32).copy_((w3 @ (a2.t() @ a2 / a2.shape[0])).flatten())
    a0 = _ws_view(lib, desc, ws, 4, 0, torch.float16)
    p0 = _ws_view(lib, desc, ws, 5, 0, torch.float16)
    idx0 = _ws_view(lib, desc, ws, 7, 0, torch.uint8)
    hs = (size + 6 - 7) // 2 + 1
    _lib.check(lib.koa_maxpool_fwd(a0.data_ptr(), p0.data_ptr(), idx0.data_ptr(), imgs.shape[0], hs, hs, 64, _stream()),
               "maxpool")
    assert torch.equal(p0.float(), _nhwc_f16(taps["_fe.pool"]).float())
    _set_bf16_copy(lib, desc, ws, 11, 0, p0.bfloat16())
    gy = _randn(*ref.shape, seed=7)
    (tok.reshape(-1, tok.shape[-1]) * gy).sum().backward()
    (ref * gy).sum().backward()
    mine = dict(enc.named_parameters())
    errs = torch.tensor([rel(mine[k[len("_fe."):]].grad, v.grad) for k, v in params.items()])
    assert torch.isfinite(errs).all()
    assert float(errs.median()) < 2e-2 and float(errs.max()) < 1e-1, (float(errs.median()), float(errs.max()))


@pytest.mark.parametrize("b,n_p,depth,with_cls,head", [(2, 5, 2, True, True), (3, 7, 1, False, False), (2, 64, 1, False, False),
                                                       (1, 123, 1, True, True), (4, 25, 1, False, True)])
def test_feat_forward_backward(cuda, b, n_p, depth, with_cls, head):
    """FeaT (embedding, CLS/pos, pre-norm blocks, head; _core_trf.py:74-205) vs the fp32 oracle: token states, logits,
    input gradient and every parameter gradient; dead heads keep grad None like the reference."""
    from oaprogressionmmf_b200.koamodels import FeaT

    dim, heads = 2048, 8
    spec = ko.feat_param_spec("_agg", n_p, dim, depth, dim, 2, with_cls)
    sd = ko.make_state_dict(spec, 21, pos_scale=0.5, device=cuda)
    mod = FeaT(n_p, dim, dim, depth, heads, dim, 2, with_cls=with_cls).to(cuda)
    mod.load_state_dict({k[len("_agg."):]: v.clone() for k, v in sd.items()})
    mod.train()
    tok = _randn(b, n_p, dim, seed=6, scale=0.7).requires_grad_(True)
    tok_ref = tok.detach().clone().requires_grad_(True)
    out, states, attns = mod.run(tok, compute_head=head)
    for v in sd.values():
        v.requires_grad_(True)
    out_ref, states_ref = ko.feat_forward(sd, "_agg", tok_ref, depth, heads, 0.0, 0.0, True)
    assert out.shape == out_ref.shape and states.shape == states_ref.shape and len(attns) == depth
    assert rel(states, states_ref) < 5e-3
    gs, go = _randn(*states.shape, seed=8), _randn(*out.shape, seed=9)
    loss = (states * gs).sum() + ((out * go).sum() if head else 0)
    loss_ref = (states_ref * gs).sum() + ((out_ref * go).sum() if head else 0)
    if head:
        assert rel(out, out_ref) < LOGIT_TOL
    loss.backward()
    loss_ref.backward()
    assert rel(tok.grad, tok_ref.grad) < 1e-2
    mine = dict(mod.named_parameters())
    for k, v in sd.items():
        gm = mine[k[len("_agg."):]].grad
        if v.grad is None:
            assert gm is None, k
            continue
        assert gm is not None, k
        assert rel(gm, v.grad) < 1e-2, (k, rel(gm, v.grad))


def test_exported_transformer_classes_run_on_their_own(cuda):
    """``koafusion.models`` exports ``Transformer``, ``Attention`` and ``FeedForward`` next to ``FeaT``
    (``models/__init__.py:1``): used on their own they compute what the reference classes compute
    (``_core_trf.py:141-205``), forward and backward, operator by operator through the C ABI (``koamodels/_ops.py``);
    ``state_dict`` keys are the reference's, so the comparison loads one into the other."""
    from oaprogressionmmf_b200.koamodels import Attention, FeedForward, Transformer

    dim, heads, depth, b, n = 512, 8, 2, 3, 37
    torch.manual_seed(5)
    mine = Transformer(dim, depth, heads, dim, 0.0).to(cuda)
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}

    def ref_transformer(x):
        attns = []
        for d in range(depth):
            p = lambda k: sd[k].requires_grad_(True)  # noqa: E731
            o = F.layer_norm(x, (dim,), p(f"prenorm_0_{d}.weight"), p(f"prenorm_0_{d}.bias"))
            qkv = F.linear(o, p(f"attn_{d}.to_qkv.weight"))
            q, k, v = qkv.view(b, n, 3, heads, dim // heads).permute(2, 0, 3, 1, 4)
            attn = torch.softmax(q @ k.transpose(-1, -2) * dim ** -0.5, dim=-1)
            o = (attn @ v).permute(0, 2, 1, 3).reshape(b, n, dim)
            x = F.linear(o, p(f"attn_{d}.to_out.0.weight"), p(f"attn_{d}.to_out.0.bias")) + x
            attns.append(attn)
            o = F.layer_norm(x, (dim,), p(f"prenorm_1_{d}.weight"), p(f"prenorm_1_{d}.bias"))
            o = F.linear(F.gelu(F.linear(o, p(f"ff_{d}.net.0.weight"), p(f"ff_{d}.net.0.bias"))), p(f"ff_{d}.net.3.weight"),
                         p(f"ff_{d}.net.3.bias"))
            x = o + x
        return x, attns

    x = _randn(b, n, dim, seed=3).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    out, attns = mine(x)
    ref, rattns = ref_transformer(xr)
    assert len(attns) == depth and attns[0].shape == (b, heads, n, n)
    assert rel(out, ref) < 2e-3 and rel(attns[1], rattns[1]) < 2e-3, (rel(out, ref), rel(attns[1], rattns[1]))
    g = _randn(*out.shape, seed=4)
    (out * g).sum().backward()
    (ref * g).sum().backward()
    assert rel(x.grad, xr.grad) < 1e-2, rel(x.grad, xr.grad)
    for k, p in mine.named_parameters():
        assert rel(p.grad, sd[k].grad) < 1.5e-2, (k, rel(p.grad, sd[k].grad))
    # the two smaller building blocks, and the errors of the interface
    ff = FeedForward(dim, 2 * dim).to(cuda)
    y = ff(x.detach())
    assert rel(y, F.linear(F.gelu(F.linear(x.detach(), ff.net[0].weight, ff.net[0].bias)), ff.net[3].weight, ff.net[3].bias)) < 2e-3
    at = Attention(dim, heads).to(cuda)
    o, a = at(x.detach())
    assert o.shape == (b, n, dim) and a.shape == (b, heads, n, n) and abs(float(a.sum(-1).mean()) - 1) < 1e-5
    with pytest.raises(ValueError):
        at(x.detach(), mask=torch.ones(b, n - 1, dtype=torch.bool, device=cuda))
    with pytest.raises(_lib.KoaError):
        ff(x.detach().cpu())


def test_resnet_classifier_forward(cuda):
    """``dict_fes[arch]()`` is a complete classifier in the reference (``_torchvision.py:227-242``): here its ``forward``
    runs the CUDA extractor + the fp32 ``fc`` kernel and matches the oracle's extractor followed by ``F.linear``."""
    from oaprogressionmmf_b200.koamodels import dict_fes

    sd, enc = _fe_pair("resnet18", cuda)
    net = dict_fes["resnet18"](pretrained=False).to(cuda).eval()
    # the oracle's keys are those of nn.Sequential(*children): position -> attribute name of the full network
    remap = {"0": "conv1", "1": "bn1", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}
    net.load_state_dict({**net.state_dict(), **{remap[k[len("_fe."):].split(".")[0]] + k[len("_fe."):][1:]: v for k, v in sd.items()}})
    x = _randn(3, 1, 64, 64, seed=6).expand(-1, 3, -1, -1)
    with torch.no_grad():
        got = net(x)
        feat = ko.fe_forward(sd, "_fe", "resnet18", x, False, True).flatten(1)
        ref = F.linear(feat, net.fc.weight, net.fc.bias)
    assert got.shape == (3, 1000) and rel(got, ref) < LOGIT_TOL


def test_feat_dropout_matches_oracle_with_the_same_masks(cuda):
    """nn.Dropout inside FeaT (emb_dropout after the positional embedding, mlp_dropout after to_out / GELU / ff out /
    head GELU; _core_trf.py:105,127,146-149,164). The engine draws counter-based Philox masks; koa_dropout_mask
    returns the very same masks, the oracle applies them instead of torch's random stream, and forward + backward
    must then agree like in the dropout-free test. Also: the mask statistics, reproducibility under a seed, and
    eval mode ignoring dropout."""
    from oaprogressionmmf_b200.koamodels import FeaT

    lib = _lib.load()
    b, n_p, dim, depth, heads, p_emb, p_mlp = 3, 9, 2048, 2, 8, 0.1, 0.2
    spec = ko.feat_param_spec("_agg", n_p, dim, depth, dim, 2, True)
    sd = ko.make_state_dict(spec, 21, pos_scale=0.5, device=cuda)
    mod = FeaT(n_p, dim, dim, depth, heads, dim, 2, emb_dropout=p_emb, mlp_dropout=p_mlp).to(cuda)
    mod.load_state_dict({k[len("_agg."):]: v.clone() for k, v in sd.items()})
    mod.train()
    tok = _randn(b, n_p, dim, seed=6, scale=0.7).requires_grad_(True)
    tok_ref = tok.detach().clone().requires_grad_(True)
    torch.manual_seed(123)
    out, states, _ = mod.run(tok, compute_head=True)
    seed = mod.last_dropout_seed
    assert seed != 0
    n = n_p + 1

    def mask(site, rows, cols, p):
        m = torch.empty(rows, cols, device=cuda)
        _lib.check(lib.koa_dropout_mask(seed, site, rows, cols, p, m.data_ptr(), _stream()), "mask")
        return m

    masks = {"emb": mask(0xE000, b * n, dim, p_emb), "head": mask(0xF000, b, dim, p_mlp)}
    for d in range(depth):
        masks[f"attn_out_{d}"] = mask(4 * d + 0, b * n, dim, p_mlp)
        masks[f"ff_act_{d}"] = mask(4 * d + 1, b * n, dim, p_mlp)
        masks[f"ff_out_{d}"] = mask(4 * d + 2, b * n, dim, p_mlp)
    for k, m in masks.items():  # values are 0 or 1/(1-p), drop rate within 4 sigma
        p = p_emb if k == "emb" else p_mlp
        vals = torch.unique(m)
        assert len(vals) == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1 / (1 - p)) < 1e-6, k
        rate = float((m == 0).float().mean())
        assert abs(rate - p) < 4 * (p * (1 - p) / m.numel()) ** 0.5 + 1e-4, (k, rate)
    assert not torch.equal(masks["ff_out_0"], masks["ff_out_1"]) and not torch.equal(masks["ff_out_0"], masks["attn_out_0"])
    for v in sd.values():
        v.requires_grad_(True)
    out_ref, states_ref = ko.feat_forward(sd, "_agg", tok_ref, depth, heads, p_emb, p_mlp, True, masks)
    assert rel(states, states_ref) < 5e-3 and rel(out, out_ref) < LOGIT_TOL
    gs, go = _randn(*states.shape, seed=8), _randn(*out.shape, seed=9)
    ((states * gs).sum() + (out * go).sum()).backward()
    ((states_ref * gs).sum() + (out_ref * go).sum()).backward()
    assert rel(tok.grad, tok_ref.grad) < 1e-2
    mine = dict(mod.named_parameters())
    for k, v in sd.items():
        assert rel(mine[k[len("_agg."):]].grad, v.grad) < 1e-2, (k, rel(mine[k[len("_agg."):]].grad, v.grad))
    # same torch seed -> same masks -> same output; different seed -> different output
    with torch.no_grad():
        torch.manual_seed(123)
        again = mod.run(tok.detach(), compute_head=True)[1]
        assert torch.equal(again, states.detach())
        torch.manual_seed(124)
        other = mod.run(tok.detach(), compute_head=True)[1]
        assert rel(other, states.detach()) > 1e-2
        # eval mode ignores dropout: equals the dropout-free oracle
        mod.eval()
        ev = mod.run(tok.detach(), compute_head=True)[1]
        ev_ref = ko.feat_forward(sd, "_agg", tok_ref.detach(), depth, heads, p_emb, p_mlp, False)[1]
        assert rel(ev, ev_ref) < 5e-3


# =====================================================================================================
# 3. models, on the fixtures generated from the unmodified reference
# =====================================================================================================
GOLDEN = sorted(f[:-5] for f in os.listdir(os.path.join(os.path.dirname(__file__), "golden")) if f.endswith(".json"))


def _golden_case(case, golden_dir, dev, pos_scale=1.0):
    from oaprogressionmmf_b200.koamodels import dict_models

    with open(os.path.join(golden_dir, case + ".json")) as f:
        gold = json.load(f)
    name = gold["model"]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in gold["config_kwargs"].items()}
    cfg = ko.make_config(name, **kw)
    spec = ko.model_param_spec(name, cfg)
    inputs, target = ko.make_inputs(name, cfg, gold["batch"], gold["seed_inputs"], device=dev)
    model = dict_models[name](to_attr(cfg), None).to(dev)
    model.load_state_dict(ko.make_state_dict(spec, gold["seed_weights"], pos_scale=pos_scale, device=dev), strict=True)
    return gold, cfg, model, inputs, target


@pytest.mark.parametrize("case", GOLDEN)
def test_model_eval_logits_match_reference(cuda, golden_dir, case):
    """Eval-mode logits of every model class against the logits the *unmodified reference* produced for the same
    weights and inputs: within 1e-2 relative (bf16) and identical class predictions."""
    gold, cfg, model, inputs, _ = _golden_case(case, golden_dir, cuda)
    model.eval()
    with torch.no_grad():
        out = model(*inputs)
    assert isinstance(out, dict) and list(out) == ["main"]
    lg = out["main"]
    ref = torch.tensor(gold["eval_logits"], device=cuda)
    assert lg.shape == ref.shape
    # The tiny fixtures have 2-3 knees x 2 logits of magnitude ~0.05 computed from 32x32 / 64x64 inputs: the relative
    # L2 error of those four numbers under bf16 storage scatters around 0.3e-2 .. 1.1e-2 (the bf16-emulating oracle,
    # which involves no CUDA-path code, lands in the same range: `floor`). The bar here is therefore 2e-2; the 1e-2
    # bar of BASELINE.json is asserted where it is defined, at full size (test_full_size_logits_match_reference).
    spec = ko.model_param_spec(gold["model"], cfg)
    sd = ko.make_state_dict(spec, gold["seed_weights"], device=cuda)
    with torch.no_grad():
        emu = ko.model_forward(gold["model"], cfg, sd, inputs, training=False, emulate_16bit=True)
    floor = rel(emu, ref)
    tol = TINY_LOGIT_TOL
    assert rel(lg, ref) < tol, (rel(lg, ref), floor)
    assert bool((lg.argmax(1) == ref.argmax(1)).all())


@pytest.mark.parametrize("case", GOLDEN)
def test_model_train_step_matches_reference_structure(cuda, golden_dir, case):
    """One train-mode step on the golden case: loss finite and close to the reference's, a gradient for exactly
    the parameters the reference has one for (None on the dead per-sequence heads), every gradient finite and of
    the reference's magnitude. Tiny train-mode batches (6 images of 32x32) sit on the chaotic BatchNorm floor described
    in the module docstring (the loss itself moves by 2-3 % from run to run with the order of the statistics atomics,
    and a whole extractor's gradient scale with it), so this is a structural check with loose magnitudes (loss within
    40 %, gradient norms within a factor 6, median within a factor 3); numerical parity of the train step is asserted by the engine tests above
    and, against the reference itself, at full size below."""
    from oaprogressionmmf_b200.losses import FocalLoss

    # the reference's train step was recorded on the sensitised weights (pos_embedding / cls_token x 0.02, so that
    # the logits depend on the image features; oracle/make_golden.py)
    gold, cfg, model, inputs, target = _golden_case(case, golden_dir, cuda, pos_scale=0.02)
    model.train()
    lg = model(*inputs)["main"]
    loss = FocalLoss(gamma=2)(lg, target)
    loss.backward()
    assert abs(float(loss) - gold["train_loss"]) < 0.4 * max(gold["train_loss"], 0.1)
    ratios = []
    for k, p in model.named_parameters():
        gref = gold["grads"][k]
        if gref is None:
            assert p.grad is None, f"{k}: reference leaves grad None"
            continue
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
        ratios.append(float(p.grad.norm()) / (gref["norm"] + 1e-30))
    r = torch.tensor(ratios)
    assert float(r.min()) > 1 / 6 and float(r.max()) < 6, (float(r.min()), float(r.max()))
    assert 1 / 3 < float(r.median()) < 3


GOLDEN_FULL = sorted(f[:-5] for f in os.listdir(os.path.join(os.path.dirname(__file__), "golden_full")) if f.endswith(".json"))
def logit_err(got, ref):
    """BASELINE.json's metric, literally: ||got - ref|| / ||ref|| over the logit tensor, no absolute floor."""
    got, ref = got.detach().double(), ref.detach().double()
    return float((got - ref).norm() / ref.norm())


class _Taps:
    """Records what the fixtures' ``taps`` hold for the reference: the output of every extractor (``_fe*``) and the token
    states of every transformer (``_agg*``) of one forward pass, keyed by the child-module name."""

    def __init__(self, model):
        from oaprogressionmmf_b200.koamodels import _feat, _fe, _small

        self.out, self.names, self.saved = {}, {id(m): k for k, m in model.named_children()}, []
        t = self

        def wrap(cls, attr, pick):
            orig = getattr(cls, attr)
            self.saved.append((cls, attr, orig))

            def f(mod, *a, **k):
                r = orig(mod, *a, **k)
                if id(mod) in t.names:
                    t.out[t.names[id(mod)]] = pick(r).detach()
                return r

            setattr(cls, attr, f)

        wrap(_fe.SliceEncoder, "encode_volume", lambda r: r)
        wrap(_fe.SliceEncoder, "encode_image", lambda r: r)
        wrap(_feat.FeaT, "run", lambda r: r[1])
        wrap(_small.FeatC1, "forward", lambda r: r)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        for cls, attr, orig in self.saved:
            setattr(cls, attr, orig)


def _probe(numel, device):
    return torch.randn(numel, generator=torch.Generator().manual_seed(numel % 9973 + 17)).to(device)


@pytest.mark.parametrize("case", GOLDEN_FULL)
def test_full_size_logits_match_reference(cuda, case):
    """BASELINE.json's sizes (XR 350x350 ResNeXt-50, DESS 160x160x64 / TSE x32 / T2 x25 ResNet-50s, D 2048, depth 4)
    against what the UNMODIFIED reference produced for the same seeded weights and inputs
    (oracle/make_golden_fullsize.py), all six reference classes and the two 3-MRI extensions.

    * eval logits: ||delta|| / ||ref|| < 1e-2, the north-star tolerance as written (no floor), also with the sensitised
      weights (pos_embedding / cls_token x 0.02, where the logits depend on the input), identical class predictions;
    * the intermediates the logits hide: every extractor output and every transformer's states within 1e-2 (norm and
      the five recorded values);
    * one train step: loss within 1 %, a gradient exactly where the reference has one, and per tensor: the norm within
      [0.6, 2.0] x reference (median within 5 %) and the DIRECTION: the projection of the gradient on a seeded probe
      vector, which estimates the relative L2 error of the tensor (a sign or permutation error inside a tensor keeps the
      norm and destroys the projection)."""
    from oaprogressionmmf_b200.losses import FocalLoss

    gold_dir = os.path.join(os.path.dirname(__file__), "golden_full")
    for tag, ps in (("eval_logits", 1.0), ("eval_logits_sensitised", 0.02)):
        gold, cfg, model, inputs, target = _golden_case(case, gold_dir, cuda, pos_scale=ps)
        model.eval()
        with torch.no_grad(), _Taps(model) as taps:
            lg = model(*inputs)["main"]
        ref = torch.tensor(gold[tag], device=cuda)
        assert logit_err(lg, ref) < LOGIT_TOL, (tag, logit_err(lg, ref))
        assert bool((lg.argmax(1) == ref.argmax(1)).all()), tag
        if ps == 1.0 and "taps" in gold:
            assert set(gold["taps"]) == set(taps.out), (sorted(gold["taps"]), sorted(taps.out))
            for k, g in gold["taps"].items():
                t = taps.out[k].flatten().double()
                assert t.numel() == int(torch.tensor(g["shape"]).prod()), k
                n = t.numel()
                idx = [0, n // 4, n // 2, (3 * n) // 4, n - 1]
                rms = g["norm"] / n ** 0.5
                assert abs(float(t.norm()) - g["norm"]) < 1e-2 * g["norm"], (k, float(t.norm()), g["norm"])
                for i, want in zip(idx, g["samples"]):
                    assert abs(float(t[i]) - want) < 1e-2 * abs(want) + 2e-2 * rms, (k, i, float(t[i]), want)
    model.train()  # the reference's train step ran on the sensitised weights
    lg = model(*inputs)["main"]
    loss = FocalLoss(gamma=2)(lg, target)
    loss.backward()
    assert abs(float(loss) - gold["train_loss"]) < 0.01 * gold["train_loss"], (float(loss), gold["train_loss"])
    assert logit_err(lg, torch.tensor(gold["train_logits"], device=cuda)) < 3e-2
    ratios, proj = [], {}
    for k, p in model.named_parameters():
        gref = gold["grads"][k]
        if gref is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
        ratios.append(float(p.grad.norm()) / (gref["norm"] + 1e-30))
        if "probe" in gref and gref["norm"] > 0:
            got = float((p.grad.flatten().double() * _probe(p.grad.numel(), cuda).double()).sum())
            proj[k] = (got - gref["probe"]) / gref["norm"]   # ~ N(0, 1) x relative L2 error of this tensor
    r = torch.tensor(ratios)
    assert float(r.min()) > 0.6 and float(r.max()) < 2.0, (float(r.min()), float(r.max()))
    assert abs(float(r.median()) - 1) < 0.05, float(r.median())
    # DIRECTION of the gradients, tensor by tensor. At random initialisation a train-mode BatchNorm ResNet is chaotic in
    # its gradients: the fp32 oracle on this GPU and the fp32 reference on the build container's CPU already disagree by
    # ~70 % (relative L2) on the early convolutions of the 2-knee fixtures, and so does ANY 16-bit arithmetic. So the bar
    # is (a) the floor: per tensor, the CUDA path is as close to the fp32 oracle (run here, same device, TF32 off) as the
    # oracle with 16-bit storage emulated (no CUDA-path code involved) is; (b) wherever that floor is low (the transformer
    # tensors: < 5e-2), the CUDA path itself is within 5e-2 (or 1.5 x the floor), the whole tensor compared element by
    # element; (c) the recorded projections of the REFERENCE's gradients on seeded probe vectors agree wherever neither the
    # machine nor 16-bit storage blurs them (fp32 oracle here and the floor both within 2e-2).
    spec = ko.model_param_spec(gold["model"], cfg)
    runs = {}
    for tag, emu in (("f32", False), ("emu", True)):
        sd = ko.make_state_dict(spec, gold["seed_weights"], pos_scale=0.02, device=cuda)
        runs[tag] = ko.train_step(gold["model"], cfg, sd, inputs, target, emulate_16bit=emu)[2]
    mine_e, floor_e, quiet_bad, probe_bad = [], [], [], []
    for k, p in model.named_parameters():
        if gold["grads"][k] is None:
            continue
        g32, gem = runs["f32"][k], runs["emu"][k]
        if float(g32.norm()) == 0:
            continue
        em, ef = rel(p.grad, g32), rel(gem, g32)
        mine_e.append(em)
        floor_e.append(ef)
        if ef < 5e-2 and em >= max(5e-2, 1.5 * ef):
            quiet_bad.append((k, em, ef))
        gref = gold["grads"][k]
        if k in proj:  # the reference's own number, wherever neither the machine (o32) nor 16-bit storage (ef) blurs it
            o32 = abs(float((g32.flatten().double() * _probe(g32.numel(), cuda).double()).sum()) - gref["probe"]) / gref["norm"]
            if o32 < 2e-2 and ef < 2e-2 and abs(proj[k]) > 8e-2:
                probe_bad.append((k, abs(proj[k]), o32, ef))
    mine_e, floor_e = torch.tensor(mine_e), torch.tensor(floor_e)
    assert float(mine_e.median()) <= 1.25 * float(floor_e.median()) + 1e-2, (float(mine_e.median()), float(floor_e.median()))
    assert float(mine_e.mean()) <= 1.25 * float(floor_e.mean()) + 1e-2, (float(mine_e.mean()), float(floor_e.mean()))
    if gold["model"] != "XR1Cnn":
        assert int((floor_e < 5e-2).sum()) >= 40, "the transformer tensors must be in the low-floor set"
    assert not quiet_bad, quiet_bad[:5]
    assert not probe_bad, probe_bad[:5]


@pytest.mark.parametrize("name", ["XR1MR2C1CnnTrf", "XR1MR3C1CnnTrf"])
def test_trained_weights_logits_match_oracle(cuda, name):
    """at random initialisation the logits barely depend on the input, so a parity claim on them says little.
    Here the full-size model (the reference's real one and the bench workload) is first TRAINED for 10 Adam steps (lr 1e-3)
    on four knees with the fp32 oracle on this GPU (the oracle is pinned to the unmodified reference by the CPU fixtures;
    a training trajectory cannot be replayed bit for bit across machines, so the trained weights are made here, not
    shipped), then evaluated on two unseen knees: logits O(0.1 .. 1) that differ between knees, CUDA path within 1e-2 of
    the oracle (literal ||delta|| / ||ref||) and the same class predictions."""
    from oaprogressionmmf_b200.koamodels import dict_models

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = ko.make_config(name)
    spec = ko.model_param_spec(name, cfg)
    sd = ko.make_state_dict(spec, 4242, pos_scale=0.02, device=cuda)
    params = [v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    for v in params:
        v.requires_grad_(True)
    opt = torch.optim.Adam(params, lr=1e-3)
    inputs, _ = ko.make_inputs(name, cfg, 4, 99, device=cuda)
    target = torch.tensor([0, 1, 1, 0], device=cuda)
    for _ in range(10):
        opt.zero_grad(set_to_none=True)
        loss = ko.focal_loss(ko.model_forward(name, cfg, sd, inputs, training=True), target)
        loss.backward()
        opt.step()
    sd = {k: v.detach() for k, v in sd.items()}
    test_in, _ = ko.make_inputs(name, cfg, 2, 100, device=cuda)
    with torch.no_grad():
        ref = ko.model_forward(name, cfg, sd, test_in, training=False)
    model = dict_models[name](to_attr(cfg), None).to(cuda)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        got = model(*test_in)["main"]
    assert float(ref.abs().max()) > 0.05, ref
    assert logit_err(got, ref) < LOGIT_TOL, (logit_err(got, ref), got, ref)
    assert bool((got.argmax(1) == ref.argmax(1)).all())


def test_output_type_main_returns_bare_tensor(cuda, golden_dir):
    gold, cfg, model, inputs, _ = _golden_case("XR1Cnn_r18", golden_dir, cuda)
    model.config["output_type"] = "main"
    model.eval()
    with torch.no_grad():
        out = model(*inputs)
    assert torch.is_tensor(out) and out.shape == (gold["batch"], 2)
    model.config["output_type"] = "nope"
    with pytest.raises(ValueError):
        model(*inputs)


def test_full_size_model_step_properties(cuda):
    """BASELINE.json's full sizes (XR 350^2 + DESS 160^2x64 + T2 160^2x25 + clinical, B = 2), where the oracle is too
    slow to be the checker: size-independent properties instead. (1) eval logits are invariant to the order of the
    knees in the batch (no cross-knee leakage through the slice batch); (2) a train step gives finite gradients for
    every live parameter and None on the dead heads; (3) the backward pass is linear in the loss scale: backward of
    2*loss doubles every gradient (a power-of-two scale survives the bf16 storage exactly; split-K atomics reorder
    fp32 sums, hence 1e-3 instead of bit equality)."""
    from oaprogressionmmf_b200.koamodels import dict_models
    from oaprogressionmmf_b200.losses import FocalLoss

    name = "XR1MR2C1CnnTrf"
    cfg = ko.make_config(name)
    spec = ko.model_param_spec(name, cfg)
    model = dict_models[name](to_attr(cfg), None).to(cuda)
    model.load_state_dict(ko.make_state_dict(spec, 778, device=cuda))
    inputs, target = ko.make_inputs(name, cfg, 2, 779, device=cuda)
    model.eval()
    with torch.no_grad():
        a = model(*inputs)["main"]
        b = model(*[t.flip(0) for t in inputs])["main"].flip(0)
    assert torch.isfinite(a).all() and rel(b, a) < 1e-6
    # backward linearity at full size. Eval-mode BatchNorm keeps the two forward passes bit-identical (train-mode
    # statistics are summed with atomics whose order varies run to run, and a 2-image batch amplifies that).
    grads = []
    for scale in (1.0, 2.0):
        model.zero_grad(set_to_none=True)
        (FocalLoss(gamma=2)(model(*inputs)["main"], target) * scale).backward()
        grads.append({k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()})
    for k, g in grads[0].items():
        dead = ("_agg_1.mlp_head0." in k) or ("_agg_2.mlp_head0." in k)
        assert (g is None) == dead, k
        if g is not None:
            assert torch.isfinite(g).all(), k
            assert rel(grads[1][k], 2 * g) < 1e-3, k
    # train mode: finite loss and gradients, BatchNorm buffers updated
    model.train()
    model.zero_grad(set_to_none=True)
    loss = FocalLoss(gamma=2)(model(*inputs)["main"], target)
    loss.backward()
    assert torch.isfinite(loss)
    for k, p in model.named_parameters():
        dead = ("_agg_1.mlp_head0." in k) or ("_agg_2.mlp_head0." in k)
        assert (p.grad is None) == dead, k
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), k
    assert int(model._fe1[1].num_batches_tracked) == 1


def test_branch_streams_do_not_change_the_results(cuda, golden_dir):
    """The fusion models run their modality branches on concurrent CUDA streams (koamodels/_models.py::_run_branches);
    the reference runs them one after the other (_xrNmrMcP.py:209-236). Eval logits are bit-identical either way, a
    train step gives the same loss and gradients up to the run-to-run noise of the statistics atomics."""
    from oaprogressionmmf_b200.koamodels import set_branch_streams
    from oaprogressionmmf_b200.losses import FocalLoss

    case = [c for c in GOLDEN if c.startswith("XR1MR3C1CnnTrf")][0]
    gold, cfg, model, inputs, target = _golden_case(case, golden_dir, cuda, pos_scale=0.02)
    try:
        res = {}
        for on in (False, True):
            set_branch_streams(on)
            model.eval()
            with torch.no_grad():
                lg = model(*inputs)["main"].clone()
            res[on] = lg
        assert torch.equal(res[False], res[True])
        # train mode on concurrent streams: finite, gradients everywhere the serial run has them, several repeats
        # (a missing cross-stream dependency shows up as NaN / garbage, not as a small deviation)
        set_branch_streams(True)
        model.train()
        for _ in range(3):
            model.zero_grad(set_to_none=True)
            loss = FocalLoss(gamma=2)(model(*inputs)["main"], target)
            loss.backward()
            torch.cuda.synchronize()
            assert bool(torch.isfinite(loss))
            assert abs(float(loss) - gold["train_loss"]) < 0.4 * max(gold["train_loss"], 0.1)
            for k, p in model.named_parameters():
                if gold["grads"][k] is not None:
                    assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    finally:
        set_branch_streams(None)
    assert _lib.debug_flag() == 0


def test_slice_encoder_rejects_a_colour_image(cuda):
    """The stem folds the three identical channels the reference feeds (`repeat(x, k=3)`, _xrNmrMcP.py:211-213) into one:
    an input whose channels differ is an error, not a silent use of channel 0."""
    sd, enc = _fe_pair("resnet18", cuda)
    enc.eval()
    grey = _randn(2, 1, 64, 64, seed=6)
    with torch.no_grad():
        a = enc(grey.expand(-1, 3, -1, -1))
        b = enc(grey.repeat(1, 3, 1, 1))
        assert torch.equal(a, b)
        enc2 = _fe_pair("resnet18", cuda)[1].eval()
        with pytest.raises(ValueError):
            enc2(_randn(2, 3, 64, 64, seed=7))
