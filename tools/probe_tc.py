"""GPU probe for the tcgen05 GEMM / implicit-GEMM kernels (run under gpurun).

Each group runs in its own subprocess with a timeout so a faulting kernel cannot take the other
groups down with it. Prints one line per case plus a coarse error map on mismatch.

    python tools/probe_tc.py            # all groups
    python tools/probe_tc.py --group gemm
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _imports():
    import torch

    from oaprogressionmmf_b200 import _lib

    return torch, _lib


def rel_err(torch, got, ref):
    got = got.float()
    ref = ref.float()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item(), (got - ref).abs().max().item()


def err_map(torch, got, ref, br=32, bc=32, max_rows=12, max_cols=16):
    got = got.float()
    ref = ref.float()
    m, n = got.shape[0], got.shape[-1]
    got = got.reshape(m, -1)
    ref = ref.reshape(m, -1)
    n = got.shape[1]
    lines = []
    for r0 in range(0, min(m, br * max_rows), br):
        row = []
        for c0 in range(0, min(n, bc * max_cols), bc):
            g = got[r0 : r0 + br, c0 : c0 + bc]
            rf = ref[r0 : r0 + br, c0 : c0 + bc]
            e = (g - rf).norm() / (rf.norm() + 1e-9)
            row.append("." if e < 2e-2 else ("x" if e < 0.9 else "X"))
        lines.append("".join(row))
    return "\n".join(lines)


def report(name, torch, _lib, got, ref, tol=1e-2):
    torch.cuda.synchronize()
    flag = _lib.debug_flag()
    rel, mx = rel_err(torch, got, ref)
    ok = rel < tol and flag == 0 and bool(torch.isfinite(got.float()).all())
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel={rel:.3e} maxabs={mx:.3e} flag=0x{flag:x}", flush=True)
    if not ok:
        print(err_map(torch, got, ref), flush=True)
    return ok


def group_gemm():
    torch, _lib = _imports()
    lib = _lib.load()
    st = _lib.current_stream()
    ok = True
    g = torch.Generator(device="cuda").manual_seed(1)
    for (m, n, k) in [(128, 128, 64), (128, 64, 64), (256, 128, 128), (1000, 192, 320), (777, 256, 2048),
                      (4096, 64, 64), (300, 2048, 512)]:
        a = torch.randn(m, k, device="cuda", generator=g).bfloat16()
        b = torch.randn(n, k, device="cuda", generator=g).bfloat16()
        ref = a.float() @ b.float().t()
        out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16)
        ep = _lib.Epilogue(out=out.data_ptr(), ldo=n)
        _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), "gemm")
        ok &= report(f"gemm bf16-out m{m} n{n} k{k}", torch, _lib, out, ref)
    # epilogue features: bias + gelu + pre_out, fp32 out + residual + bf16 copy, stats
    m, n, k = 520, 256, 192
    a = torch.randn(m, k, device="cuda", generator=g).bfloat16()
    b = (torch.randn(n, k, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g)
    ref_pre = a.float() @ b.float().t() + bias
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ep = _lib.Epilogue(out=out.data_ptr(), ldo=n, act=_lib.ACT_GELU, bias=bias.data_ptr(), pre_out_bf16=pre.data_ptr())
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), "gemm")
    ok &= report("gemm bias+gelu (act)", torch, _lib, out, torch.nn.functional.gelu(ref_pre))
    ok &= report("gemm bias+gelu (pre)", torch, _lib, pre, ref_pre)
    outf = torch.empty(m, n, device="cuda", dtype=torch.float32)
    cpy = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ep = _lib.Epilogue(out=outf.data_ptr(), ldo=n, out_fp32=1, bias=bias.data_ptr(), residual_f32=res.data_ptr(),
                       out_bf16_copy=cpy.data_ptr())
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), "gemm")
    ok &= report("gemm fp32 out + bias + residual", torch, _lib, outf, ref_pre + res, tol=1e-5)
    ok &= report("gemm fp32 out bf16 copy", torch, _lib, cpy, ref_pre + res)
    # gelu-grad epilogue + masked bf16 add
    h = torch.randn(m, n, device="cuda", generator=g).bfloat16()
    ep = _lib.Epilogue(out=out.data_ptr(), ldo=n, act=_lib.ACT_GELU_GRAD, aux_bf16=h.data_ptr())
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), "gemm")
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).sum().backward()
    ok &= report("gemm gelu-grad epilogue", torch, _lib, out, (a.float() @ b.float().t()) * hf.grad)
    add = torch.randn(m, n, device="cuda", generator=g).bfloat16()
    mask = torch.randn(m, n, device="cuda", generator=g).relu().bfloat16()
    ep = _lib.Epilogue(out=out.data_ptr(), ldo=n, add_bf16=add.data_ptr(), gate_bf16=mask.data_ptr())
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), "gemm")
    ok &= report("gemm masked add", torch, _lib, out, (a.float() @ b.float().t() + add.float()) * (mask.float() > 0))
    # column statistics
    cs = torch.zeros(n, device="cuda")
    cq = torch.zeros(n, device="cuda")
    ep = _lib.Epilogue(out=out.data_ptr(), ldo=n, col_sum=cs.data_ptr(), col_sumsq=cq.data_ptr())
    _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), "gemm")
    torch.cuda.synchronize()
    of = out.float()
    ok &= report("gemm col_sum", torch, _lib, cs[None], of.sum(0)[None], tol=1e-4)
    ok &= report("gemm col_sumsq", torch, _lib, cq[None], (of * of).sum(0)[None], tol=1e-4)
    return ok


def _conv_ref(torch, x_nhwc, w_krsc, stride, pad):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_krsc.float().permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(x, w, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def group_conv():
    torch, _lib = _imports()
    lib = _lib.load()
    st = _lib.current_stream()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ok = True
    g = torch.Generator(device="cuda").manual_seed(2)
    cases = [
        # n, h, w, cin, cout, r, s, stride, pad
        (2, 8, 8, 64, 64, 3, 3, 1, 1),
        (3, 10, 10, 64, 128, 3, 3, 1, 1),
        (2, 20, 20, 128, 128, 3, 3, 2, 1),
        (5, 40, 40, 64, 64, 3, 3, 1, 1),
        (2, 40, 40, 256, 512, 1, 1, 2, 0),
        (3, 22, 22, 128, 64, 3, 3, 2, 1),
        (4, 5, 5, 512, 512, 3, 3, 1, 1),
        (64, 40, 40, 64, 64, 3, 3, 1, 1),
    ]
    for (n, h, w, cin, cout, r, s, stride, pad) in cases:
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, r, s, cin, device="cuda", generator=g) * 0.05).bfloat16()
        ref = _conv_ref(torch, x, wt, stride, pad)
        ho, wo = ref.shape[1], ref.shape[2]
        out = torch.full((n, ho, wo, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        ep = _lib.Epilogue(out=out.data_ptr(), ldo=cout)
        rc = lib.koa_conv_fprop_bf16(x.data_ptr(), wt.data_ptr(), n, h, w, cin, cout, r, s, stride, pad, C.byref(ep), st)
        _lib.check(rc, "conv_fprop")
        ok &= report(f"conv n{n} {h}x{w} c{cin}->{cout} k{r} s{stride} p{pad}", torch, _lib,
                     out.reshape(-1, cout), ref.reshape(-1, cout))
    return ok


WGRAD_VARIANTS = [(8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 256), (1024, 8192, 256)]


def group_wgrad():
    torch, _lib = _imports()
    lib = _lib.load()
    st = _lib.current_stream()
    g = torch.Generator(device="cuda").manual_seed(3)
    good = None
    for var in WGRAD_VARIANTS:
        lib.koa_debug_set_wgrad_desc(*var)
        ok = True
        for (p, cout, cin) in [(64, 128, 64), (128, 128, 64), (640, 128, 128), (5000, 256, 64), (3001, 64, 256)]:
            dy = torch.randn(p, cout, device="cuda", generator=g).bfloat16()
            x = torch.randn(p, cin, device="cuda", generator=g).bfloat16()
            ref = dy.float().t() @ x.float()
            dw = torch.zeros(cout, cin, device="cuda")
            _lib.check(lib.koa_gemm_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), p, cout, cin, st), "wgrad")
            ok &= report(f"wgrad var{var} p{p} cout{cout} cin{cin}", torch, _lib, dw, ref, tol=2e-3)
            if not ok:
                break
        if ok:
            good = var
            break
    print("WGRAD_GOOD_VARIANT", good, flush=True)
    if good is None:
        return False
    # convolution weight gradients with the good variant
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    cases = [
        (2, 8, 8, 64, 64, 3, 3, 1, 1),
        (3, 20, 20, 128, 128, 3, 3, 2, 1),
        (2, 40, 40, 256, 512, 1, 1, 2, 0),
        (6, 10, 10, 64, 256, 3, 3, 1, 1),
        (40, 40, 40, 64, 64, 3, 3, 1, 1),
    ]
    for (n, h, w, cin, cout, r, s, stride, pad) in cases:
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        ho = (h + 2 * pad - r) // stride + 1
        wo = (w + 2 * pad - s) // stride + 1
        dy = torch.randn(n, ho, wo, cout, device="cuda", generator=g).bfloat16()
        wt = torch.zeros(cout, cin, r, s, device="cuda", requires_grad=True)
        y = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt, stride=stride, padding=pad)
        y.backward(dy.float().permute(0, 3, 1, 2))
        ref = wt.grad.permute(0, 2, 3, 1).contiguous()  # [cout, r, s, cin]
        dw = torch.zeros(cout, r, s, cin, device="cuda")
        rc = lib.koa_conv_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), n, h, w, cin, cout, r, s, stride, pad, st)
        _lib.check(rc, "conv_wgrad")
        ok &= report(f"conv_wgrad n{n} {h}x{w} c{cin}->{cout} k{r} s{stride}", torch, _lib,
                     dw.reshape(cout, -1), ref.reshape(cout, -1), tol=2e-3)
    return ok


GROUPS = {"gemm": group_gemm, "conv": group_conv, "wgrad": group_wgrad}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    args = ap.parse_args()
    if args.group:
        ok = GROUPS[args.group]()
        sys.exit(0 if ok else 1)
    status = {}
    for name in GROUPS:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", name], timeout=240)
            status[name] = r.returncode
        except subprocess.TimeoutExpired:
            status[name] = "timeout"
    print("PROBE_STATUS", status, flush=True)
    sys.exit(0 if all(v == 0 for v in status.values()) else 1)


if __name__ == "__main__":
    main()
