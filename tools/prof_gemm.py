"""Stand-alone driver of koa_gemm_bf16 for ncu: one warm-up + one measured launch per epilogue variant.
    python tools/prof_gemm.py M N K [variants...]   variants: plain stats addgate bwdstats bnapply (conv -> BatchNorm -> + identity
    -> ReLU in the epilogue, fp16, gemm_conv_kernel MODE 2) kcat (K-concatenated data gradient of the y-free tail: [A | A2] with
    K2 = N, column bias, gate, BatchNorm-backward sums)
Prints CUDA-event times (not valid under ncu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oaprogressionmmf_b200 import _lib

def main():
    m, n, k = map(int, sys.argv[1:4])
    variants = sys.argv[4:] or ["plain", "stats", "addgate", "bwdstats"]
    lib = _lib.load()
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    a = torch.randn(m, k, generator=g).to(dev).bfloat16()
    b = (torch.randn(n, k, generator=g) * k ** -0.5).to(dev).bfloat16()
    out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    add = torch.randn(m, n, generator=g).to(dev).bfloat16()
    gate = torch.randn(m, n, generator=g).to(dev).bfloat16()
    y = torch.randn(m, n, generator=g).to(dev).bfloat16()
    s, q = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    mean, invstd = torch.zeros(n, device=dev), torch.ones(n, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    a16, b16 = a.half(), b.half()
    res16, out16 = add.half(), y.half()
    scale, shift = torch.rand(n, device=dev) + 0.5, torch.randn(n, device=dev) * 0.1
    a2 = torch.randn(m, n, generator=g).to(dev).bfloat16()
    bcat = (torch.randn(n, k + n, generator=g) * (k + n) ** -0.5).to(dev).bfloat16()
    bias = torch.randn(n, device=dev) * 0.1
    for v in variants:
        ep = _lib.Epilogue(out=out.data_ptr(), ldo=n)
        if v == "bnapply":
            ep = _lib.Epilogue(out=out16.data_ptr(), ldo=n, a_f16=1, b_f16=1, out_f16=1, act_f16=1, bn_scale=scale.data_ptr(),
                               bn_shift=shift.data_ptr(), add_bf16=res16.data_ptr(), act=_lib.ACT_RELU, out_bf16_copy=out.data_ptr())
        if v == "kcat":
            ep = _lib.Epilogue(out=out.data_ptr(), ldo=n, act_f16=1, col_bias=bias.data_ptr(), gate_bf16=res16.data_ptr(),
                               col_sum=s.data_ptr(), col_sumsq=q.data_ptr(), stat_y=out16.data_ptr(), stat_mean=mean.data_ptr(),
                               stat_invstd=invstd.data_ptr())
        if v in ("stats", "bwdstats"):
            ep.col_sum, ep.col_sumsq = s.data_ptr(), q.data_ptr()
        if v in ("addgate", "bwdstats"):
            ep.add_bf16, ep.gate_bf16 = add.data_ptr(), gate.data_ptr()
        if v == "bwdstats":
            ep.stat_y, ep.stat_mean, ep.stat_invstd = y.data_ptr(), mean.data_ptr(), invstd.data_ptr()
        times = []
        for it in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if v == "bnapply":
                _lib.check(lib.koa_gemm_bf16(a16.data_ptr(), b16.data_ptr(), m, n, k, C.byref(ep), st), v)
            elif v == "kcat":
                _lib.check(lib.koa_gemm_kcat_bf16(a.data_ptr(), k, a2.data_ptr(), n, bcat.data_ptr(), m, n, C.byref(ep), st), v)
            else:
                _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), v)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        extra = {"plain": 0, "stats": 0, "addgate": 2, "bwdstats": 3, "bnapply": 2, "kcat": 3}[v]
        byts = 2 * (m * k + n * k + m * n * (1 + extra))
        t = min(times[1:])
        print(f"{v:9s} {m}x{n}x{k}: {t*1e3:8.1f} us  {2*m*n*k/t/1e9:7.1f} TFLOP/s  {byts/t/1e6:7.0f} GB/s", flush=True)
    print("flag", hex(_lib.debug_flag()))

main()
