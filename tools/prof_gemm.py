"""Stand-alone driver of koa_gemm_bf16 for ncu: one warm-up + one measured launch per epilogue variant.
    python tools/prof_gemm.py M N K [variants...]   variants: plain stats addgate bwdstats
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
    for v in variants:
        ep = _lib.Epilogue(out=out.data_ptr(), ldo=n)
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
            _lib.check(lib.koa_gemm_bf16(a.data_ptr(), b.data_ptr(), m, n, k, C.byref(ep), st), v)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        extra = {"plain": 0, "stats": 0, "addgate": 2, "bwdstats": 3}[v]
        byts = 2 * (m * k + n * k + m * n * (1 + extra))
        t = min(times[1:])
        print(f"{v:9s} {m}x{n}x{k}: {t*1e3:8.1f} us  {2*m*n*k/t/1e9:7.1f} TFLOP/s  {byts/t/1e6:7.0f} GB/s", flush=True)
    print("flag", hex(_lib.debug_flag()))

main()
