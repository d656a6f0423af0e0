"""CUDA-event timing of koa_attention_{fwd,bwd}_fmt at the transformer shapes of the full model (KOA_ATTN_TC selects the kernels)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oaprogressionmmf_b200 import _lib
lib = _lib.load()
st = _lib.current_stream()
for b, n in ((16, 124), (16, 64), (16, 32), (16, 25), (32, 124)):
    heads, hd = 8, 256
    d = heads * hd
    for f16 in (1, 0):
        qkv = torch.randn(b * n, 3 * d, device="cuda").to(torch.float16 if f16 else torch.bfloat16)
        out = torch.empty(b * n, d, dtype=qkv.dtype, device="cuda")
        probs = torch.empty(b, heads, n, n, device="cuda")
        dout = torch.randn(b * n, d, device="cuda").bfloat16()
        dqkv = torch.empty(b * n, 3 * d, dtype=torch.bfloat16, device="cuda")
        def fwd():
            _lib.check(lib.koa_attention_fwd_fmt(qkv.data_ptr(), out.data_ptr(), probs.data_ptr(), b, n, heads, hd, d ** -0.5, f16, st), "f")
        def bwd():
            _lib.check(lib.koa_attention_bwd_fmt(qkv.data_ptr(), probs.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), b, n, heads, hd, d ** -0.5, f16, st), "b")
        res = []
        for fn in (fwd, bwd):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 20 * 1e3)
        print(f"TC={os.environ.get('KOA_ATTN_TC','1')} b={b} n={n} f16={f16}: fwd {res[0]:.1f} us, bwd {res[1]:.1f} us", flush=True)
