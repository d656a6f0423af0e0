"""Diagnostic: per-unit BatchNorm statistics produced by the fused GEMM epilogue vs statistics recomputed from the stored conv output."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oaprogressionmmf_b200 import _lib
from oaprogressionmmf_b200.koamodels import SliceEncoder, dict_fes
from oracle import koa_oracle as ko

def run(arch, b, s, size):
    lib = _lib.load()
    dev = "cuda"
    spec = ko.fe_param_spec(arch, "_fe")
    sd = ko.make_state_dict(spec, 11, device=dev)
    enc = SliceEncoder(dict_fes[arch](pretrained=False), with_gap=True).to(dev)
    enc.load_state_dict({k[len("_fe."):]: v.clone() for k, v in sd.items()})
    enc.train()
    g = torch.Generator().manual_seed(5)
    vol = torch.randn(b, 1, size, size, s, generator=g).to(dev)
    tok = enc.encode_volume(vol)
    fn = tok.grad_fn
    while fn is not None and not hasattr(fn, "ws"):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    ws, desc = fn.ws, fn.desc
    n_units = lib.koa_fe_num_units(C.byref(desc))
    off, nb = C.c_size_t(), C.c_size_t()
    print(f"== {arch} b={b} s={s} size={size} units={n_units}")
    for u in range(n_units):
        lib.koa_fe_debug_offset(C.byref(desc), 6, u, C.byref(off), C.byref(nb))
        coef = ws[off.value:off.value + nb.value].view(torch.float32).view(7, -1)
        c = coef.shape[1]
        lib.koa_fe_debug_offset(C.byref(desc), 0, u, C.byref(off), C.byref(nb))
        y = ws[off.value:off.value + nb.value].view(torch.bfloat16).view(-1, c).float()
        mean = y.mean(0); var = y.var(0, unbiased=False)
        em = float((coef[2] - mean).abs().max() / (mean.abs().max() + 1e-6))
        ev = float((coef[3] - torch.rsqrt(var + 1e-5)).abs().max() / torch.rsqrt(var + 1e-5).abs().max())
        flag = "  <<<<" if (em > 1e-3 or ev > 1e-3) else ""
        print(f"unit {u:2d} rows={y.shape[0]:6d} c={c:4d} mean_err={em:.2e} invstd_err={ev:.2e}{flag}")
    torch.cuda.synchronize()
    print("flag", hex(_lib.debug_flag()))

run("resnet50", 2, 3, 64)
run("resnet50", 4, 4, 64)
