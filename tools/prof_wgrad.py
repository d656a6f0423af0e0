"""Stand-alone timing of the weight-gradient kernels (CUDA events, L2 flushed between launches).
    python tools/prof_wgrad.py            # a fixed list of shapes from ResNet-50 @160 (B=8) and the transformers"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oaprogressionmmf_b200 import _lib

lib = _lib.load()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream

def timeit(fn):
    ts = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])

# (pixels, cout, cin) 1x1
for pixels, cout, cin in [(819200, 256, 64), (819200, 64, 256), (204800, 512, 128), (51200, 1024, 256), (51200, 256, 1024), (12800, 2048, 512), (520, 2048, 2048), (520, 6144, 2048)]:
    dy = torch.randn(pixels, cout, device=dev).bfloat16(); x = torch.randn(pixels, cin, device=dev).bfloat16()
    dw = torch.zeros(cout, cin, device=dev)
    t = timeit(lambda: _lib.check(lib.koa_gemm_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), pixels, cout, cin, 0, st), "w"))
    by = 2 * pixels * (cout + cin) + 4 * cout * cin
    print(f"1x1 P={pixels:7d} Cout={cout:5d} Cin={cin:5d}: {t*1e3:8.1f} us {2*pixels*cout*cin/t/1e9:7.1f} TF {by/t/1e6:6.0f} GB/s", flush=True)
# 3x3 (n_img, h, cin, cout, stride)
for n, h, cin, cout, stride in [(512, 40, 64, 64, 1), (512, 20, 128, 128, 1), (512, 10, 256, 256, 1), (512, 5, 512, 512, 1), (512, 40, 128, 128, 2)]:
    ho = (h + 2 - 3) // stride + 1
    dy = torch.randn(n, ho, ho, cout, device=dev).bfloat16(); x = torch.randn(n, h, h, cin, device=dev).bfloat16()
    dw = torch.zeros(cout, 3, 3, cin, device=dev)
    t = timeit(lambda: _lib.check(lib.koa_conv_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), n, h, h, cin, cout, 3, 3, stride, 1, 0, st), "w"))
    fl = 2 * n * ho * ho * cout * cin * 9
    print(f"3x3 n={n} h={h:3d} Cin={cin:4d} Cout={cout:4d} s={stride}: {t*1e3:8.1f} us {fl/t/1e9:7.1f} TF", flush=True)
print("flag", hex(_lib.debug_flag()))
