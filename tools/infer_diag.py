"""Where does the time of a small eval-mode forward go? device-resident forward vs predict_batched, wall clock and CUDA events."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oaprogressionmmf_b200.evalpath import predict_batched
from oaprogressionmmf_b200.koamodels import dict_models, set_branch_streams
from oaprogressionmmf_b200.synthetic import model_config, synthetic_batch, to_attr
from oaprogressionmmf_b200 import _lib
dev = torch.device("cuda", 0)
cfg = model_config("XR1MR3C1CnnTrf")
model = dict_models["XR1MR3C1CnnTrf"](to_attr(cfg), None).to(dev).eval()
lib = _lib.load()
def timeit(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    l0 = lib.koa_launch_count()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, (lib.koa_launch_count() - l0) // n
for streams in (True, False):
    set_branch_streams(streams)
    for b in (1, 2, 4, 16):
        ins_h, _ = synthetic_batch(cfg, b, 5, pin=True)
        ins = [t.to(dev) for t in ins_h]
        with torch.no_grad():
            r = timeit(lambda: model(*ins)["main"])
            print(f"branch_streams={streams} batch {b}: device-resident forward {r[0]:.2f} ms (wall {r[1]:.2f}), {r[2]} launches", flush=True)
            r = timeit(lambda: predict_batched(model, ins_h, dev, micro_batch=16))
            print(f"   predict_batched {r[0]:.2f} ms (wall {r[1]:.2f})", flush=True)
            for i, name in enumerate(["xr", "dess", "tse", "t2"]):
                fe = getattr(model, f"_fe{i}")
                x = ins[i]
                f = (lambda: fe.encode_image(x)) if i == 0 else (lambda: fe.encode_volume(x))
                r = timeit(f)
                print(f"   extractor {name}: {r[0]:.2f} ms (wall {r[1]:.2f}), {r[2]} launches", flush=True)
