import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from oaprogressionmmf_b200.koamodels import dict_models
from oaprogressionmmf_b200.synthetic import model_config, synthetic_batch, to_attr
dev = torch.device("cuda", 0)
cfg = model_config("XR1MR3C1CnnTrf")
model = dict_models["XR1MR3C1CnnTrf"](to_attr(cfg), None).to(dev).eval()
b, chunk = 128, 32
ins_h, _ = synthetic_batch(cfg, b, 5, pin=True)
print("pinned", [t.is_pinned() for t in ins_h])
with torch.no_grad():
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ins = [t.to(dev, non_blocking=True) for t in ins_h]
        torch.cuda.synchronize(); t1 = time.perf_counter()
        times = []
        for i in range(0, b, chunk):
            out = model(*[t[i:i + chunk] for t in ins])["main"]
            torch.cuda.synchronize(); times.append(time.perf_counter())
        print("h2d %.1f ms" % ((t1 - t0) * 1e3), "chunks", ["%.1f" % ((b2 - a) * 1e3) for a, b2 in zip([t1] + times[:-1], times)])
