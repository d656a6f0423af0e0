"""Batched inference sweep of the full fusion model (BASELINE.json config 5): eval-mode forward, knees/s per batch
size on one GPU (the 8-GPU sweep is 8 independent replicas: no collective in inference).
    python tools/infer_sweep.py [--workload XR1MR3C1CnnTrf] [--batches 1,2,4,8,16,32,64] [--max-gb 150]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oaprogressionmmf_b200.koamodels import dict_models
from oaprogressionmmf_b200.evalpath import predict_batched
from oaprogressionmmf_b200.synthetic import model_config, synthetic_batch, to_attr

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="XR1MR3C1CnnTrf")
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
ap.add_argument("--chunk", type=int, default=32, help="larger batches run as micro-batches of this many knees (eval mode has "
                "no coupling between knees: BatchNorm uses its running statistics)")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--max-gb", type=float, default=150.0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
cfg = model_config(args.workload)
model = dict_models[args.workload](to_attr(cfg), None).to(dev).eval()
rows = []
for b in [int(x) for x in args.batches.split(",")]:
    ins_h, _ = synthetic_batch(cfg, b, 5, pin=True)
    try:
        with torch.no_grad():
            ins = [t.to(dev, non_blocking=True) for t in ins_h]
            for _ in range(2):
                outs = [model(*[t[i:i + args.chunk] for t in ins])["main"] for i in range(0, b, args.chunk)]
            torch.cuda.synchronize()
            if torch.cuda.max_memory_allocated() / 2**30 > args.max_gb:
                raise RuntimeError("over the memory budget")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                # host -> device inside the timed region, pipelined per micro-batch; predictions read back once per batch
                pred, _ = predict_batched(model, ins_h, dev, micro_batch=args.chunk)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        rows.append(dict(batch=b, ms=ms, knees_per_s=b / ms * 1e3, peak_gb=torch.cuda.max_memory_allocated() / 2**30))
        print(json.dumps(rows[-1]), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(batch=b, error=str(e)[:200])), flush=True)
        break
    del ins, outs
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
