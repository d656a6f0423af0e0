"""Per-parameter gradient-norm ratios (ours / reference fixture) of one golden train step.
    python tools/diag_train_case.py MR3CnnTrf"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_parity import _golden_case
from oaprogressionmmf_b200.losses import FocalLoss

case = sys.argv[1]
gdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
gold, cfg, model, inputs, target = _golden_case(case, gdir, torch.device("cuda"), pos_scale=0.02)
model.train()
lg = model(*inputs)["main"]
loss = FocalLoss(gamma=2)(lg, target)
loss.backward()
print("loss", float(loss), "ref", gold["train_loss"])
groups = {}
for k, p in model.named_parameters():
    gref = gold["grads"][k]
    if gref is None:
        continue
    r = float(p.grad.norm()) / (gref["norm"] + 1e-30)
    groups.setdefault(k.split(".")[0], []).append(r)
for g, v in groups.items():
    t = torch.tensor(v)
    print(f"{g:12s} n={len(v):4d} min {t.min():.3f} med {t.median():.3f} max {t.max():.3f}")
