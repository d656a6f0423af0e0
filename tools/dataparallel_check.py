"""nn.DataParallel (single process, one Python thread per GPU: what the reference's trainer does, run/train_prog_fus.py:84)
around a koamodels class on 2 GPUs: eval-mode logits equal those of the unwrapped model, a train step produces finite
gradients for every live parameter. (Multi-GPU training should use dataparallel.wrap = one process per GPU; this checks that
the per-device library state added in round 1 holds when the reference's own wrapper is kept.)
    python tools/dataparallel_check.py        # needs >= 2 visible GPUs"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oaprogressionmmf_b200 import _lib
from oaprogressionmmf_b200.koamodels import dict_models
from oaprogressionmmf_b200.losses import FocalLoss
from oracle import koa_oracle as ko
from tests.util import rel, to_attr

assert torch.cuda.device_count() >= 2, "needs two GPUs"
name = "XR1MR2C1CnnTrf"
cfg = ko.make_config(name, xr_size=160, mr_size=96, slices=(12, 8), depth=2)
torch.manual_seed(778)
model = dict_models[name](to_attr(cfg), None).to("cuda:0")
inputs, target = ko.make_inputs(name, cfg, 4, 100, device="cuda:0")
model.eval()
with torch.no_grad():
    ref = model(*inputs)["main"]
dp = torch.nn.DataParallel(model, device_ids=[0, 1])
with torch.no_grad():
    got = dp(*inputs)["main"]
print("eval logits DataParallel vs single device: rel", rel(got, ref))
assert rel(got, ref) < 1e-5
dp.train()
loss = FocalLoss(gamma=2)(dp(*inputs)["main"], target)
loss.backward()
torch.cuda.synchronize()
n = 0
for k, p in model.named_parameters():
    if k.startswith(("_agg_1.mlp_head0", "_agg_2.mlp_head0")):
        continue
    assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    n += 1
assert _lib.debug_flag() == 0
print(f"train step through nn.DataParallel on 2 GPUs: loss {float(loss):.4f}, {n} finite gradients; dataparallel_check OK")
