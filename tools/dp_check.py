"""2+ ranks: the staged, overlapped gradient all-reduce of dataparallel.KneeParallel gives the mean of the ranks' local
gradients (computed with the synchronisation off and averaged by hand), up to the run-to-run scatter of the backward pass.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oaprogressionmmf_b200 import dataparallel as dp
from oaprogressionmmf_b200.koamodels import dict_models
from oaprogressionmmf_b200.losses import FocalLoss
from oracle import koa_oracle as ko
from tests.util import rel, to_attr

rank, local, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "NONE")
dist.init_process_group("nccl", device_id=dev)
name = "XR1MR2C1CnnTrf"
cfg = ko.make_config(name, xr_size=160, mr_size=96, slices=(12, 8), depth=2)
torch.manual_seed(778)
# eval mode with gradients enabled: BatchNorm uses its running statistics and dropout is off, so the backward pass is well
# conditioned and repeatable (a train-mode BatchNorm ResNet at random initialisation amplifies the summation-order noise
# of its statistics to tens of percent of the early gradients on a batch this small, which would hide any error here)
model = dict_models[name](to_attr(cfg), None).to(dev).eval()
wrapped = dp.wrap(model)
inputs, target = ko.make_inputs(name, cfg, 4, 100 + rank, device=dev)
loss_fn = FocalLoss(gamma=2)

def grads(sync):
    dp._state.enabled = sync
    wrapped.zero_grad(set_to_none=True)
    loss_fn(wrapped(*inputs)["main"], target).backward()
    torch.cuda.synchronize()
    return {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}

a, b = grads(False), grads(False)
mean = {}
for k, g in a.items():
    if g is None:
        continue
    m = g.clone()
    dist.all_reduce(m)
    mean[k] = m / ws
s = grads(True)
assert dp.collectives_issued() > 0
worst, scat = 0.0, 0.0
for k, m in mean.items():
    assert s[k] is not None, k
    scat = max(scat, rel(b[k], a[k]))
    worst = max(worst, rel(s[k], m))
dead = [k for k, g in s.items() if g is None]
# every rank holds the same reduced gradients, bit for bit (nothing was left unreduced)
same = 0.0
for k, m in mean.items():
    g0 = s[k].clone()
    dist.broadcast(g0, src=0)
    same = max(same, float((g0 - s[k]).abs().max()))
t = torch.tensor([worst, scat, same], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_check: {ws} ranks, {len(mean)} tensors, worst rel(synced, mean of local) = {float(t[0]):.3e}, local run-to-run "
          f"scatter = {float(t[1]):.3e}, dead heads without gradient: {len(dead)}, collectives per step: {dp.collectives_issued()}")
    assert float(t[0]) <= 3 * float(t[1]) + 1e-4, "the staged all-reduce does not reproduce the mean of the local gradients"
    assert float(t[1]) < 1e-2, "the eval-mode backward pass should be repeatable"
    assert float(t[2]) == 0.0, "ranks disagree on the reduced gradients"
    print("dp_check OK")
dist.destroy_process_group()
