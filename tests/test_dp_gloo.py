"""N > 1 host logic on CPU: world_size-2 gloo run of the knee-wise data-parallel wrapper
(oaprogressionmmf_b200/dataparallel.py). The CUDA engines cannot run here, so the wrapped module is a small CPU
stand-in with the same structure the wrapper cares about: BatchNorm buffers, a live fusion head, dead per-sequence
heads named like the reference's, and one sub-module that behaves like an engine: its backward fills ONE flat
gradient buffer and hands it to ``dataparallel.sync_flat`` (the path every koa_fe_backward / koa_feat_backward takes)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from oaprogressionmmf_b200 import dataparallel as dp


class _Head(nn.Module):
    def __init__(self):
        super().__init__()
        self.mlp_head0 = nn.Sequential(nn.LayerNorm(8), nn.Linear(8, 8), nn.GELU(), nn.Dropout(0.0), nn.Linear(8, 2))
        self.body = nn.Linear(8, 8)


class _EngineFn(torch.autograd.Function):
    """y = x @ W^T + b computed like an engine: gradients of (W, b) land in one flat buffer that is all-reduced async."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        from oaprogressionmmf_b200 import _lib

        (gw, gb), flat = _lib.zeros_like_flat([w, gy.new_zeros(w.shape[0])])
        gw.copy_(gy.t() @ x)
        gb.copy_(gy.sum(0))
        # in two slices of the one buffer, as the extractors hand over their stages (koamodels/_fe.py: layer4 first)
        cut = (gb.data_ptr() - flat.data_ptr()) // 4
        dp.sync_flat(flat[:cut], [ctx.w_param])
        dp.sync_flat(flat[cut:], [ctx.b_param])
        return gy @ w, gw, gb


class _EngineLike(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(8, 8) * 0.3)
        self.bias = nn.Parameter(torch.zeros(8))

    def forward(self, x):
        fn = _EngineFn
        out = fn.apply(x, self.weight, self.bias)
        out.grad_fn.w_param, out.grad_fn.b_param = self.weight, self.bias
        return out


class _Toy(nn.Module):
    """Hierarchical like XR1MR2C1CnnTrf: two per-sequence aggregators whose heads are dead, one fusion head."""

    def __init__(self):
        super().__init__()
        self._fe1 = nn.Sequential(nn.Linear(4, 8), nn.BatchNorm1d(8))
        self._agg_1, self._agg_2, self._agg_final = _Head(), _Head(), _Head()
        self._eng = _EngineLike()

    def forward(self, x0, x1):
        a = self._agg_1.body(self._fe1(x0))
        b = self._agg_2.body(self._fe1(x1))
        return self._agg_final.mlp_head0(self._eng(self._agg_final.body(a + b)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = _Toy()
        with torch.no_grad():  # ranks start with different BN buffers; rank 0's must win
            model._fe1[1].running_mean.fill_(float(rank))
        dp._state.__init__()
        wrapped = dp.wrap(model)
        assert isinstance(wrapped, dp.KneeParallel)
        # the engine-like parameters are not "loose": pretend the type check found them
        dp._state.loose = [p for n, p in model.named_parameters() if p.requires_grad and not n.startswith("_eng.")]
        g = torch.Generator().manual_seed(1)
        x0, x1 = torch.randn(8, 4, generator=g), torch.randn(8, 4, generator=g)
        tgt = torch.randint(0, 2, (8,), generator=g)
        (s0, s1), st = dp.shard_knees([x0, x1], tgt, rank, world)
        assert s0.shape[0] == 4
        loss = nn.functional.cross_entropy(wrapped(s0, s1), st)
        loss.backward()
        grads = {n: (None if p.grad is None else p.grad.clone()) for n, p in dp.unwrap(wrapped).named_parameters()}
        res = dict(grads=grads, loss=float(loss.detach()), rm=model._fe1[1].running_mean.clone(),
                   keys=list(dp.unwrap(wrapped).state_dict().keys()), collectives=dp.collectives_issued())
        # accumulating into existing gradients is refused (the async collective would race with autograd's +=)
        try:
            nn.functional.cross_entropy(wrapped(s0, s1), st).backward()
            res["second_backward"] = "ok"
        except RuntimeError as e:
            res["second_backward"] = str(e)
        dp.disable()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ddp_wrapper_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    # gradients are identical on both ranks (all-reduced) and dead heads stay None, like the reference
    for n, g in r0["grads"].items():
        if n.startswith(dp.DEAD_HEAD_PREFIXES):
            assert g is None and r1["grads"][n] is None, n
        else:
            assert g is not None, n
            assert torch.equal(g, r1["grads"][n]), n
    # all-reduced gradient == mean of the per-rank gradients computed without DDP on the same shards
    torch.manual_seed(0)
    ref = _Toy()
    dp.freeze_dead_heads(ref)
    gen = torch.Generator().manual_seed(1)
    x0, x1 = torch.randn(8, 4, generator=gen), torch.randn(8, 4, generator=gen)
    tgt = torch.randint(0, 2, (8,), generator=gen)
    acc = None
    for rank in range(world):
        ref.zero_grad()
        (s0, s1), st = dp.shard_knees([x0, x1], tgt, rank, world)
        nn.functional.cross_entropy(ref(s0, s1), st).backward()
        gs = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
        acc = gs if acc is None else {n: acc[n] + gs[n] for n in gs}
    for n, g in acc.items():
        assert torch.allclose(r0["grads"][n], g / world, rtol=1e-5, atol=1e-6), n
    assert r0["collectives"] == 3  # two stage slices of the engine buffer + one flat buffer of loose parameters
    assert "zero_grad" in r0["second_backward"]
    # checkpoint keys are those of the unwrapped module (no "module." prefix)
    assert all(not k.startswith("module.") for k in r0["keys"])
    # rank 0's parameters and BatchNorm buffers were broadcast at wrap time: rank 1 started from rank 0's running_mean
    assert float(r1["rm"].abs().max()) < 1.0


def test_shard_knees_rejects_ragged_batches():
    x = torch.zeros(5, 1)
    with pytest.raises(ValueError):
        dp.shard_knees([x], torch.zeros(5), 0, 2)


def test_dead_head_names_follow_the_reference():
    m = _Toy()
    names = dp.freeze_dead_heads(m)
    assert len(names) == 12  # 6 tensors per dead head, two per-sequence transformers
    assert all(not dict(m.named_parameters())[n].requires_grad for n in names)
    assert all(p.requires_grad for n, p in m.named_parameters() if n.startswith("_agg_final."))
    assert dp.wrap(m) is m  # no process group -> unchanged
