"""Knee-wise data parallelism: one process per GPU (the reference wraps the model in single-process
``nn.DataParallel``, ``koafusion/run/train_prog_fus.py:84``; ``eval_prog_fus.py:173-174``).

The path shards by knee with no data-path collective; the only exchange is the gradient all-reduce (mean) over
NCCL / NVLink. It is issued per *engine call*, not per parameter: every ``koa_fe_backward`` / ``koa_feat_backward``
accumulates the gradients of its whole extractor / transformer into ONE contiguous fp32 buffer (``_lib.zeros_like_flat``),
and as soon as that call returns the buffer goes to ``all_reduce(AVG, async_op=True)``. Backward runs the fusion
transformer first, then the per-sequence transformers, then the CNNs, so the transformers (82 % of the gradient bytes)
are on the wire while the CNN backward (91 % of the FLOPs) still computes; there is no bucket copy, no per-parameter
hook and nine collectives per step for the full model. The few parameters outside the engines (clinical token, XR head)
are reduced in one small flat buffer by an end-of-backward callback, which also waits for the outstanding collectives.

What this module keeps from the reference's DataParallel semantics:

* the heads of the per-sequence transformers (``_agg_1/_agg_2[/_agg_3].mlp_head0.*``) are dead compute in the
  reference (``_xrNmrMcP.py:239-240``) and never receive gradients -> frozen, ``.grad`` stays ``None`` on every rank;
* BatchNorm statistics stay per replica (DataParallel has no SyncBN) and rank 0's running statistics are the ones that
  persist: nothing is exchanged during training, ``sync_buffers`` broadcasts rank 0's buffers on demand (before an
  evaluation on every rank);
* checkpoints are written from the unwrapped module, so ``state_dict`` keys stay those of the reference
  (``koafusion/various/_checkpoint.py:56-59``).

Gradient accumulation over several backward passes is not supported while the sync is active (the collective would
race with autograd's in-place accumulation): call ``zero_grad(set_to_none=True)`` every step, as the reference does.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

DEAD_HEAD_PREFIXES = ("_agg_1.mlp_head0.", "_agg_2.mlp_head0.", "_agg_3.mlp_head0.")


def dead_head_parameters(model: nn.Module) -> List[str]:
    """Names of the parameters that the reference computes with but whose result it discards."""
    return [n for n, _ in model.named_parameters() if n.startswith(DEAD_HEAD_PREFIXES)]


def freeze_dead_heads(model: nn.Module) -> List[str]:
    names = dead_head_parameters(model)
    params = dict(model.named_parameters())
    for n in names:
        params[n].requires_grad_(False)
    return names


def shard_knees(inputs: Sequence[torch.Tensor], target: torch.Tensor, rank: int, world: int
                ) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Contiguous knee shard of a global batch (dim 0 of every modality), as DataParallel's scatter does.
    The global batch must divide evenly: mean-of-local-means then equals the reference's global mean loss."""
    b = target.shape[0]
    if b % world != 0:
        raise ValueError(f"global batch {b} does not divide over {world} ranks")
    per = b // world
    sl = slice(rank * per, (rank + 1) * per)
    return [t[sl] for t in inputs], target[sl]


# ----------------------------------------------------------------------------------------------------------------
# gradient synchronisation state (one model per process)
# ----------------------------------------------------------------------------------------------------------------
class _SyncState:
    def __init__(self):
        self.group = None
        self.world = 1
        self.enabled = False
        self.use_avg = False          # NCCL has ReduceOp.AVG; gloo (CPU tests) sums and divides
        self.handles = []             # (work, flat) of the engine buffers in flight
        self.covered = set()          # ids of parameters whose gradient lives in a synced flat buffer
        self.loose: List[nn.Parameter] = []   # parameters outside the engines
        self.callback_queued = False
        self.collectives = 0          # issued since the last reset (bench / tests)


_state = _SyncState()


def active() -> bool:
    """True while a data-parallel wrapper synchronises gradients (the extractors then run their backward in stages)."""
    return _state.enabled


def sync_flat(flat: Optional[torch.Tensor], params: Sequence[nn.Parameter]) -> None:
    """Called by the engine autograd Functions right after (a stage of) their backward call returned: ``flat`` holds
    the freshly computed gradients of ``params`` (contiguous fp32). No-op unless ``wrap`` activated the synchronisation."""
    st = _state
    if not st.enabled or flat is None or flat.numel() == 0:
        return
    import torch.distributed as dist

    for p in params:
        if p.grad is not None:
            raise RuntimeError("gradient accumulation across backward passes is not supported with the data-parallel "
                               "gradient sync: call zero_grad(set_to_none=True) before every backward")
        st.covered.add(id(p))
    op = dist.ReduceOp.AVG if st.use_avg else dist.ReduceOp.SUM
    work = dist.all_reduce(flat, op=op, group=st.group, async_op=True)
    st.handles.append((work, flat))
    st.collectives += 1
    _queue_finish()


def _queue_finish() -> None:
    st = _state
    if not st.callback_queued:
        st.callback_queued = True
        torch.autograd.Variable._execution_engine.queue_callback(_finish)


def _finish() -> None:
    """End of the backward pass: reduce the gradients of the loose parameters in one flat buffer, then wait for every
    outstanding collective (the current stream waits, the host does not)."""
    st = _state
    import torch.distributed as dist

    st.callback_queued = False
    loose = [p for p in st.loose if p.grad is not None and id(p) not in st.covered]
    if loose:
        flat = torch.cat([p.grad.reshape(-1) for p in loose])
        op = dist.ReduceOp.AVG if st.use_avg else dist.ReduceOp.SUM
        dist.all_reduce(flat, op=op, group=st.group)
        st.collectives += 1
        if not st.use_avg:
            flat.div_(st.world)
        off = 0
        for p in loose:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    for work, flat in st.handles:
        work.wait()
        if not st.use_avg:
            flat.div_(st.world)
    st.handles.clear()
    st.covered.clear()


def _loose_hook(grad):
    if _state.enabled:
        _queue_finish()
    return grad


class KneeParallel(nn.Module):
    """Thin data-parallel wrapper: forwards to the wrapped module, activates the flat-buffer gradient sync, makes all
    ranks start from rank 0's parameters and buffers. ``.module`` is the wrapped model (as with DDP / DataParallel)."""

    def __init__(self, module: nn.Module, process_group=None):
        super().__init__()
        import torch.distributed as dist

        self.module = module
        st = _state
        st.group = process_group
        st.world = dist.get_world_size(process_group)
        st.use_avg = dist.get_backend(process_group) == "nccl"
        st.enabled = True
        st.handles.clear()
        st.covered.clear()
        st.collectives = 0
        # parameters owned by an engine (SliceEncoder / FeaT) arrive through sync_flat; the rest is "loose"
        from .koamodels import FeaT, SliceEncoder  # local import: koamodels imports this module

        owned = set()
        for m in module.modules():
            if isinstance(m, (FeaT, SliceEncoder)):
                owned.update(id(p) for p in m.parameters())
        st.loose = [p for p in module.parameters() if p.requires_grad and id(p) not in owned]
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
        # a backward pass that reaches a loose parameter needs the end-of-backward reduction even without an engine call
        for p in st.loose:
            p.register_hook(_loose_hook)

    def forward(self, *args, **kwargs):
        # A backward pass that raised (e.g. an out-of-memory error the caller caught) never ran its final callback: drain
        # what it left behind so that this step queues its own callback and no stale collective is mistaken for a new one.
        st = _state
        if st.callback_queued or st.handles:
            for work, _ in st.handles:
                work.wait()
            st.handles.clear()
            st.covered.clear()
            st.callback_queued = False
        return self.module(*args, **kwargs)

    def sync_buffers(self) -> None:
        """Rank 0's BatchNorm running statistics to every rank (DataParallel keeps replica 0's)."""
        import torch.distributed as dist

        group = _state.group
        src = dist.get_global_rank(group, 0) if group is not None else 0  # group rank 0, as in __init__
        with torch.no_grad():
            for b in self.module.buffers():
                dist.broadcast(b, src=src, group=group)


def wrap(model: nn.Module, device_ids=None, process_group=None) -> nn.Module:
    """Data-parallel wrapper with the reference's DataParallel semantics (see module docstring). ``model`` must already
    be on its device; with world size 1 (or no initialised process group) the model is returned unchanged."""
    import torch.distributed as dist

    freeze_dead_heads(model)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return model
    return KneeParallel(model, process_group)


def unwrap(model: nn.Module) -> nn.Module:
    return model.module if isinstance(model, (KneeParallel, nn.parallel.DistributedDataParallel, nn.DataParallel)) else model


def disable() -> None:
    """Turn the gradient sync off (tests; a process that tears its process group down)."""
    _state.enabled = False
    _state.handles.clear()
    _state.covered.clear()
    _state.loose = []


def collectives_issued() -> int:
    return _state.collectives
