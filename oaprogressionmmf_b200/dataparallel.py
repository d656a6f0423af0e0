"""Knee-wise data parallelism: one process per GPU (the reference wraps the model in single-process
``nn.DataParallel``, ``koafusion/run/train_prog_fus.py:84``; ``eval_prog_fus.py:173-174``).

The path shards by knee with no data-path collective; the only exchange is the gradient all-reduce (mean) of
``torch.nn.parallel.DistributedDataParallel`` over NCCL/NVLink, whose buckets fill in reverse execution order:
the three transformers (82 % of the gradient bytes) finish first and are reduced while the CNN backward (91 % of
the FLOPs) still runs. What this module adds on top of DDP is the reference-specific bookkeeping:

* the heads of the per-sequence transformers (``_agg_1/_agg_2[/_agg_3].mlp_head0.*``) are dead compute in the
  reference (``_xrNmrMcP.py:239-240``) and never receive gradients -> frozen before wrapping so DDP does not wait
  for them (the reference leaves ``.grad = None`` there, and Adam skips them);
* BatchNorm statistics stay per replica (DataParallel has no SyncBN) and rank 0's running statistics are the ones
  that persist (``broadcast_buffers=True`` mirrors DataParallel re-broadcasting replica 0 every step);
* checkpoints are written from the unwrapped module, so ``state_dict`` keys stay those of the reference
  (``koafusion/various/_checkpoint.py:56-59``).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

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


def wrap(model: nn.Module, device_ids=None, process_group=None) -> nn.Module:
    """DDP with the reference's DataParallel semantics (see module docstring). ``model`` must already be on its
    device; with world size 1 (or no initialised process group) the model is returned unchanged."""
    import torch.distributed as dist

    freeze_dead_heads(model)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return model
    return nn.parallel.DistributedDataParallel(model, device_ids=device_ids, process_group=process_group,
                                               broadcast_buffers=True, gradient_as_bucket_view=True)


def unwrap(model: nn.Module) -> nn.Module:
    return model.module if isinstance(model, (nn.parallel.DistributedDataParallel, nn.DataParallel)) else model
