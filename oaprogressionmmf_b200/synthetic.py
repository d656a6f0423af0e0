"""Synthetic stand-in for the OAI data pipeline (``koafusion/datasets``): batches with the shapes, dtypes and
value ranges ``DatasetOAI3d.__getitem__`` + the transform chain deliver to the train loop
(``koafusion/datasets/oai/_dataset.py:250-329``, ``_data_provider.py:297-334``): z-normalised images / volumes
(slice axis innermost), the 9-element clinical vector (three z-scores, three one-hot pairs) and Bernoulli(0.12)
progression targets (prior from ``koafusion/various/_metrics_stat_anlys.py:107``)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def model_config(name: str, *, xr_size=350, mr_size=160, slices=(64, 32, 25), depth=4, heads=8, mlp_dim=2048,
                 xr_arch="resnext50_32x4d", mr_arch="resnet50", dropout=0.0, output_type="dict") -> dict:
    """Model config (the keys the koafusion constructors read) for the full-size recipes of runner.sh:85-363.
    ``slices``: slice counts of the MRI inputs in the order the class takes them."""
    base = dict(name=name, debug=False, downscale=False, input_channels=1, output_channels=2, output_type=output_type,
                pretrained=False, path_pretrained=None, restore_weights=False)
    agg = dict(depth=depth, heads=heads, emb_dropout=dropout, mlp_dim=mlp_dim, mlp_dropout=dropout)
    mr = lambda s: [mr_size, mr_size, s]  # noqa: E731
    fe2 = dict(xr=dict(arch=xr_arch, pretrained=False, with_gap=True, dropout=dropout),
               mr=dict(arch=mr_arch, pretrained=False, with_gap=True, dropout=dropout))
    clin = dict(dim_in=9, dim_out=2048, dropout=dropout)
    if name == "XR1Cnn":
        return dict(base, input_size=[[xr_size, xr_size]], fe=dict(arch=xr_arch, pretrained=False, with_gap=True, dropout=0.0),
                    agg=dict(hidden_size=512, dropout=0.5 if dropout else 0.0))
    if name == "MR1CnnTrf":
        return dict(base, input_size=[mr(slices[0])],
                    fe=dict(arch=mr_arch, pretrained=False, with_gap=True, dropout=dropout, dims_view="rc"),
                    agg=dict(agg, num_slices=slices[0]))
    if name == "MR2CnnTrf":
        return dict(base, input_size=[mr(slices[0]), mr(slices[1])],
                    fe=dict(arch=mr_arch, pretrained=False, with_gap=True, dropout=dropout),
                    agg=dict(agg, num_slices=[slices[0], slices[1]]))
    if name == "XR1MR1CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size], mr(slices[0])], fe=fe2, agg=dict(agg, num_slices=[1, slices[0]]))
    if name == "XR1MR2CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size], mr(slices[0]), mr(slices[1])], fe=fe2,
                    agg=dict(agg, num_slices=[1, slices[0], slices[1]]))
    if name == "XR1MR2C1CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size], mr(slices[0]), mr(slices[1]), [16]], fe=dict(fe2, clin=clin),
                    agg=dict(agg, num_slices=[1, slices[0], slices[1], 1]))
    if name == "MR3CnnTrf":
        return dict(base, input_size=[mr(s) for s in slices[:3]], fe=fe2, agg=dict(agg, num_slices=list(slices[:3])))
    if name == "XR1MR3C1CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size]] + [mr(s) for s in slices[:3]] + [[16]], fe=dict(fe2, clin=clin),
                    agg=dict(agg, num_slices=[1] + list(slices[:3]) + [1]))
    raise ValueError(name)


class AttrDict(dict):
    """Config container with item and attribute access (the reference reads its OmegaConf both ways)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return [to_attr(v) for v in d]
    return d


def synthetic_batch(cfg: dict, batch: int, seed: int, pin: bool = False) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """One host-side batch: positional model inputs + target, in the layout the train loop moves to the device
    (``koafusion/run/train_prog_fus.py:136-140``)."""
    g = torch.Generator().manual_seed(seed)
    ins = []
    for shape in cfg["input_size"]:
        if len(shape) == 1:
            z = torch.randn(batch, 3, generator=g)
            oh = [torch.nn.functional.one_hot(torch.randint(0, 2, (batch,), generator=g), 2).float() for _ in range(3)]
            t = torch.cat([z[:, 0:1], oh[0], z[:, 1:2], oh[1], oh[2], z[:, 2:3]], dim=1)[:, None, :]
        else:
            t = torch.randn(batch, 1, *shape, generator=g)
        ins.append(t.pin_memory() if pin else t)
    target = (torch.rand(batch, generator=g) < 0.12).long()
    return ins, (target.pin_memory() if pin else target)


class SyntheticKneeLoader:
    """Endless iterator of host batches (optionally pinned), ``n_distinct`` distinct batches cycled."""

    def __init__(self, cfg: dict, batch: int, seed: int = 779, n_distinct: int = 2, pin: bool = True):
        self.batches = [synthetic_batch(cfg, batch, seed + i, pin=pin) for i in range(n_distinct)]
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return b


class DevicePrefetcher:
    """Host batches -> device batches with the copy of batch i+1 issued on a copy stream while step i computes (what
    the reference's ``DataLoader(pin_memory=True)`` + ``.to(device, non_blocking=True)`` loop amounts to,
    ``koafusion/run/train_prog_fus.py:136-140``). Every batch is copied from pinned host memory, none is reused."""

    def __init__(self, loader, device):
        self.loader = iter(loader)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self._issue()

    def _issue(self):
        ins_h, tgt_h = next(self.loader)
        with torch.cuda.stream(self.copy_stream):
            ins = [t.to(self.device, non_blocking=True) for t in ins_h]
            tgt = tgt_h.to(self.device, non_blocking=True)
        self.next = (ins, tgt)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.copy_stream)
        ins, tgt = self.next
        for t in ins + [tgt]:
            t.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
        self._issue()
        return ins, tgt


def input_bytes(ins: Sequence[torch.Tensor], target: torch.Tensor) -> int:
    return sum(t.numel() * t.element_size() for t in ins) + target.numel() * target.element_size()
