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
    ``koafusion/run/train_prog_fus.py:136-140``). Every batch is copied from pinned host memory inside the step that
    precedes its use; the device tensors handed out are two alternating buffer sets."""

    def __init__(self, loader, device):
        self.loader = iter(loader)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # Two sets of device buffers allocated once on the compute stream and filled in turn by the copy stream (no tensor
        # changes streams; blocks migrating between the per-stream pools of the caching allocator stall a model that
        # runs its branches on several streams). The batch handed out by __next__ stays valid until the call after next.
        self.bufs = [None, None]
        self.filled = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self.calls = 0
        self._issue()

    def _issue(self):
        ins_h, tgt_h = next(self.loader)
        k = self.k
        cur = torch.cuda.current_stream(self.device)
        if self.bufs[k] is None or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(self.bufs[k][0], ins_h)):
            self.bufs[k] = ([torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in ins_h],
                            torch.empty(tgt_h.shape, dtype=tgt_h.dtype, device=self.device))
        # everything queued on the compute stream so far (the step that read this set two calls ago) precedes the refill
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.bufs[k][0], ins_h):
                dst.copy_(src, non_blocking=True)
            self.bufs[k][1].copy_(tgt_h, non_blocking=True)
            self.filled[k].record(self.copy_stream)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        k = self.k
        cur.wait_event(self.filled[k])
        ins, tgt = self.bufs[k]
        self.k ^= 1
        self._issue()  # the next batch goes into the other set while this one is consumed
        return ins, tgt


def input_bytes(ins: Sequence[torch.Tensor], target: torch.Tensor) -> int:
    return sum(t.numel() * t.element_size() for t in ins) + target.numel() * target.element_size()


# ---------------------------------------------------------------------------------------------------------------------
# the item / batch contract of the reference datasets
# ---------------------------------------------------------------------------------------------------------------------
# per sequence: stored (on-disk) minimum size, storage type, PTNormalize statistics, gamma augmentation or not
# (koafusion/datasets/oai/_dataset.py:277-291, koafusion/datasets/_data_provider.py:297-334)
MODAL_SPECS = {
    "sag_3d_dess": dict(stored=(320, 320, 128), dtype=torch.uint8, mean=0.257, std=0.235, gamma=True),
    "cor_iw_tse": dict(stored=(320, 320, 32), dtype=torch.uint8, mean=0.455, std=0.290, gamma=True),
    "sag_t2_map": dict(stored=(320, 320, 25), dtype=torch.uint8, mean=0.259, std=0.345, gamma=False),
    "xr_pa": dict(stored=(700, 700), dtype=torch.uint16, mean=0.543, std=0.296, gamma=True),
}


def encode_clinical(age: float, sex: str, bmi: float, inj: int, surg: int, womac: float) -> torch.Tensor:
    """The 9-element clinical vector of ``DatasetOAI3d.__getitem__`` (``koafusion/datasets/oai/_dataset.py:254-266``):
    z-scored age, one-hot sex (MALE first), z-scored BMI, one-hot injury, one-hot surgery, z-scored WOMAC."""
    vec = [(age - 60.945) / 9.209]
    vec += [1, 0] if sex == "MALE" else [0, 1]
    vec.append((bmi - 28.734) / 4.917)
    for flag in (inj, surg):
        onehot = [0.0, 0.0]
        onehot[flag] = 1.0
        vec += onehot
    vec.append((womac - 10.940) / 14.573)
    return torch.tensor(vec, dtype=torch.float32)


class SyntheticKneeDataset(torch.utils.data.Dataset):
    """Map-style dataset with the item contract of ``DatasetOAI3d.__getitem__`` (``_dataset.py:250-329``): a dict with
    ``image__{modal}`` per modality ((CH, R, C[, S]) image, (1, 9) clinical vector), ``target`` (numpy array of one
    element), ``("-", "exam_knee_id")``, ``("-", "side")``. A ``DataLoader`` with the default collate turns it into the
    batches the reference's train / eval loops read (``run/train_prog_fus.py:136-140``, ``run/eval_prog_fus.py:250-312``).

    ``stored=False``: images as the reference's CPU transform chain delivers them (z-normalised fp32 of ``sizes[modal]``).
    ``stored=True``: images *as stored* (integer type and on-disk size of ``MODAL_SPECS``, or ``stored_sizes[modal]``) plus,
    under ``state__{modal}``, the random state of the training transforms drawn like the reference draws it
    (``preproc.draw_train_state``; centre crop and no augmentation when ``train=False``); the chain itself then runs on
    the device (``device_transforms``). RIGHT knees carry their mirror axis in the state instead of being flipped here."""

    def __init__(self, modals: Sequence[str], sizes: dict, n: int, seed: int = 779, stored: bool = False, train: bool = True,
                 stored_sizes: dict = None):
        self.modals, self.sizes, self.n, self.seed = list(modals), dict(sizes), int(n), int(seed)
        self.stored, self.train = bool(stored), bool(train)
        self.stored_sizes = dict(stored_sizes or {})

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        import random

        import numpy as np

        from . import preproc

        g = torch.Generator().manual_seed(self.seed * 100003 + idx)
        rng = random.Random(self.seed * 100003 + idx)
        side = "RIGHT" if rng.random() < 0.5 else "LEFT"
        item = {("-", "exam_knee_id"): f"{9000000 + idx}__{side}", ("-", "side"): side}
        for m in self.modals:
            if m == "clin":
                vec = encode_clinical(rng.gauss(60.9, 9.2), rng.choice(["MALE", "FEMALE"]), rng.gauss(28.7, 4.9),
                                      int(rng.random() < 0.3), int(rng.random() < 0.1), abs(rng.gauss(10.9, 14.6)))
                item["image__clin"] = vec.unsqueeze(0)
                continue
            if not self.stored:
                item[f"image__{m}"] = torch.randn((1,) + tuple(self.sizes[m]), generator=g)
                continue
            spec = MODAL_SPECS[m]
            size = tuple(self.stored_sizes.get(m, spec["stored"]))
            hi = 256 if spec["dtype"] == torch.uint8 else 4096
            item[f"image__{m}"] = torch.randint(0, hi, (1,) + size, generator=g, dtype=torch.int32).to(spec["dtype"])
            if self.train:
                st = preproc.draw_train_state(rng, size, self.sizes[m], with_gamma=spec["gamma"])
            else:
                st = {"offsets": preproc.crop_offsets(size, self.sizes[m]), "theta": None, "gamma": None}
            st["flip"] = preproc.FLIP_AXIS[m] if side == "RIGHT" else 0
            item[f"state__{m}"] = st
        item["target"] = np.asarray([int(rng.random() < 0.12)])
        return item


def preproc_flip_axis(modal: str) -> int:
    """Axis mirrored for RIGHT knees of this sequence (``preproc.FLIP_AXIS``)."""
    from . import preproc

    return preproc.FLIP_AXIS[modal]


def collate_knees(items: Sequence[dict]) -> dict:
    """Collate for ``stored=True`` items: tensors are stacked like the default collate does, the per-sample transform
    states stay a list of dicts (one per knee), identifiers a list."""
    import numpy as np

    out = {}
    for k in items[0]:
        vals = [it[k] for it in items]
        if isinstance(vals[0], torch.Tensor):
            out[k] = torch.stack(vals)
        elif isinstance(vals[0], np.ndarray):
            out[k] = torch.from_numpy(np.stack(vals))
        else:
            out[k] = vals
    return out


def device_transforms(batch: dict, modals: Sequence[str], sizes: dict, downscale=None, device=None) -> List[torch.Tensor]:
    """Positional model inputs from a ``stored=True`` batch: per image modality ONE ``koa_augment_resample`` call (mirror,
    crop, unit range, rotation, gamma, z-score of ``MODAL_SPECS`` and the downscale factor of the recipe,
    ``run/train_prog_fus.py:143-146``) on the stored integers after their host-to-device copy; the clinical vector passes
    through. ``sizes[modal]`` is the crop (``config.model.input_size`` before the downscale)."""
    from . import preproc

    xs = []
    for i, m in enumerate(modals):
        x = batch[f"image__{m}"]
        if device is not None:
            x = x.to(device, non_blocking=True)
        if m != "clin":
            spec = MODAL_SPECS[m]
            factor = downscale[i] if downscale else None
            x = preproc.augment_normalize_downscale(x, sizes[m], batch[f"state__{m}"], spec["mean"], spec["std"], factor)
        xs.append(x)
    return xs
