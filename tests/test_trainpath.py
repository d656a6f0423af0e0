"""``trainpath.train_epoch`` / ``val_epoch`` against a transcription-free replay of the reference loop's per-batch order
(``koafusion/run/train_prog_fus.py:119-240``): same optimiser trajectory, same logged losses, one host read-back per
epoch. CPU: a small torch model stands in for the CUDA model classes (the loops only need ``model(*xs)["main"]``), the
softmax of the validation loop is answered by the oracle stand-in of ``tests/test_step_glue.py``."""
import copy
import functools

import numpy as np
import torch
from torch import nn

from oaprogressionmmf_b200 import synthetic as sy, trainpath
from tests.test_step_glue import host  # noqa: F401  (fixture: host stand-in for the C entry points)

MODALS = ["xr_pa", "sag_t2_map", "clin"]
SIZES = {"xr_pa": (8, 8), "sag_t2_map": (6, 6, 3)}


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b, self.c = nn.Linear(64, 4), nn.Linear(108, 4), nn.Linear(9, 4)
        self.head = nn.Linear(4, 2)

    def forward(self, xr, t2, clin):
        h = self.a(xr.flatten(1)) + self.b(t2.flatten(1)) + self.c(clin.flatten(1))
        return {"main": self.head(torch.relu(h))}          # (B, classes), as the model classes return it


class CE(nn.Module):
    def forward(self, input, target):
        return nn.functional.cross_entropy(input, target)


def _loader():
    ds = sy.SyntheticKneeDataset(MODALS, SIZES, n=10, seed=5)
    return torch.utils.data.DataLoader(ds, batch_size=4, drop_last=True)


def test_train_epoch_follows_the_reference_order():
    torch.manual_seed(0)
    mine = Tiny().train()
    ref = copy.deepcopy(mine)
    om, orf = torch.optim.Adam(mine.parameters(), lr=1e-2), torch.optim.Adam(ref.parameters(), lr=1e-2)
    metrics = trainpath.train_epoch(mine, _loader(), MODALS, CE(), om)
    logged = []
    for batch in _loader():                                 # run/train_prog_fus.py:132-166, statement by statement
        orf.zero_grad()
        xs = tuple(batch[f"image__{m}"] for m in MODALS)
        ys = batch["target"]
        pred = ref(*xs)["main"]
        loss = CE()(input=pred.squeeze(1), target=ys.long().squeeze(1))
        logged.append(loss.item())
        loss.backward()
        orf.step()
    assert set(metrics) == {"batch-w", "epoch-w"} and len(metrics["batch-w"]["loss_prog"]) == 2
    np.testing.assert_allclose(metrics["batch-w"]["loss_prog"], logged, rtol=1e-6)
    for a, b in zip(mine.parameters(), ref.parameters()):
        assert torch.equal(a, b)
    assert np.isfinite(np.mean(np.asarray(metrics["batch-w"]["loss_prog"])))      # what fit() does with the list (:257-259)


def test_val_epoch_accumulates_like_the_reference(host):  # noqa: F811
    torch.manual_seed(1)
    model = Tiny().eval()
    before = copy.deepcopy(model.state_dict())
    seen = {}

    def metrics_fn(prog_target, prog_pred_proba):
        seen["target"], seen["proba"] = prog_target, prog_pred_proba
        return {"n": len(prog_target)}

    metrics = trainpath.val_epoch(model, _loader(), MODALS, CE(), metrics_fn=metrics_fn)
    losses, targets, probas = [], [], []
    with torch.no_grad():
        for batch in _loader():                             # run/train_prog_fus.py:186-225
            xs = tuple(batch[f"image__{m}"] for m in MODALS)
            pred = model(*xs)["main"]
            loss = CE()(input=pred.squeeze(1), target=batch["target"].long().squeeze(1))
            losses.append(np.round(loss.item(), 3))
            targets.append(batch["target"].numpy())
            probas.append(torch.softmax(pred, dim=1).numpy())
    assert metrics["epoch-w"] == {"n": 8}
    np.testing.assert_allclose(metrics["batch-w"]["loss_prog"], losses, atol=1e-9)
    np.testing.assert_array_equal(seen["target"], np.concatenate(targets, axis=0))
    np.testing.assert_allclose(seen["proba"], np.concatenate(probas, axis=0), rtol=1e-5, atol=1e-7)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k])                    # no_grad, no optimiser: nothing moved


def test_transforms_hook_replaces_extract_and_downscale():
    calls = []

    def transforms(batch):
        calls.append(len(batch["target"]))
        return tuple(batch[f"image__{m}"] for m in MODALS)

    torch.manual_seed(2)
    model = Tiny().train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    m = trainpath.train_epoch(model, _loader(), MODALS, CE(), opt, transforms=transforms)
    assert calls == [4, 4] and len(m["batch-w"]["loss_prog"]) == 2
    assert functools.partial(sy.device_transforms, modals=MODALS, sizes=SIZES) is not None
