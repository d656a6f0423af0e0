"""GPU parity of the rows next to the hot path through the C ABI: ``koa_adam_step``,
``koa_resample_linear``, ``koa_unit_range_affine``, ``koa_predict``, ``koa_ensemble_proba``.

Checked against (a) the golden vectors made from the unmodified reference (``tests/golden_step/step_rows.json``), (b) the
numpy oracle on seeded inputs, (c) ``torch.optim.Adam`` / ``AdamW`` — which *are* the reference's optimisers — on the same
device, and (d) size-independent properties at the sizes of the reference recipes (box-mean equivalence of the 0.5
factors, identity of factor 1, exact minimum / maximum, probabilities summing to one)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oaprogressionmmf_b200 import evalpath, optim as koptim, preproc
from oracle import step_oracle as so
from oracle.make_golden_step import seeded_volume

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# fp32 outputs of magnitude ~1. Where the size ratio is not a power of two the source coordinate scale * (dst + 0.5) - 0.5
# is itself rounded (one ulp of a coordinate ~8 moves an interpolation weight by 1e-6, with or without FMA contraction),
# hence the absolute term; a wrong tap or weight is an error of order 0.1 - 1.
FP32_TOL = dict(rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden_step", "step_rows.json")) as f:
        return json.load(f)


# ---------------------------------------------------------------------------------------------------------------------
# Adam
# ---------------------------------------------------------------------------------------------------------------------
def _adam_params(dev, seed=0):
    """Sizes around the 8192-element chunk and the 4-element vector width, > 64 tensors (several launches), and one
    parameter that is an unaligned view (scalar path of the kernel)."""
    g = torch.Generator().manual_seed(seed)
    sizes = [1, 3, 4, 5, 255, 8191, 8192, 8193, 16384 + 7, 100003, 64 * 3 * 7 * 7] + [64, 128, 256, 512] * 15
    ps = [torch.randn(n, generator=g).to(dev) for n in sizes]
    backing = torch.randn(1001, generator=g).to(dev)
    ps.append(backing[1:])          # 4-byte aligned only
    ps.append(torch.randn(33, 65, generator=g).to(dev))
    return ps


@pytest.mark.parametrize("cls,ref_cls,wd", [(koptim.Adam, torch.optim.Adam, 0.0), (koptim.Adam, torch.optim.Adam, 1e-4),
                                            (koptim.AdamW, torch.optim.AdamW, 1e-2)])
def test_adam_matches_torch_optim(cuda, cls, ref_cls, wd):
    mine = [torch.nn.Parameter(p.clone()) if p.is_contiguous() and p.storage_offset() == 0 else p.requires_grad_()
            for p in _adam_params(cuda)]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    om, orf = cls(mine, lr=1e-3, weight_decay=wd), ref_cls(ref, lr=1e-3, weight_decay=wd, foreach=False)
    g = torch.Generator().manual_seed(1)
    for it in range(4):
        for k, (a, b) in enumerate(zip(mine, ref)):
            if k == 5 and it == 0:
                a.grad = b.grad = None      # skipped once: its step count lags behind
                continue
            gr = (torch.randn(a.shape, generator=g) * (3.0 if it == 2 else 0.1)).to(cuda)
            a.grad, b.grad = gr.clone(), gr.clone()
        om.step()
        orf.step()
        for a, b in zip(mine, ref):
            torch.testing.assert_close(a.detach(), b.detach(), rtol=2e-6, atol=2e-7)
            torch.testing.assert_close(om.state[a]["exp_avg"], orf.state[b]["exp_avg"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(om.state[a]["exp_avg_sq"], orf.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-7)
    assert float(om.state[mine[5]]["step"]) == 3 and float(om.state[mine[0]]["step"]) == 4
    fresh = ref_cls([torch.nn.Parameter(p.detach().clone()) for p in mine], lr=1e-3, weight_decay=wd)
    fresh.load_state_dict(om.state_dict())


def test_adam_matches_the_oracle(cuda):
    g = torch.Generator().manual_seed(2)
    p0, gr = torch.randn(20001, generator=g), torch.randn(20001, generator=g)
    p = torch.nn.Parameter(p0.clone().to(cuda))
    opt = koptim.Adam([p], lr=3e-3, betas=(0.8, 0.99), eps=1e-6, weight_decay=5e-4)
    m = v = np.zeros(20001, np.float32)
    q = p0.numpy()
    for step in (1, 2, 3):
        p.grad = gr.to(cuda) * step
        opt.step()
        q, m, v = so.adam_step(q, gr.numpy() * np.float32(step), m, v, step, lr=3e-3, betas=(0.8, 0.99), eps=1e-6,
                               weight_decay=5e-4)
        np.testing.assert_allclose(p.detach().cpu().numpy(), q, rtol=2e-6, atol=2e-7)


def test_adam_drives_a_model_step(cuda):
    """One optimiser step on the gradients of the CUDA path (XR1Cnn, resnet18) equals torch.optim.Adam on the same
    gradients; the dead-weight case (no gradient) stays untouched."""
    from oaprogressionmmf_b200 import koamodels
    from oaprogressionmmf_b200.losses import FocalLoss
    from oracle import koa_oracle as ko
    from tests.util import to_attr

    cfg = ko.make_config("XR1Cnn", xr_size=64, xr_arch="resnet18")   # the configuration of tests/golden/XR1Cnn_r18.json
    torch.manual_seed(778)
    model = koamodels.dict_models["XR1Cnn"](to_attr(cfg), None).to(cuda).train()
    x = torch.randn(4, 1, 64, 64, device=cuda)
    y = torch.tensor([0, 1, 1, 0], device=cuda)
    loss = FocalLoss(num_classes=2)(model(x)["main"].squeeze(1), y)
    loss.backward()
    live = [p for p in model.parameters() if p.grad is not None]
    assert live
    ref = [torch.nn.Parameter(p.detach().clone()) for p in live]
    for r, p in zip(ref, live):
        r.grad = p.grad.clone()
    koptim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4).step()
    torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-4, foreach=False).step()
    for r, p in zip(ref, live):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=2e-6, atol=2e-7)


# ---------------------------------------------------------------------------------------------------------------------
# resampling / normalisation
# ---------------------------------------------------------------------------------------------------------------------
def test_interpolate_matches_reference_golden(cuda, gold):
    for case in gold["interp"]:
        x = torch.from_numpy(seeded_volume(case["seed"], tuple(case["shape"]), case["kind"])).to(cuda)
        y = preproc.PTInterpolate(tuple(case["factor"]))(x)
        assert list(y.shape) == case["out_shape"], case["name"]
        ref = torch.tensor(case["out"], dtype=torch.float32).reshape(case["out_shape"])
        torch.testing.assert_close(y.cpu(), ref, msg=case["name"], **FP32_TOL)


def test_normalize_downscale_matches_reference_golden(cuda, gold):
    for case in gold["norm"]:
        x = torch.from_numpy(seeded_volume(case["seed"], tuple(case["shape"]), case["kind"])).to(cuda)
        y = preproc.unit_range_normalize_downscale(x, case["mean"], case["std"], case["factor"])
        ref = torch.tensor(case["out"], dtype=torch.float32).reshape(case["out_shape"])
        torch.testing.assert_close(y.cpu(), ref, rtol=1e-5, atol=2e-5, msg=case["name"])


@pytest.mark.parametrize("shape,factor,dtype", [
    ((3, 1, 21, 18, 7), (0.5, 0.5, 0.5), torch.float32),     # odd sizes: not a box mean
    ((2, 2, 30, 31), (0.75, 0.4), torch.float32),
    ((2, 1, 16, 12, 5), (0.5, 0.5, 1.0), torch.int16),
    ((5, 1, 9), (1.0,), torch.float32),
])
def test_interpolate_matches_the_oracle(cuda, shape, factor, dtype):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(shape, generator=g) if dtype == torch.float32 else \
        torch.randint(-2000, 2000, shape, generator=g, dtype=dtype)
    y = preproc.PTInterpolate(factor)(x.to(cuda))
    ref = so.interpolate_linear(x.numpy(), factor)
    torch.testing.assert_close(y.cpu(), torch.from_numpy(ref), rtol=1e-5, atol=1e-5 * float(x.abs().max()))


def test_recipe_sizes_box_mean_identity_and_minmax(cuda):
    """Full sizes of the reference recipes (runner.sh:347-362): DESS 320x320x128 uint8 at factor (0.5, 0.5, 0.5) and the
    XR 700x700 image at (0.5, 0.5) are exact box means; factor 1 along the slice axis is the identity; the per-volume
    minimum / maximum are exact."""
    g = torch.Generator().manual_seed(6)
    dess = torch.randint(0, 256, (2, 1, 320, 320, 128), dtype=torch.uint8, generator=g).to(cuda)
    dess[0, 0, 17, 5, 3] = 0
    dess[1].clamp_(9, 201)
    y = preproc.PTInterpolate((0.5, 0.5, 0.5))(dess)
    assert y.shape == (2, 1, 160, 160, 64)
    torch.testing.assert_close(y, F.avg_pool3d(dess.float(), 2), rtol=1e-6, atol=1e-4)
    _, _, mm = preproc.unit_range_affine(dess, 0.257, 0.235)
    assert mm.cpu().tolist() == [[0.0, 255.0], [9.0, 201.0]]
    z = preproc.unit_range_normalize_downscale(dess, 0.257, 0.235, (0.5, 0.5, 0.5))
    lo = mm[:, 0].view(2, 1, 1, 1, 1)
    hi = mm[:, 1].view(2, 1, 1, 1, 1)
    ref = ((F.avg_pool3d(dess.float(), 2) - lo) / (hi - lo) - 0.257) / 0.235
    torch.testing.assert_close(z, ref, rtol=1e-5, atol=2e-5)
    t2 = torch.randn(2, 1, 320, 320, 25, generator=g).to(cuda)    # odd slice count, factor 1 along it
    y = preproc.downscale_x(t2, [0.5, 0.5, 1.0])
    assert y.shape == (2, 1, 160, 160, 25)
    ref = F.avg_pool3d(t2, (2, 2, 1))
    torch.testing.assert_close(y, ref, rtol=1e-6, atol=1e-6)
    xr = torch.randn(3, 1, 700, 700, generator=g).to(cuda)
    torch.testing.assert_close(preproc.downscale_x(xr, [0.5, 0.5]), F.avg_pool2d(xr, 2), rtol=1e-6, atol=1e-6)
    clin = torch.randn(16, 1, 9, generator=g).to(cuda)
    assert torch.equal(preproc.downscale_x(clin, [1.0]), clin)
    # unaligned volumes (odd element counts, uint16): scalar path of the min / max kernel
    odd = torch.randint(100, 4000, (3, 1, 7, 9, 5), dtype=torch.int32, generator=g).to(torch.uint16).to(cuda)
    _, _, mm = preproc.unit_range_affine(odd, 0.5, 0.25)
    as_i = odd.cpu().to(torch.int32).reshape(3, -1)
    assert mm.cpu().tolist() == [[float(a.min()), float(a.max())] for a in as_i]
    neg = torch.randn(2, 1, 50, 41, generator=g).to(cuda) - 3.0     # negative fp32 values: ordered-integer atomics
    _, _, mm = preproc.unit_range_affine(neg, 0.5, 0.25)
    torch.testing.assert_close(mm, torch.stack([neg.amin((1, 2, 3)), neg.amax((1, 2, 3))], 1), rtol=0, atol=0)


def test_downscaled_input_feeds_the_model(cuda):
    """The row in its place: integer volume -> fused normalise + downscale -> MR1CnnTrf forward equals the model on the
    oracle-preprocessed fp32 volume."""
    from oaprogressionmmf_b200 import koamodels
    from oracle import koa_oracle as ko
    from tests.util import to_attr

    cfg = ko.make_config("MR1CnnTrf", mr_size=64, slices=(4,), depth=2, output_type="main")  # as tests/golden/MR1CnnTrf.json
    torch.manual_seed(778)
    model = koamodels.dict_models["MR1CnnTrf"](to_attr(cfg), None).to(cuda).eval()
    raw = seeded_volume(7, (2, 1, 128, 128, 8), "u8")
    x_ref = torch.from_numpy(so.interpolate_linear(so.unit_range_normalize(raw, 0.257, 0.235), (0.5, 0.5, 0.5))).to(cuda)
    x = preproc.unit_range_normalize_downscale(torch.from_numpy(raw).to(cuda), 0.257, 0.235, (0.5, 0.5, 0.5))
    torch.testing.assert_close(x, x_ref, rtol=1e-5, atol=2e-5)
    with torch.no_grad():
        a, b = model(x), model(x_ref)
    torch.testing.assert_close(a, b, rtol=0, atol=5e-3)  # plumbing check: the two inputs differ by fp32 round-off


# ---------------------------------------------------------------------------------------------------------------------
# predictions
# ---------------------------------------------------------------------------------------------------------------------
def test_predict_and_ensemble(cuda, gold):
    g = torch.Generator().manual_seed(8)
    logits = torch.randn(300, 2, generator=g) * 3
    logits[5] = torch.tensor([1.25, 1.25])           # tie: first maximum, as torch / numpy argmax
    proba, pred = evalpath.predict(logits.to(cuda))
    rp, ra = so.predict(logits.numpy())
    np.testing.assert_allclose(proba.cpu().numpy(), rp, rtol=1e-5, atol=1e-7)
    assert pred.cpu().tolist() == ra.tolist() and int(pred[5]) == 0
    torch.testing.assert_close(proba.sum(1), torch.ones(300, device=cuda), rtol=0, atol=1e-6)
    many = torch.randn(7, 5, generator=g)            # more than two classes
    proba, pred = evalpath.predict(many.to(cuda))
    torch.testing.assert_close(proba.cpu(), torch.softmax(many, 1), rtol=1e-5, atol=1e-7)
    assert pred.cpu().tolist() == many.argmax(1).tolist()

    raw = {int(k): v for k, v in gold["ensemble"]["raw"].items()}
    ref = gold["ensemble"]["out"]
    ens = evalpath.ensemble_eval_foldw(raw, device=cuda)
    assert ens["exam_knee_id"] == ref["exam_knee_id"] and ens["target"] == ref["target"]
    assert ens["predict"] == ref["predict"]
    np.testing.assert_allclose(ens["predict_proba"], ref["predict_proba"], rtol=1e-5, atol=1e-7)
    folds = torch.softmax(torch.randn(5, 64, 3, generator=g), -1)
    out, pred = evalpath.ensemble_proba(folds.to(cuda))
    rp, ra = so.ensemble(folds.numpy())
    np.testing.assert_allclose(out.cpu().numpy(), rp, rtol=1e-5, atol=1e-7)
    assert pred.cpu().tolist() == ra.tolist()
