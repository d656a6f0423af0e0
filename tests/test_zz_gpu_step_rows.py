"""GPU parity of the rows next to the hot path through the C ABI: ``koa_adam_step``,
``koa_resample_linear``, ``koa_unit_range_affine``, ``koa_augment_resample``, ``koa_predict``, ``koa_ensemble_proba``.

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
            if not orf.state[b]:            # the parameter skipped so far: no state on either side
                assert not om.state[a]
                continue
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
# the loader's transform chain fused with the downscale
# ---------------------------------------------------------------------------------------------------------------------
def test_augment_matches_reference_golden(cuda, gold):
    """Outputs of the unmodified reference transform classes (random state set by hand); two knees per call with different
    states, the second one the validation chain."""
    for case in gold["augment"]:
        x = torch.from_numpy(seeded_volume(case["seed"], tuple(case["stored"]), case["kind"]))
        batch = torch.stack([x, x])[:, None].to(cuda)
        states = [{"offsets": case["offsets"], "theta": case["theta"], "gamma": case["gamma"], "flip": case["flip"]},
                  {"offsets": preproc.crop_offsets(case["stored"], case["crop"]), "theta": None, "gamma": None}]
        y = preproc.augment_normalize_downscale(batch, case["crop"], states, case["mean"], case["std"], case["factor"])
        assert list(y.shape) == [2] + case["out_shape"], case["name"]
        ref = torch.tensor(case["out"], dtype=torch.float32).reshape(case["out_shape"])
        torch.testing.assert_close(y[0].cpu(), ref, rtol=1e-5, atol=5e-5, msg=case["name"])
        plain = so.augment_chain(x.numpy(), states[1]["offsets"], case["crop"], None, None, case["mean"], case["std"],
                                 case["factor"])
        torch.testing.assert_close(y[1].cpu(), torch.from_numpy(plain), rtol=1e-5, atol=5e-5, msg=case["name"])


@pytest.mark.parametrize("stored,crop,dtype,factor", [
    ((40, 36, 10), (32, 33, 7), torch.uint8, (0.5, 0.5, 0.5)),      # odd crop sizes: not a box mean
    ((50, 44), (41, 37), torch.uint16, (0.5, 0.5)),                 # 2-D
    ((24, 24, 5), (20, 22, 4), torch.int16, None),                  # signed storage, no downscale
    ((18, 20, 4), (16, 16, 4), torch.float32, (0.75, 0.4, 1.0)),    # non-dyadic factors
])
def test_augment_matches_the_oracle_on_random_states(cuda, stored, crop, dtype, factor):
    import random

    rng = random.Random(12)
    g = torch.Generator().manual_seed(13)
    b = 5
    if dtype == torch.float32:
        x = torch.randn((b, 1) + stored, generator=g)
    else:
        lo, hi = {torch.uint8: (3, 250), torch.uint16: (100, 4000), torch.int16: (-1500, 2500)}[dtype]
        x = torch.randint(lo, hi, (b, 1) + stored, generator=g, dtype=torch.int32).to(dtype)
    states = [preproc.draw_train_state(rng, stored, crop, rotate_prob=0.75, gamma_prob=0.75) for _ in range(b)]
    states[-1] = {"offsets": preproc.crop_offsets(stored, crop), "theta": 0.26, "gamma": 0.5}    # extreme angle / gamma
    states[1]["flip"] = 1                                    # RIGHT knees: mirrored columns / slices
    states[2]["flip"] = 2 if len(stored) == 3 else 1
    y = preproc.augment_normalize_downscale(x.to(cuda), crop, states, 0.4, 0.25, factor).cpu()
    for k, st in enumerate(states):
        ref = so.augment_chain(x[k, 0].numpy(), st["offsets"], crop, st["theta"], st["gamma"], 0.4, 0.25, factor,
                               st.get("flip") or 0)
        torch.testing.assert_close(y[k], torch.from_numpy(ref), rtol=1e-5, atol=5e-5, msg=f"volume {k}: {st}")


def _torch_chain(vol, st, crop, mean, std, factor):
    """The reference chain written with the torch calls the reference makes (F.affine_grid / F.grid_sample /
    F.interpolate), on the device: the full-size counterpart of the numpy oracle."""
    if st.get("flip"):
        vol = vol.flip(st["flip"])
    sel = tuple(slice(o, o + c) for o, c in zip(st["offsets"], crop))
    v = vol[sel].float()
    v = (v - v.min()) / (v.max() - v.min())
    if st["theta"] is not None:
        th = torch.tensor(st["theta"])
        mat = torch.tensor([[torch.cos(th), -torch.sin(th), 0], [torch.sin(th), torch.cos(th), 0]], device=v.device)
        img = v.permute(2, 0, 1)[:, None] if v.ndim == 3 else v[None, None]          # (S, 1, R, C)
        grid = F.affine_grid(mat[None].repeat(img.shape[0], 1, 1), img.size(), align_corners=False)
        img = F.grid_sample(img, grid, align_corners=False)
        v = img[:, 0].permute(1, 2, 0) if v.ndim == 3 else img[0, 0]
    if st["gamma"] is not None:
        v = v ** (1.0 / st["gamma"])
    v = ((v - mean) / std)[None, None]
    if factor:
        v = F.interpolate(v, scale_factor=tuple(factor), mode="trilinear" if v.ndim == 5 else "bilinear",
                          align_corners=False, recompute_scale_factor=True)
    return v[0]


def test_augment_recipe_sizes(cuda):
    """Full sizes of the reference recipes: DESS 320x320x128 crops of uint8 volumes at (0.5, 0.5, 0.5), XR 700x700 crops of
    uint16 images at (0.5, 0.5). Properties that need no oracle — the chain without rotation and gamma is
    unit_range_normalize_downscale of the crop; rotation by 0 is the identity — and the torch calls of the reference on
    the same device for the rotated / gamma-corrected samples."""
    g = torch.Generator().manual_seed(21)
    stored, crop = (352, 340, 136), (320, 320, 128)
    dess = torch.randint(0, 256, (3, 1) + stored, dtype=torch.uint8, generator=g).to(cuda)
    states = [{"offsets": [7, 13, 5], "theta": None, "gamma": None},
              {"offsets": [32, 20, 8], "theta": 0.0, "gamma": None},
              {"offsets": [0, 0, 0], "theta": -0.21, "gamma": 1.7, "flip": 2}]
    y = preproc.augment_normalize_downscale(dess, crop, states, 0.257, 0.235, (0.5, 0.5, 0.5))
    assert y.shape == (3, 1, 160, 160, 64)
    for k in (0, 1):
        o = states[k]["offsets"]
        cut = dess[k:k + 1, :, o[0]:o[0] + 320, o[1]:o[1] + 320, o[2]:o[2] + 128].contiguous()
        ref = preproc.unit_range_normalize_downscale(cut, 0.257, 0.235, (0.5, 0.5, 0.5))
        # k = 0: same arithmetic up to the folded affine map; k = 1: the identity grid lands on the voxel centres up to
        # one ulp of a coordinate ~320 (3e-5 of a unit-range step of at most 1 / 0.235)
        torch.testing.assert_close(y[k:k + 1], ref, rtol=1e-5, atol=2e-5 if k == 0 else 5e-4)
    # coordinates ~320 carry an ulp of 3e-5, a unit-range difference between neighbours of up to 1, divided by std
    torch.testing.assert_close(y[2], _torch_chain(dess[2, 0], states[2], crop, 0.257, 0.235, (0.5, 0.5, 0.5)),
                               rtol=1e-4, atol=1e-3)
    xr_i32 = torch.randint(0, 4096, (2, 1, 720, 712), dtype=torch.int32, generator=g)
    xr = xr_i32.to(torch.uint16).to(cuda)         # uint16 only ever crosses to the device as bytes: no CUDA cast kernels
    states = [{"offsets": [20, 12], "theta": 0.15, "gamma": 0.6, "flip": 1}, {"offsets": [3, 0], "theta": None, "gamma": 1.9}]
    y = preproc.augment_normalize_downscale(xr, (700, 700), states, 0.543, 0.296, (0.5, 0.5))
    assert y.shape == (2, 1, 350, 350)
    for k in range(2):
        ref = _torch_chain(xr_i32[k, 0].to(cuda), states[k], (700, 700), 0.543, 0.296, (0.5, 0.5))
        torch.testing.assert_close(y[k], ref, rtol=1e-4, atol=1e-3)
    assert torch.isfinite(y).all()


def test_stored_batches_transformed_on_the_device(cuda):
    """The loader contract end to end: stored integer volumes + per-sample transform state from the
    dataset, one koa_augment_resample call per modality on the device, against the numpy oracle knee by knee."""
    from oaprogressionmmf_b200 import synthetic as sy

    modals = ["xr_pa", "sag_3d_dess", "cor_iw_tse", "sag_t2_map", "clin"]
    sizes = {"xr_pa": (20, 20), "sag_3d_dess": (16, 16, 8), "cor_iw_tse": (16, 16, 4), "sag_t2_map": (16, 16, 3)}
    stored = {"xr_pa": (24, 22), "sag_3d_dess": (18, 20, 9), "cor_iw_tse": (16, 18, 4), "sag_t2_map": (17, 16, 4)}
    downscale = [(0.5, 0.5), (0.5, 0.5, 0.5), (0.5, 0.5, 1.0), (0.5, 0.5, 1.0), (1.0,)]
    for train in (True, False):
        ds = sy.SyntheticKneeDataset(modals, sizes, n=8, stored=True, train=train, stored_sizes=stored)
        dl = torch.utils.data.DataLoader(ds, batch_size=4, drop_last=True, collate_fn=sy.collate_knees)
        saw_right = saw_rot = saw_gamma = False
        for batch in dl:
            assert batch["image__sag_3d_dess"].dtype == torch.uint8 and batch["image__xr_pa"].dtype == torch.uint16
            xs = [x.cpu() for x in sy.device_transforms(batch, modals, sizes, downscale, device=cuda)]
            assert [tuple(x.shape) for x in xs] == [(4, 1, 10, 10), (4, 1, 8, 8, 4), (4, 1, 8, 8, 4), (4, 1, 8, 8, 3), (4, 1, 9)]
            for m, x, f in zip(modals[:-1], xs, downscale):
                spec = sy.MODAL_SPECS[m]
                for k, st in enumerate(batch[f"state__{m}"]):
                    side = batch[("-", "side")][k]
                    assert st["flip"] == (sy.preproc_flip_axis(m) if side == "RIGHT" else 0)
                    if not train:
                        assert st["theta"] is None and st["gamma"] is None
                    if m == "sag_t2_map":
                        assert st["gamma"] is None              # no gamma augmentation for T2 maps
                    saw_right |= side == "RIGHT"
                    saw_rot |= st["theta"] is not None
                    saw_gamma |= st["gamma"] is not None
                    raw = batch[f"image__{m}"][k, 0].numpy()
                    ref = so.augment_chain(raw, st["offsets"], sizes[m], st["theta"], st["gamma"], spec["mean"], spec["std"],
                                           f, st["flip"])
                    np.testing.assert_allclose(x[k].numpy(), ref, rtol=1e-5, atol=5e-5, err_msg=f"{m} knee {k}: {st}")
        assert saw_right and (saw_rot and saw_gamma) == train


def test_epoch_loops_on_the_cuda_path(cuda):
    """trainpath.train_epoch / val_epoch (the reference trainer's loops with one read-back per epoch) around the CUDA
    model, FocalLoss and the fused Adam: same losses as a replay that reads loss.item() every step, as the reference does
    (weight gradients are summed with fp32 atomics, so equal up to summation order)."""
    from oaprogressionmmf_b200 import koamodels, synthetic as sy, trainpath
    from oaprogressionmmf_b200.losses import FocalLoss
    from oracle import koa_oracle as ko
    from tests.util import to_attr

    cfg = ko.make_config("XR1Cnn", xr_size=64, xr_arch="resnet18")
    modals, sizes = ["xr_pa"], {"xr_pa": (64, 64)}

    def loader():
        return torch.utils.data.DataLoader(sy.SyntheticKneeDataset(modals, sizes, n=12, seed=3), batch_size=4, drop_last=True)

    def build():
        torch.manual_seed(778)
        model = koamodels.dict_models["XR1Cnn"](to_attr(cfg), None).to(cuda).train()
        # (lr 1e-4, the reference's: with 1e-3 the first Adam update, lr * sign(g) almost everywhere, turns the summation-order
        # noise of near-zero gradient entries into 2 % of the next step's loss)
        return model, koptim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)

    loss_fn = FocalLoss(num_classes=2)
    model, opt = build()
    metrics = trainpath.train_epoch(model, loader(), modals, loss_fn, opt, device=cuda)
    ref_model, ref_opt = build()
    logged = []
    for batch in loader():
        ref_opt.zero_grad()
        x = batch["image__xr_pa"].to(cuda)
        loss = loss_fn(input=ref_model(x)["main"].squeeze(1), target=batch["target"].to(cuda).long().squeeze(1))
        logged.append(loss.item())
        loss.backward()
        ref_opt.step()
    got = metrics["batch-w"]["loss_prog"]
    assert len(got) == 3 and all(np.isfinite(got))
    np.testing.assert_allclose(got, logged, rtol=2e-2, atol=1e-3)
    val = trainpath.val_epoch(model.eval(), loader(), modals, loss_fn, device=cuda)
    assert len(val["batch-w"]["loss_prog"]) == 3
    proba = val["epoch-w"]["predict_proba"]
    assert proba.shape == (12, 2) and val["epoch-w"]["target"].shape == (12, 1)
    np.testing.assert_allclose(proba.sum(1), np.ones(12), rtol=0, atol=1e-5)


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
