"""CPU tests of the rows next to the hot path: the oracle restatements against the golden vectors made
from the unmodified reference (``oracle/make_golden_step.py`` -> ``tests/golden_step/step_rows.json``) and against
``torch.optim`` itself, plus the host logic of the product side (sizes, schedules, fold alignment, loud failure without a
GPU). The CUDA kernels are compared with the same oracle in ``tests/test_zz_gpu_step_rows.py`` (``-m gpu``)."""
import json
import os

import numpy as np
import pytest
import torch

from oaprogressionmmf_b200 import _lib, evalpath, optim as koptim, preproc
from oracle import step_oracle as so
from oracle.make_golden_step import seeded_volume

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden_step", "step_rows.json")) as f:
        return json.load(f)


def test_interpolate_oracle_matches_reference(gold):
    assert len(gold["interp"]) >= 6
    for case in gold["interp"]:
        x = seeded_volume(case["seed"], tuple(case["shape"]), case["kind"])
        y = so.interpolate_linear(x, case["factor"])
        assert list(y.shape) == case["out_shape"], case["name"]
        ref = np.asarray(case["out"], dtype=np.float32).reshape(y.shape)
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=2e-6, err_msg=case["name"])
        assert preproc.output_size(case["shape"][2:], case["factor"]) == case["out_shape"][2:]


def test_half_factors_are_box_means(gold):
    """The factors of the reference recipes on even sizes: 2x2(x2) box mean."""
    x = seeded_volume(1, (2, 1, 8, 6, 4), "f32")
    y = so.interpolate_linear(x, (0.5, 0.5, 0.5))
    box = x.reshape(2, 1, 4, 2, 3, 2, 2, 2).mean(axis=(3, 5, 7))
    np.testing.assert_allclose(y, box, rtol=1e-6, atol=1e-6)
    y = so.interpolate_linear(x, (0.5, 0.5, 1.0))
    np.testing.assert_allclose(y, x.reshape(2, 1, 4, 2, 3, 2, 4).mean(axis=(3, 5)), rtol=1e-6, atol=1e-6)


def test_unit_range_normalize_oracle_matches_reference(gold):
    for case in gold["norm"]:
        x = seeded_volume(case["seed"], tuple(case["shape"]), case["kind"])
        y = so.unit_range_normalize(x, case["mean"], case["std"])
        if case["factor"]:
            y = so.interpolate_linear(y, case["factor"])
        ref = np.asarray(case["out"], dtype=np.float32).reshape(case["out_shape"])
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5, err_msg=case["name"])


def test_affine_form_of_the_normalisation_commutes_with_resampling(gold):
    """What the fused kernel computes: scale * interpolate(raw integers) + shift == the reference order of operations."""
    for case in gold["norm"]:
        x = seeded_volume(case["seed"], tuple(case["shape"]), case["kind"]).astype(np.float32)
        f = np.float32
        y = so.interpolate_linear(x, case["factor"]) if case["factor"] else x
        out = np.empty_like(y)
        for b in range(x.shape[0]):
            lo, hi = x[b].min(), x[b].max()
            scale = f(1) / ((hi - lo) * f(case["std"]))
            shift = (-lo / (hi - lo) - f(case["mean"])) / f(case["std"])
            out[b] = y[b] * scale + shift
        ref = np.asarray(case["out"], dtype=np.float32).reshape(case["out_shape"])
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=2e-5, err_msg=case["name"])


def test_augment_chain_oracle_matches_reference(gold):
    """crop -> unit range -> rotation -> gamma -> z-score -> downscale against the reference transform classes run with
    the same (hand-set) random state; every stage switched on and off, 2-D and 3-D, integer and fp32 storage."""
    assert {c["name"] for c in gold["augment"]} >= {"dess_all", "tse_rot_only", "t2_gamma_only", "xr_2d", "val_center",
                                                        "dess_right", "tse_right", "xr_right", "t2_right_val"}
    for case in gold["augment"]:
        x = seeded_volume(case["seed"], tuple(case["stored"]), case["kind"])
        y = so.augment_chain(x, case["offsets"], case["crop"], case["theta"], case["gamma"], case["mean"], case["std"],
                             case["factor"], case["flip"])
        ref = np.asarray(case["out"], dtype=np.float32).reshape(case["out_shape"])
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5, err_msg=case["name"])
    # offsets as the reference computes them
    assert preproc.crop_offsets((14, 12, 6), (12, 10, 4), (0.3, 0.9, 0.5)) == gold["augment"][0]["offsets"]
    assert preproc.crop_offsets((12, 12, 4), (8, 10, 4)) == [2, 1, 0]
    with pytest.raises(ValueError):
        preproc.crop_offsets((4, 4), (5, 4))


def test_train_state_draws_follow_the_reference_order():
    """Same consumption of the ``random`` stream as the reference transforms: ratio per crop axis, (p, theta), (p, gamma)."""
    import math
    import random

    a, b = random.Random(5), random.Random(5)
    st = preproc.draw_train_state(a, (20, 20, 8), (16, 16, 8))
    ratios = [b.random() for _ in range(3)]
    p_rot, theta = b.random(), b.uniform(math.radians(-15.0), math.radians(15.0))
    p_gam, gamma = b.random(), b.uniform(0.5, 2.0)
    assert st["offsets"] == [int(math.floor(r * d)) for r, d in zip(ratios, (4, 4, 0))]
    assert st["theta"] == (theta if p_rot < 0.5 else None) and st["gamma"] == (gamma if p_gam < 0.5 else None)
    assert a.random() == b.random()
    st = preproc.draw_train_state(a, (20, 20, 8), (16, 16, 8), with_gamma=False)
    assert st["gamma"] is None


@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (False, 1e-4), (True, 1e-2)])
def test_adam_oracle_matches_torch_optim(decoupled, wd):
    """The reference's optimiser IS torch.optim.Adam (dict_optimizers): five steps on the same gradients."""
    g = torch.Generator().manual_seed(5)
    shapes = [(7,), (33, 5), (4, 3, 3, 3)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    ref = cls(params, lr=1e-3, weight_decay=wd)
    mine = [(p.detach().numpy().copy(), np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)) for p in params]
    for step in range(1, 6):
        grads = [torch.randn(s, generator=g) * (0.1 if step % 2 else 3.0) for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        ref.step()
        mine = [so.adam_step(p, gr.numpy(), m, v, step, lr=1e-3, weight_decay=wd, decoupled=decoupled)
                for (p, m, v), gr in zip(mine, grads)]
        for (p, m, v), q in zip(mine, params):
            # fp32 round-off of operands of magnitude <= 10 (the moments cancel): a few ulps of the operands
            np.testing.assert_allclose(p, q.detach().numpy(), rtol=2e-6, atol=2e-7)
            np.testing.assert_allclose(m, ref.state[q]["exp_avg"].numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(v, ref.state[q]["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-7)


def test_lr_lambdas_match_reference(gold):
    dummy = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    a, b = gold["lr"]["static_decay"], gold["lr"]["multistep"]
    sa = koptim.CustomWarmupStaticDecayLR(dummy, **a["params"])
    sb = koptim.CustomWarmupMultiStepLR(dummy, **b["params"])
    for e in range(40):
        assert sa.lr_lambdas[0](e) == pytest.approx(a["factors"][e], rel=1e-12)
        assert sb.lr_lambdas[0](e) == pytest.approx(b["factors"][e], rel=1e-12)
        pa = {k: v for k, v in a["params"].items() if k != "epochs_decay"}
        assert so.lr_lambda_warmup_static_decay(e, **pa) == pytest.approx(a["factors"][e], rel=1e-12)
        assert so.lr_lambda_warmup_multistep(e, **b["params"]) == pytest.approx(b["factors"][e], rel=1e-12)
    assert set(koptim.dict_optimizers) == {"SGD", "Adam", "AdamW", "RMSprop"}
    assert len(koptim.dict_schedulers) == 12


def test_ensemble_oracle_and_fold_alignment_match_reference(gold):
    raw = {int(k): v for k, v in gold["ensemble"]["raw"].items()}
    ref = gold["ensemble"]["out"]
    got = so.ensemble_eval_foldw(raw)
    assert got["exam_knee_id"] == ref["exam_knee_id"] and got["target"] == ref["target"]
    assert got["predict"] == ref["predict"]
    np.testing.assert_allclose(got["predict_proba"], ref["predict_proba"], rtol=1e-12)
    # host half of the product: the same join, without touching the device
    ids, target, rows = evalpath.align_folds(raw)
    assert ids == ref["exam_knee_id"] and target == ref["target"]
    for f, idx in rows.items():
        assert [raw[f]["exam_knee_id"][i] for i in idx] == ids
    # a knee missing from one fold drops out (inner join); a duplicated id is an error (validate="1:1")
    short = {k: {c: v[:-1] if k == 1 else v for c, v in d.items()} for k, d in raw.items()}
    ids2, _, _ = evalpath.align_folds(short)
    assert set(ids2) == set(raw[1]["exam_knee_id"][:-1]) and len(ids2) == len(ids) - 1
    dup = {k: dict(d) for k, d in raw.items()}
    dup[2]["exam_knee_id"] = [dup[2]["exam_knee_id"][0]] * len(ids)
    with pytest.raises(ValueError):
        evalpath.align_folds(dup)


def test_predict_oracle():
    logits = np.array([[0.2, -1.0], [3.0, 3.5], [1.0, 1.0]], dtype=np.float32)
    proba, pred = so.predict(logits)
    np.testing.assert_allclose(proba, torch.softmax(torch.from_numpy(logits), 1).numpy(), rtol=1e-6)
    assert pred.tolist() == [0, 1, 0]


def test_no_cpu_fallback_on_these_rows():
    """Without a GPU the product must fail loudly, not compute on the host."""
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    opt = koptim.Adam([p], lr=1e-3)
    assert set(opt.param_groups[0]) >= {"lr", "betas", "eps", "weight_decay", "amsgrad"}
    with pytest.raises(_lib.KoaError):
        opt.step()
    with pytest.raises(_lib.KoaError):
        preproc.PTInterpolate((0.5, 0.5))(torch.zeros(1, 1, 4, 4))
    with pytest.raises(_lib.KoaError):
        preproc.unit_range_normalize_downscale(torch.zeros(1, 1, 4, 4, dtype=torch.uint8), 0.5, 0.2, (0.5, 0.5))
    with pytest.raises(_lib.KoaError):
        evalpath.predict(torch.zeros(2, 2))
    with pytest.raises(ValueError):
        koptim.Adam([p], amsgrad=True)
    with pytest.raises(ValueError):
        koptim.Adam([p], lr=-1.0)
    assert preproc.downscale_x(torch.zeros(2, 1, 9), None).shape == (2, 1, 9)  # empty factor: identity, no launch
