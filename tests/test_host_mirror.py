"""Host-side mirror of ``koafusion.models`` (oaprogressionmmf_b200/koamodels): registry, constructor contract,
``state_dict`` layout and error behaviour. CPU only: the modules are parameter holders and their compute lives in
the CUDA library, so nothing here runs a forward pass (except to check that it refuses to run on the CPU).

The key/shape specification used as ground truth (oracle.model_param_spec) is itself pinned against the unmodified
reference by tests/test_oracle_golden.py (``num_keys`` and every gradient key of the golden fixtures)."""
import json
import os

import pytest
import torch

from oaprogressionmmf_b200 import _lib, koamodels
from oaprogressionmmf_b200.synthetic import model_config
from oracle import koa_oracle as ko
from tests.util import to_attr

REFERENCE_NAMES = ["XR1Cnn", "MR1CnnTrf", "MR2CnnTrf", "XR1MR1CnnTrf", "XR1MR2CnnTrf", "XR1MR2C1CnnTrf"]
EXTENSIONS = ["MR3CnnTrf", "XR1MR3C1CnnTrf"]


def test_registry_matches_reference_names():
    # koafusion/models/__init__.py:8-15 (+ the two 3-MRI extensions BASELINE.json asks for)
    assert list(koamodels.dict_models)[:6] == REFERENCE_NAMES
    assert set(koamodels.dict_models) == set(REFERENCE_NAMES + EXTENSIONS)
    for name in ("Transformer", "FeaT", "FeedForward", "Attention", "dict_fes"):
        assert hasattr(koamodels, name)
    assert {"resnet18", "resnet34", "resnet50", "resnext50_32x4d"} <= set(koamodels.dict_fes)


@pytest.mark.parametrize("name", REFERENCE_NAMES + EXTENSIONS)
def test_state_dict_layout_equals_reference(name):
    kw = dict(xr_size=64, mr_size=64, slices=(3, 2, 2), depth=1)
    cfg = ko.make_config(name, **kw)
    spec = ko.model_param_spec(name, cfg)
    model = koamodels.dict_models[name](to_attr(cfg), None)
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    for k, shape in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    # round trip both ways, strict
    osd = ko.make_state_dict(spec, 3)
    model.load_state_dict(osd, strict=True)
    for k, v in model.state_dict().items():
        assert torch.equal(v, osd[k]), k


def test_golden_key_counts(golden_dir):
    for fn in sorted(os.listdir(golden_dir)):
        gold = json.load(open(os.path.join(golden_dir, fn)))
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in gold["config_kwargs"].items()}
        cfg = ko.make_config(gold["model"], **kw)
        model = koamodels.dict_models[gold["model"]](to_attr(cfg), None)
        assert len(model.state_dict()) == gold["num_keys"], fn
        got = {k for k, _ in model.named_parameters()}
        assert got == set(gold["grads"]), fn  # same trainable tensors as the reference


def test_synthetic_config_equals_oracle_config():
    for name in REFERENCE_NAMES + EXTENSIONS:
        assert model_config(name) == ko.make_config(name)


def test_output_type_switch_and_errors():
    cfg = ko.make_config("XR1Cnn", xr_size=64, xr_arch="resnet18")
    cfg["fe"]["arch"] = "vgg16"
    with pytest.raises(ValueError):  # _xr1_cnn.py:21-29: unknown extractor
        koamodels.XR1Cnn(to_attr(cfg), None)
    cfg = ko.make_config("MR1CnnTrf", mr_size=64, slices=(2,), depth=1)
    cfg["fe"]["arch"] = "resnext50_32x4d"
    with pytest.raises(ValueError):  # _mrN_cnn_trf.py:38-43
        koamodels.MR1CnnTrf(to_attr(cfg), None)
    cfg = ko.make_config("MR1CnnTrf", mr_size=64, slices=(2,), depth=1)
    cfg["fe"]["dims_view"] = "zz"
    with pytest.raises(ValueError):  # _mrN_cnn_trf.py:60-71
        koamodels.MR1CnnTrf(to_attr(cfg), None)
    cfg = ko.make_config("XR1MR2C1CnnTrf", xr_size=100, mr_size=64, slices=(2, 2), depth=1)
    with pytest.raises(AssertionError):  # _xrNmrMcP.py:104-109: spatial-size lookup
        koamodels.XR1MR2C1CnnTrf(to_attr(cfg), None)
    with pytest.raises(RuntimeError):  # pretrained weights need the reference's download
        koamodels.dict_fes["resnet50"](pretrained=True)


def test_feat_signature_and_limits():
    f = koamodels.FeaT(num_patches=5, patch_dim=256, emb_dim=256, depth=1, heads=8, mlp_dim=256, num_classes=2)
    keys = list(f.state_dict())
    assert keys[:4] == ["cls_token", "pos_embedding", "patch_to_embedding.weight", "patch_to_embedding.bias"]
    assert "transformer.attn_0.to_qkv.weight" in keys and "transformer.attn_0.to_qkv.bias" not in keys
    assert f.pos_embedding.shape == (1, 6, 256)
    f2 = koamodels.FeaT(num_patches=5, patch_dim=256, emb_dim=256, depth=1, heads=8, mlp_dim=256, num_classes=2,
                        with_cls=False)
    assert "cls_token" not in f2.state_dict() and f2.pos_embedding.shape == (1, 5, 256)
    with pytest.raises(ValueError):
        f(torch.zeros(1, 5, 256), mask=torch.ones(1, 5))


def test_cpu_forward_fails_loudly_instead_of_falling_back():
    cfg = ko.make_config("XR1Cnn", xr_size=64, xr_arch="resnet18")
    model = koamodels.XR1Cnn(to_attr(cfg), None)
    with pytest.raises((_lib.KoaError, RuntimeError)):
        model(torch.zeros(1, 1, 64, 64))
    enc = koamodels.SliceEncoder(koamodels.dict_fes["resnet18"](), with_gap=True)
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 2, 64, 64))


def test_multi_device_dataparallel_is_refused_with_guidance():
    """The reference wraps its model in single-process nn.DataParallel (run/train_prog_fus.py:84). Over one device that
    wrapper never replicates and works; over several devices this path is one process per GPU (dataparallel.wrap), and the
    replication hook says so instead of failing inside a kernel."""
    import pytest
    from oaprogressionmmf_b200.koamodels import dict_models
    from oracle import koa_oracle as ko
    from tests.util import to_attr

    model = dict_models["XR1Cnn"](to_attr(ko.make_config("XR1Cnn", xr_size=64, xr_arch="resnet18")), None)
    with pytest.raises(TypeError, match="one process per GPU"):
        model._replicate_for_data_parallel()


def test_backward_stage_bounds_partition_the_flat_gradient_buffer():
    """Host logic of the staged backward pass (koamodels/_fe.py::_stage_bounds, used by the data-parallel wrapper): the four
    stages (layer4, layer3, layer2, layer1 + stem) cover every block once, in backward order, and their flat-buffer
    slices cover every gradient tensor exactly once without overlapping."""
    import torch
    from oaprogressionmmf_b200 import _lib
    from oaprogressionmmf_b200.koamodels import SliceEncoder, dict_fes

    for arch, layers in (("resnet50", (3, 4, 6, 3)), ("resnext50_32x4d", (3, 4, 6, 3)), ("resnet18", (2, 2, 2, 2))):
        enc = SliceEncoder(dict_fes[arch](pretrained=False))
        params = enc._trainable()
        grads, flat = _lib.zeros_like_flat(params)
        stages = enc._stage_bounds(params, grads, flat)
        assert [(b, e, s) for b, e, s, *_ in stages] == [
            (sum(layers[:3]), sum(layers), 0), (sum(layers[:2]), sum(layers[:3]), 0), (layers[0], sum(layers[:2]), 0),
            (0, layers[0], 1)]
        covered = torch.zeros(flat.numel(), dtype=torch.int32)
        seen = []
        prev_lo = flat.numel()
        for _, _, _, lo, hi, ps in stages:
            assert 0 <= lo < hi <= prev_lo      # earlier layers sit earlier in the buffer: stages walk it backwards
            prev_lo = lo
            covered[lo:hi] += 1
            seen += ps
        assert int(covered.max()) == 1
        assert len(seen) == len(params) and {id(p) for p in seen} == {id(p) for p in params}
        for g in grads:
            off = (g.data_ptr() - flat.data_ptr()) // 4
            assert bool((covered[off:off + g.numel()] == 1).all())
        # a frozen tensor simply drops out of its stage
        params[5].requires_grad_(False)
        grads2, flat2 = _lib.zeros_like_flat([p if p.requires_grad else None for p in params])
        stages2 = enc._stage_bounds(params, grads2, flat2)
        assert sum(len(ps) for *_, ps in stages2) == len(params) - 1


def test_fp16_range_guard_of_the_slice_encoder():
    """`poison_if_out_of_fp16_range` (pure torch, device-side in production): features pass through untouched while
    max|x| * max_c sum|w_c| is below the fp16 range, become NaN beyond it and for non-finite inputs, and stay differentiable."""
    from oaprogressionmmf_b200.koamodels._fe import poison_if_out_of_fp16_range

    torch.manual_seed(0)
    w = torch.randn(64, 3, 7, 7) * 0.025
    l1 = float(w.abs().sum(dim=(1, 2, 3)).max())
    feat = torch.randn(6, 2048, requires_grad=True)
    x = torch.randn(6, 32, 32)
    out = poison_if_out_of_fp16_range(x, w, feat)
    assert torch.equal(out, feat)
    out.sum().backward()
    assert torch.equal(feat.grad, torch.ones_like(feat))
    just_below = x / x.abs().max() * (65504.0 / l1) * 0.999
    assert torch.equal(poison_if_out_of_fp16_range(just_below, w, feat), feat)
    assert torch.isnan(poison_if_out_of_fp16_range(just_below * 1.01, w, feat)).all()
    assert torch.isnan(poison_if_out_of_fp16_range(-just_below * 1.01, w, feat)).all()
    bad = x.clone()
    bad[0, 0, 0] = float("nan")
    assert torch.isnan(poison_if_out_of_fp16_range(bad, w, feat)).all()
    bad[0, 0, 0] = float("inf")
    assert torch.isnan(poison_if_out_of_fp16_range(bad, w, feat)).all()
    assert torch.equal(poison_if_out_of_fp16_range(torch.rand(2, 8, 8) * 255.0, w, feat), feat)  # raw 8-bit data
