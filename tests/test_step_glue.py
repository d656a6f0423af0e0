"""CPU test of the Python -> C-ABI glue of the rows next to the hot path (``optim.Adam``, ``preproc``, ``evalpath``).

No GPU here, so the five C entry points are replaced by a host stand-in that reads the *same ctypes arguments* the real
library receives (pointer tables, size arrays, hyper-parameter structs) from host memory and answers with the oracle.
What is under test is the marshalling and the bookkeeping of the host side (step counters, state layout, per-step
grouping, shapes, scheduler interplay) — the kernels themselves are tested on the GPU in ``test_zz_gpu_step_rows.py``.
The stand-in lives in the test only; the product has no CPU path (``test_no_cpu_fallback_on_these_rows``)."""
import contextlib
import ctypes as C

import numpy as np
import pytest
import torch

from oaprogressionmmf_b200 import _lib, evalpath, optim as koptim, preproc
from oracle import step_oracle as so

_NP = {_lib.DT_F32: np.float32, _lib.DT_U8: np.uint8, _lib.DT_U16: np.uint16, _lib.DT_I16: np.int16}


def _view(ptr, n, dtype):
    size = int(n) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * size).from_address(ptr), dtype=dtype)


class HostStandIn:
    """Answers the calls of ``step_ops.cu`` on host pointers with the oracle."""

    def __init__(self):
        self.adam_calls = []

    def koa_adam_step(self, table, n, hyper, stream):
        h = C.cast(hyper, C.POINTER(_lib.AdamHyper)).contents
        arr = C.cast(table, C.POINTER(_lib.AdamTensor * n)).contents
        self.adam_calls.append((n, h.step, h.lr, h.weight_decay, h.decoupled_weight_decay))
        for t in arr:
            if not t.grad:
                continue
            p, g = _view(t.param, t.numel, np.float32), _view(t.grad, t.numel, np.float32)
            m, v = _view(t.exp_avg, t.numel, np.float32), _view(t.exp_avg_sq, t.numel, np.float32)
            scale = h.grad_scale if h.grad_scale != 0 else 1.0
            p2, m2, v2 = so.adam_step(p, g * np.float32(scale), m, v, h.step, lr=h.lr, betas=(h.beta1, h.beta2),
                                      eps=h.eps, weight_decay=h.weight_decay, decoupled=bool(h.decoupled_weight_decay))
            p[:], m[:], v[:] = p2, m2, v2
        return 0

    def koa_resample_linear(self, src, dtype, out, batch, din, dout, scale, shift, stream):
        din, dout = list(din), list(dout)
        x = _view(src, batch * int(np.prod(din)), _NP[dtype]).reshape([batch, 1] + din).astype(np.float32)
        y = x
        for ax in range(3):  # the oracle's separable taps, driven by explicit sizes
            i0, i1, l0, l1 = so._taps(din[ax], dout[ax])
            shape = [1] * 5
            shape[2 + ax] = dout[ax]
            y = (np.take(y, i0, axis=2 + ax) * l0.reshape(shape) + np.take(y, i1, axis=2 + ax) * l1.reshape(shape))
        if scale:
            y = y * _view(scale, batch, np.float32).reshape(batch, 1, 1, 1, 1) + \
                _view(shift, batch, np.float32).reshape(batch, 1, 1, 1, 1)
        _view(out, y.size, np.float32)[:] = y.astype(np.float32).ravel()
        return 0

    def koa_unit_range_affine(self, src, dtype, batch, n_per, mean, std, ws, scale, shift, minmax, stream):
        x = _view(src, batch * n_per, _NP[dtype]).reshape(batch, n_per).astype(np.float32)
        lo, hi = x.min(1), x.max(1)
        f = np.float32
        _view(scale, batch, np.float32)[:] = f(1) / ((hi - lo) * f(std))
        _view(shift, batch, np.float32)[:] = (-lo / (hi - lo) - f(mean)) / f(std)
        if minmax:
            _view(minmax, 2 * batch, np.float32)[:] = np.stack([lo, hi], 1).ravel()
        return 0

    def koa_augment_resample(self, src, dtype, out, params, batch, sdims, cdims, odims, mean, std, ws, stream):
        sdims, cdims, odims = list(sdims), list(cdims), list(odims)
        x = _view(src, batch * int(np.prod(sdims)), _NP[dtype]).reshape([batch] + sdims)
        table = C.cast(params, C.POINTER(_lib.Augment * batch)).contents
        res = []
        for b, t in enumerate(table):
            theta = float(np.arctan2(t.sin_t, t.cos_t)) if t.rotate else None
            gamma = 1.0 / t.inv_gamma if t.inv_gamma != 0 else None
            factor = [o / c for o, c in zip(odims, cdims)]
            res.append(so.augment_chain(x[b], [t.off0, t.off1, t.off2], cdims, theta, gamma, mean, std, factor, t.flip))
        y = np.stack(res)
        assert list(y.shape[2:]) == odims
        _view(out, y.size, np.float32)[:] = y.astype(np.float32).ravel()
        return 0

    def koa_predict(self, logits, proba, pred, b, c, stream):
        p, a = so.predict(_view(logits, b * c, np.float32).reshape(b, c))
        _view(proba, b * c, np.float32)[:] = p.astype(np.float32).ravel()
        _view(pred, b, np.int64)[:] = a
        return 0

    def koa_ensemble_proba(self, proba, out, pred, f, b, c, stream):
        p, a = so.ensemble(_view(proba, f * b * c, np.float32).reshape(f, b, c))
        _view(out, b * c, np.float32)[:] = p.astype(np.float32).ravel()
        _view(pred, b, np.int64)[:] = a
        return 0

    def koa_last_error(self):
        return b""


@pytest.fixture
def host(monkeypatch):
    stub = HostStandIn()
    monkeypatch.setattr(_lib, "load", lambda: stub)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "on_device", lambda device: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "current_stream", lambda: None)
    return stub


@pytest.mark.parametrize("cls,ref_cls,wd", [(koptim.Adam, torch.optim.Adam, 1e-4), (koptim.AdamW, torch.optim.AdamW, 1e-2)])
def test_adam_host_side_tracks_torch_optim(host, cls, ref_cls, wd):
    g = torch.Generator().manual_seed(11)
    shapes = [(5,), (17, 3), (2, 2, 3, 3), (1,)]
    mine = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    om, orf = cls(mine, lr=2e-3, weight_decay=wd), ref_cls(ref, lr=2e-3, weight_decay=wd)
    sm = koptim.CustomWarmupStaticDecayLR(om, epochs_warmup=2, epochs_static=1, epochs_decay=3)
    sr = koptim.CustomWarmupStaticDecayLR(orf, epochs_warmup=2, epochs_static=1, epochs_decay=3)
    for it in range(6):
        for k, (a, b) in enumerate(zip(mine, ref)):
            if k == 3 and it < 2:          # a parameter without gradient is skipped and keeps its own step count
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g)
            a.grad, b.grad = gr.clone(), gr.clone()
        om.step()
        orf.step()
        sm.step()
        sr.step()
        for a, b in zip(mine, ref):
            torch.testing.assert_close(a, b, rtol=2e-6, atol=2e-7)
    # two distinct step counts in one group -> two library calls per step once the late parameter joined
    assert host.adam_calls[0][:2] == (3, 1) and sorted(c[1] for c in host.adam_calls[-2:]) == [4, 6]
    assert host.adam_calls[-1][4] == int(cls is koptim.AdamW)
    assert om.param_groups[0]["lr"] == pytest.approx(orf.param_groups[0]["lr"])
    # optimiser checkpoints interoperate with torch's (same keys, same per-parameter state layout)
    sd_m, sd_r = om.state_dict(), orf.state_dict()
    assert sd_m["state"].keys() == sd_r["state"].keys()
    for k in sd_r["state"]:
        assert sd_m["state"][k].keys() == sd_r["state"][k].keys()
        assert float(sd_m["state"][k]["step"]) == float(sd_r["state"][k]["step"])
        torch.testing.assert_close(sd_m["state"][k]["exp_avg"], sd_r["state"][k]["exp_avg"], rtol=1e-5, atol=1e-6)
    fresh = ref_cls([torch.nn.Parameter(p.detach().clone()) for p in mine], lr=2e-3, weight_decay=wd)
    fresh.load_state_dict(sd_m)  # torch accepts our state as its own


def test_preproc_glue(host):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 1, 8, 6, 4, generator=g)
    y = preproc.PTInterpolate((0.5, 0.5, 1.0))(x)
    assert y.shape == (2, 1, 4, 3, 4)
    np.testing.assert_allclose(y.numpy(), so.interpolate_linear(x.numpy(), (0.5, 0.5, 1.0)), rtol=1e-6, atol=1e-6)
    xr = torch.randn(3, 2, 10, 12, generator=g)
    y = preproc.downscale_x(xr, [0.5, 0.5])
    assert y.shape == (3, 2, 5, 6)
    np.testing.assert_allclose(y.numpy(), so.interpolate_linear(xr.numpy(), (0.5, 0.5)), rtol=1e-6, atol=1e-6)
    clin = torch.randn(4, 1, 9, generator=g)
    torch.testing.assert_close(preproc.downscale_x(clin, (1.0,)), clin)
    vol = torch.randint(5, 250, (2, 1, 8, 6, 4), dtype=torch.uint8, generator=g)
    z = preproc.unit_range_normalize_downscale(vol, 0.257, 0.235, (0.5, 0.5, 0.5))
    ref = so.interpolate_linear(so.unit_range_normalize(vol.numpy(), 0.257, 0.235), (0.5, 0.5, 0.5))
    np.testing.assert_allclose(z.numpy(), ref, rtol=1e-5, atol=2e-5)
    scale, shift, mm = preproc.unit_range_affine(vol, 0.257, 0.235)
    assert mm.shape == (2, 2) and float(mm[0, 0]) == float(vol[0].min()) and float(mm[1, 1]) == float(vol[1].max())
    with pytest.raises(ValueError):
        preproc.PTInterpolate((0.5, 0.5))(x)            # factor / dimension mismatch
    with pytest.raises(NotImplementedError):
        preproc.PTInterpolate((0.5, 0.5, 0.5))(x, mask=x)
    with pytest.raises(_lib.KoaError):
        preproc.PTInterpolate((0.5, 0.5, 0.5))(x.double())


def test_evalpath_glue(host):
    class Toy(torch.nn.Module):
        def forward(self, a, b):
            return {"main": torch.stack([a.flatten(1).mean(1), b.flatten(1).mean(1)], 1).unsqueeze(1)}

    g = torch.Generator().manual_seed(9)
    batches = []
    for i in range(3):
        n = 2 if i < 2 else 1   # ragged last batch (drop_last=False on the test loader)
        batches.append({"image__xr_pa": torch.randn(n, 1, 8, 8, generator=g), "image__clin": torch.randn(n, 1, 9, generator=g),
                        "target": torch.randint(0, 2, (n, 1)), ("-", "exam_knee_id"): [f"k{i}_{j}" for j in range(n)]})
    model = Toy().train()
    acc = evalpath.eval_epoch(model, batches, ["xr_pa", "clin"], downscale=[[0.5, 0.5], [1.0]])
    assert model.training and acc["exam_knee_id"] == ["k0_0", "k0_1", "k1_0", "k1_1", "k2_0"]
    assert len(acc["predict"]) == 5 and len(acc["predict_proba"]) == 5 and acc["target"][0] == batches[0]["target"][0].tolist()
    for row, pred in zip(acc["predict_proba"], acc["predict"]):
        assert abs(sum(row) - 1) < 1e-6 and pred == int(np.argmax(row))
    raw = {f: {"exam_knee_id": acc["exam_knee_id"][::-1] if f else acc["exam_knee_id"],
               "target": acc["target"][::-1] if f else acc["target"],
               "predict": acc["predict"][::-1] if f else acc["predict"],
               "predict_proba": acc["predict_proba"][::-1] if f else acc["predict_proba"]} for f in range(2)}
    ens = evalpath.ensemble_eval_foldw(raw, device="cpu")
    ref = so.ensemble_eval_foldw(raw)
    assert ens["exam_knee_id"] == ref["exam_knee_id"] and ens["predict"] == ref["predict"]
    np.testing.assert_allclose(ens["predict_proba"], ref["predict_proba"], rtol=1e-6)
    assert set(ens) == set(ref)


def test_augment_glue(host):
    import json
    import os

    from oracle.make_golden_step import seeded_volume

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = json.load(open(os.path.join(root, "tests", "golden_step", "step_rows.json")))
    for case in gold["augment"]:
        x = torch.from_numpy(seeded_volume(case["seed"], tuple(case["stored"]), case["kind"]))
        batch = torch.stack([x, x])[:, None]     # two knees with different states: the second one is the plain chain
        states = [{"offsets": case["offsets"], "theta": case["theta"], "gamma": case["gamma"], "flip": case["flip"]},
                  {"offsets": preproc.crop_offsets(case["stored"], case["crop"]), "theta": None, "gamma": None}]
        y = preproc.augment_normalize_downscale(batch, case["crop"], states, case["mean"], case["std"], case["factor"])
        ref = np.asarray(case["out"], dtype=np.float32).reshape(case["out_shape"])
        assert list(y.shape) == [2] + case["out_shape"], case["name"]
        np.testing.assert_allclose(y[0].numpy(), ref, rtol=1e-4, atol=1e-4, err_msg=case["name"])
        plain = so.augment_chain(x.numpy(), states[1]["offsets"], case["crop"], None, None, case["mean"], case["std"],
                                 case["factor"])
        np.testing.assert_allclose(y[1].numpy(), plain, rtol=1e-5, atol=1e-5, err_msg=case["name"])
    with pytest.raises(ValueError):
        preproc.augment_normalize_downscale(batch, case["crop"], [{"offsets": [5, 5, 5], "theta": None, "gamma": None}] * 2,
                                            0.3, 0.2)
    with pytest.raises(ValueError):
        preproc.augment_normalize_downscale(batch, case["crop"], states[:1], 0.3, 0.2)
    with pytest.raises(ValueError):       # no slice axis to mirror in a 2-D image
        xr = torch.zeros(1, 1, 8, 8, dtype=torch.uint8)
        preproc.augment_normalize_downscale(xr, (8, 8), [{"offsets": [0, 0], "theta": None, "gamma": None, "flip": 2}], 0.3, 0.2)
    assert preproc.FLIP_AXIS == {"sag_3d_dess": 2, "cor_iw_tse": 1, "sag_t2_map": 2, "xr_pa": 1}
