"""The per-element arithmetic of the resampling / augmentation / Adam kernels, checked on the CPU.

``csrc/step_arith.cuh`` holds the ``__host__ __device__`` functions the kernels of ``step_ops.cu`` call for one output
element. ``tests/host_emul/step_emul.cpp`` compiles that same header with g++ and loops over the elements the way the
grid-stride loops do; here its output is compared with the golden vectors of the unmodified reference
(``tests/golden_step/step_rows.json``) and with the numpy oracle on seeded inputs. What this proves without a GPU:
indexing (crop origin, strides, 2-D as one-slice volumes), tap order, weights, the rotation grid, zero padding. What is
left to the ``-m gpu`` tests: the launch geometry, the device min / max reduction, device ``powf`` round-off.

Test infrastructure only — the package never loads this harness (``test_abi.py::test_product_never_imports_the_oracle``
style check below)."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest

from oaprogressionmmf_b200 import _lib, preproc
from oracle import step_oracle as so
from oracle.make_golden_step import seeded_volume

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DT = {"f32": _lib.DT_F32, "u8": _lib.DT_U8, "u16": _lib.DT_U16, "i16": _lib.DT_I16}
_I3 = C.c_int * 3


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_emul") / "step_emul.so")
    src = os.path.join(ROOT, "tests", "host_emul", "step_emul.cpp")
    # -ffp-contract=off: plain IEEE fp32, one rounding per operation (the device contracts a*b+c into FMAs; the
    # tolerances below cover that difference)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.emul_resample_linear.restype = lib.emul_augment_resample.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden_step", "step_rows.json")) as f:
        return json.load(f)


def _pad3(dims):
    return list(dims) + [1] * (3 - len(dims))


def _resample(lib, x, kind, factor, scale=None, shift=None):
    x = np.ascontiguousarray(x)
    b = x.shape[0] * x.shape[1]
    size_out = preproc.output_size(x.shape[2:], tuple(factor)) if factor else list(x.shape[2:])
    out = np.empty((x.shape[0], x.shape[1]) + tuple(size_out), np.float32)
    sc = scale.ctypes.data if scale is not None else None
    sh = shift.ctypes.data if shift is not None else None
    rc = lib.emul_resample_linear(C.c_void_p(x.ctypes.data), _DT[kind], C.c_void_p(out.ctypes.data), b,
                                  _I3(*_pad3(x.shape[2:])), _I3(*_pad3(size_out)), C.c_void_p(sc), C.c_void_p(sh))
    assert rc == 0
    return out


def _augment(lib, x, kind, crop, states, mean, std, factor):
    """x: (B, R, C[, S]) stored volumes; the same marshalling as preproc.augment_normalize_downscale."""
    x = np.ascontiguousarray(x)
    b = x.shape[0]
    size_out = preproc.output_size(crop, tuple(factor)) if factor else list(crop)
    table = (_lib.Augment * b)()
    for t, st in zip(table, states):
        t.off0, t.off1, t.off2 = _pad3(st["offsets"])[:len(st["offsets"])] + [0] * (3 - len(st["offsets"]))
        th, gm = st.get("theta"), st.get("gamma")
        t.rotate = int(th is not None)
        t.cos_t, t.sin_t = (math.cos(th), math.sin(th)) if th is not None else (1.0, 0.0)
        t.inv_gamma = 1.0 / gm if gm is not None else 0.0
        t.flip = st.get("flip") or 0
    out = np.empty((b, 1) + tuple(size_out), np.float32)
    rc = lib.emul_augment_resample(C.c_void_p(x.ctypes.data), _DT[kind], C.c_void_p(out.ctypes.data), table, b,
                                   _I3(*_pad3(x.shape[1:])), _I3(*_pad3(crop)), _I3(*_pad3(size_out)), C.c_float(mean),
                                   C.c_float(std))
    assert rc == 0
    return out, table


def test_resample_arithmetic_matches_reference_golden(emul, gold):
    for case in gold["interp"]:
        x = seeded_volume(case["seed"], tuple(case["shape"]), case["kind"])
        y = _resample(emul, x, case["kind"], case["factor"])
        ref = np.asarray(case["out"], np.float32).reshape(case["out_shape"])
        assert list(y.shape) == case["out_shape"], case["name"]
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5, err_msg=case["name"])


def test_normalize_downscale_arithmetic_matches_reference_golden(emul, gold):
    """Unit range + z-score folded into the resampling as a per-volume affine map (what koa_unit_range_affine feeds)."""
    for case in gold["norm"]:
        x = seeded_volume(case["seed"], tuple(case["shape"]), case["kind"])
        flat = x.reshape(x.shape[0], -1).astype(np.float32)
        lo, hi = flat.min(1), flat.max(1)
        rng = (hi - lo).astype(np.float32)
        scale = (np.float32(1) / (rng * np.float32(case["std"]))).astype(np.float32)
        shift = ((-lo / rng - np.float32(case["mean"])) / np.float32(case["std"])).astype(np.float32)
        y = _resample(emul, x, case["kind"], case["factor"], scale, shift)
        ref = np.asarray(case["out"], np.float32).reshape(case["out_shape"])
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=2e-5, err_msg=case["name"])


def test_augment_arithmetic_matches_reference_golden(emul, gold):
    for case in gold["augment"]:
        x = seeded_volume(case["seed"], tuple(case["stored"]), case["kind"])
        states = [{"offsets": case["offsets"], "theta": case["theta"], "gamma": case["gamma"], "flip": case["flip"]},
                  {"offsets": preproc.crop_offsets(case["stored"], case["crop"]), "theta": None, "gamma": None}]
        y, table = _augment(emul, np.stack([x, x]), case["kind"], case["crop"], states, case["mean"], case["std"],
                            case["factor"])
        ref = np.asarray(case["out"], np.float32).reshape(case["out_shape"])
        assert list(y.shape[1:]) == case["out_shape"], case["name"]
        np.testing.assert_allclose(y[0], ref, rtol=1e-5, atol=2e-5, err_msg=case["name"])
        plain = so.augment_chain(x, states[1]["offsets"], case["crop"], None, None, case["mean"], case["std"], case["factor"])
        np.testing.assert_allclose(y[1], plain, rtol=1e-5, atol=2e-5, err_msg=case["name"])
        sel = tuple(slice(o, o + c) for o, c in zip(case["offsets"], case["crop"]))
        xm = np.flip(x, axis=case["flip"]) if case["flip"] else x
        assert table[0].lo == float(xm[sel].min()) and table[0].range == float(xm[sel].max()) - float(xm[sel].min())


@pytest.mark.parametrize("stored,crop,kind,factor", [
    ((40, 36, 10), (32, 33, 7), "u8", (0.5, 0.5, 0.5)),      # odd crop sizes: not a box mean
    ((31, 29, 6), (31, 29, 6), "u16", (0.5, 0.5, 1.0)),      # crop = stored volume
    ((50, 44), (41, 37), "u16", (0.5, 0.5)),                 # 2-D
    ((24, 24, 5), (20, 22, 4), "i16", None),                 # no downscale
    ((18, 20, 4), (16, 16, 4), "f32", (0.75, 0.4, 1.0)),     # non-dyadic factors
])
def test_augment_arithmetic_matches_the_oracle_on_random_states(emul, stored, crop, kind, factor):
    import random

    rng = random.Random(hash((stored, crop)) & 0xffff)
    if kind == "i16":    # signed storage (CT-like offsets): not a kind of seeded_volume
        vols = np.random.default_rng(77).integers(-1500, 2500, size=(4,) + stored, dtype=np.int16)
    else:
        vols = np.stack([seeded_volume(900 + k, stored, kind) for k in range(4)])
    states = [preproc.draw_train_state(rng, stored, crop, rotate_prob=0.75, gamma_prob=0.75) for _ in range(4)]
    states[3] = {"offsets": preproc.crop_offsets(stored, crop), "theta": 0.26, "gamma": 0.5}   # extreme angle / gamma
    states[1]["flip"] = 1                                   # RIGHT knees: mirrored columns / slices
    states[2]["flip"] = 2 if len(stored) == 3 else 1
    y, _ = _augment(emul, vols, kind, crop, states, 0.4, 0.25, factor)
    for k, st in enumerate(states):
        ref = so.augment_chain(vols[k], st["offsets"], crop, st["theta"], st["gamma"], 0.4, 0.25, factor, st.get("flip") or 0)
        np.testing.assert_allclose(y[k], ref, rtol=1e-5, atol=5e-5, err_msg=f"volume {k}: {st}")


def test_harness_is_test_only():
    """Neither the package nor the built library knows the harness."""
    pkg = os.path.join(ROOT, "oaprogressionmmf_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            assert "host_emul" not in open(os.path.join(pkg, name)).read(), name
    with open(os.path.join(pkg, "libkoa_b200.so"), "rb") as f:
        assert b"emul_augment_resample" not in f.read()


def test_gpu_tests_of_the_step_rows_dry_run_on_the_host(emul, gold, monkeypatch):
    """The ``-m gpu`` tests of ``koa_augment_resample`` (golden, oracle on random states, recipe sizes against the torch
    calls of the reference) executed here with the C entry point answered by the host build of the kernel arithmetic:
    their tolerances and their torch-side reference chain are exercised at full size before they ever see a GPU."""
    import contextlib

    import torch

    import tests.test_step_glue as glue
    import tests.test_zz_gpu_step_rows as gpu_tests

    _I, _P, _F = C.c_int, C.c_void_p, C.c_float
    emul.emul_augment_resample.argtypes = [_P, _I, _P, _P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _F, _F]
    emul.emul_resample_linear.argtypes = [_P, _I, _P, _I, C.POINTER(_I), C.POINTER(_I), _P, _P]
    emul.emul_adam_step.argtypes = [_P, _I, _P]
    emul.emul_predict.argtypes = [_P, _P, _P, _I, _I]
    emul.emul_unit_range_affine.argtypes = [_P, _I, _I, C.c_longlong, _F, _F, _P, _P, _P]
    emul.emul_ensemble_proba.argtypes = [_P, _P, _P, _I, _I, _I]

    class Stub(glue.HostStandIn):
        def koa_adam_step(self, table, n, hyper, stream):          # the kernel's own coefficient set-up and update
            return emul.emul_adam_step(table, n, hyper)

        def koa_augment_resample(self, src, dtype, out, params, batch, s, c, o, mean, std, ws, stream):
            return emul.emul_augment_resample(src, dtype, out, params, batch, s, c, o, mean, std)

        def koa_unit_range_affine(self, src, dtype, batch, n_per, mean, std, ws, scale, shift, minmax, stream):
            return emul.emul_unit_range_affine(src, dtype, batch, n_per, mean, std, scale, shift, minmax)

        def koa_predict(self, logits, proba, pred, b, c, stream):
            return emul.emul_predict(logits, proba, pred, b, c)

        def koa_ensemble_proba(self, proba, out, pred, f, b, c, stream):
            return emul.emul_ensemble_proba(proba, out, pred, f, b, c)

        def koa_resample_linear(self, src, dtype, out, batch, di, do, scale, shift, stream):
            return emul.emul_resample_linear(src, dtype, out, batch, di, do, scale, shift)

    stub = Stub()
    monkeypatch.setattr(_lib, "load", lambda: stub)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "on_device", lambda device: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "current_stream", lambda: None)
    cpu = torch.device("cpu")
    gpu_tests.test_augment_matches_reference_golden(cpu, gold)
    gpu_tests.test_augment_matches_the_oracle_on_random_states(cpu, (40, 36, 10), (32, 33, 7), torch.uint8, (0.5, 0.5, 0.5))
    gpu_tests.test_augment_matches_the_oracle_on_random_states(cpu, (50, 44), (41, 37), torch.uint16, (0.5, 0.5))
    gpu_tests.test_augment_recipe_sizes(cpu)
    gpu_tests.test_interpolate_matches_reference_golden(cpu, gold)
    gpu_tests.test_normalize_downscale_matches_reference_golden(cpu, gold)
    gpu_tests.test_interpolate_matches_the_oracle(cpu, (3, 1, 21, 18, 7), (0.5, 0.5, 0.5), torch.float32)
    gpu_tests.test_interpolate_matches_the_oracle(cpu, (2, 2, 30, 31), (0.75, 0.4), torch.float32)
    gpu_tests.test_interpolate_matches_the_oracle(cpu, (2, 1, 16, 12, 5), (0.5, 0.5, 1.0), torch.int16)
    gpu_tests.test_recipe_sizes_box_mean_identity_and_minmax(cpu)
    # Adam: adam_update / make_adam_coef of the kernel (host build) against torch.optim on the same device and against the
    # numpy oracle, with the tolerances the GPU tests use; softmax / argmax / fold ensemble likewise (predict_row, ensemble_row)
    from oaprogressionmmf_b200 import optim as koptim

    gpu_tests.test_adam_matches_torch_optim(cpu, koptim.Adam, torch.optim.Adam, 0.0)
    gpu_tests.test_adam_matches_torch_optim(cpu, koptim.Adam, torch.optim.Adam, 1e-4)
    gpu_tests.test_adam_matches_torch_optim(cpu, koptim.AdamW, torch.optim.AdamW, 1e-2)
    gpu_tests.test_adam_matches_the_oracle(cpu)
    gpu_tests.test_predict_and_ensemble(cpu, gold)
