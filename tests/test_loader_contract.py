"""The loader contract: items / batches with the reference's keys, the clinical vector of
``DatasetOAI3d.__getitem__`` (golden vectors made by executing the reference's own method source, ``oracle/
make_golden_step.py``), and the stored-volume variant whose transform chain runs on the device. CPU: the device call is
answered by the host build of the kernel arithmetic (``tests/host_emul``), the result checked against the numpy oracle."""
import contextlib
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oaprogressionmmf_b200 import _lib, synthetic as sy
from oracle import step_oracle as so
from tests.test_kernel_arith_host import emul  # noqa: F401  (module-scoped fixture: builds the harness once)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden_step", "step_rows.json")) as f:
        return json.load(f)


def test_clinical_vector_matches_the_reference_dataset(gold):
    assert len(gold["clin"]) == 3
    for case in gold["clin"]:
        r = case["row"]
        vec = sy.encode_clinical(r["AGE"], r["P02SEX"], r["P01BMI"], r["inj"], r["surg"], r["WOMTS"])
        assert vec.dtype == torch.float32 and vec.shape == (9,)
        np.testing.assert_allclose(vec.numpy(), np.asarray(case["vec"], np.float32), rtol=1e-6, atol=1e-7)


def test_item_and_batch_contract(gold):
    modals = ["xr_pa", "sag_3d_dess", "sag_t2_map", "clin"]
    sizes = {"xr_pa": (20, 20), "sag_3d_dess": (16, 16, 4), "sag_t2_map": (16, 16, 3)}
    ds = sy.SyntheticKneeDataset(modals, sizes, n=7)
    item = ds[3]
    ref_keys = set(gold["clin"][0]["keys"])            # keys of a reference item (clinical modality only)
    mine = {k if isinstance(k, str) else "|".join(k) for k in item}
    assert {"image__clin", "target", "-|side"} <= (ref_keys & mine)
    assert ("-", "exam_knee_id") in item
    assert item["image__clin"].shape == (1, 9) and item["target"].shape == (1,)
    assert item["image__xr_pa"].shape == (1, 20, 20) and item["image__sag_3d_dess"].shape == (1, 16, 16, 4)
    again = ds[3]
    assert torch.equal(item["image__sag_3d_dess"], again["image__sag_3d_dess"])       # deterministic per index
    # the reference's loaders: default collate, drop_last for train / val (datasets/_data_provider.py:478-498)
    dl = torch.utils.data.DataLoader(ds, batch_size=3, drop_last=True)
    batches = list(dl)
    assert len(batches) == 2
    b = batches[0]
    assert b["target"].shape == (3, 1) and b["target"].dtype == torch.int64
    assert b["image__clin"].shape == (3, 1, 9) and b["image__sag_t2_map"].shape == (3, 1, 16, 16, 3)
    assert len(b[("-", "exam_knee_id")]) == 3
    xs = tuple(b[f"image__{m}"] for m in modals)       # run/train_prog_fus.py:136-137
    assert [x.shape[0] for x in xs] == [3] * 4


def test_stored_batches_transformed_on_the_device(emul, monkeypatch):  # noqa: F811
    import tests.test_step_glue as glue

    _I, _P, _F = C.c_int, C.c_void_p, C.c_float
    emul.emul_augment_resample.argtypes = [_P, _I, _P, _P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _F, _F]

    class Stub(glue.HostStandIn):
        def koa_augment_resample(self, src, dtype, out, params, batch, s, c, o, mean, std, ws, stream):
            return emul.emul_augment_resample(src, dtype, out, params, batch, s, c, o, mean, std)

    stub = Stub()
    monkeypatch.setattr(_lib, "load", lambda: stub)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "on_device", lambda device: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "current_stream", lambda: None)

    import tests.test_zz_gpu_step_rows as gpu_tests

    gpu_tests.test_stored_batches_transformed_on_the_device(torch.device("cpu"))     # the -m gpu test, dry-run on the host
