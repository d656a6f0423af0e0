"""The oracle (oracle/koa_oracle.py) replayed against fixtures produced by the unmodified reference
(oracle/make_golden.py). CPU only. Tolerances: both sides are fp32 on the same CPU, so 1e-4 relative."""
import glob
import json
import os

import pytest
import torch

from oracle import koa_oracle as ko

CASES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))


def _close(a, b, rel=1e-4, abs_=1e-6):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return bool(((a - b).norm() <= rel * b.norm() + abs_))


def test_fixture_inventory():
    # every reference class (koafusion/models/__init__.py:8-15) + the two 3-MRI extensions
    for name in ko.MODEL_NAMES:
        assert name in CASES


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference(case, golden_dir):
    with open(os.path.join(golden_dir, f"{case}.json")) as f:
        gold = json.load(f)
    name = gold["model"]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in gold["config_kwargs"].items()}
    cfg = ko.make_config(name, **kw)
    spec = ko.model_param_spec(name, cfg)
    assert len(spec) == gold["num_keys"]
    inputs, target = ko.make_inputs(name, cfg, gold["batch"], gold["seed_inputs"])

    sd = ko.make_state_dict(spec, gold["seed_weights"])
    with torch.no_grad():
        logits = ko.model_forward(name, cfg, sd, inputs, training=False)
    assert _close(logits, gold["eval_logits"]), (logits, gold["eval_logits"])

    sd = ko.make_state_dict(spec, gold["seed_weights"], pos_scale=0.02)
    with torch.no_grad():
        logits = ko.model_forward(name, cfg, sd, inputs, training=False)
    assert _close(logits, gold["eval_logits_sensitised"])

    sd = ko.make_state_dict(spec, gold["seed_weights"], pos_scale=0.02)
    logits, loss, grads = ko.train_step(name, cfg, sd, inputs, target)
    assert _close(logits, gold["train_logits"])
    assert _close(loss, gold["train_loss"])
    assert set(grads) == set(gold["grads"])
    for k, g in gold["grads"].items():
        if g is None:
            # dead-head parameters of the per-sequence transformers never receive a gradient
            assert grads[k] is None, k
            continue
        flat = grads[k].flatten()
        assert _close(flat.norm(), g["norm"], rel=2e-3, abs_=1e-9), (k, float(flat.norm()), g["norm"])
        idx = [0, flat.numel() // 2, flat.numel() - 1]
        got = torch.stack([flat[i] for i in idx])
        assert _close(got, g["samples"], rel=5e-3, abs_=1e-6 + 1e-3 * g["norm"] / max(1.0, flat.numel() ** 0.5)), k
    for k, v in gold["bn_after"].items():
        assert _close(sd[k].sum(), v["sum"], rel=1e-4, abs_=1e-4), k
        assert _close(sd[k].flatten()[0], v["first"], rel=1e-4, abs_=1e-5), k
    nbt = [k for k in sd if k.endswith("num_batches_tracked")][0]
    assert int(sd[nbt]) == gold["num_batches_tracked_after"]
