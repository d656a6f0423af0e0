"""Generates tests/golden/*.json from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

For every reference model class this script
  1. builds ``koafusion.models.dict_models[name]`` from /root/reference with ``pretrained=False``,
  2. checks that ``oracle.koa_oracle.model_param_spec`` lists exactly the reference's state_dict keys
     and shapes, in order,
  3. loads seeded weights (``make_state_dict``) and seeded inputs (``make_inputs``),
  4. records eval-mode logits (after BN running stats were perturbed by the seeded init), one
     train-mode step with FocalLoss (logits, loss, per-parameter gradient L2 norm and three sampled
     gradient values, updated BN running statistics of the first/last BN layer), and a
     "sensitised" eval run (pos_embedding / cls_token scaled by 0.02).
The fixtures are small (no weights): both sides regenerate weights/inputs from (spec, seed).
The 3-MRI extension models have no reference class; their fixtures are produced from a composition of
the reference's own FeaT + dict_fes blocks, assembled the way _xrNmrMcP.py:40-179,209-255 does.
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import koa_oracle as ko  # noqa: E402


from oracle.ref_loader import build_model, focal_loss, to_attr  # noqa: E402


def tensor_summary(t):
    """What a fixture keeps of an intermediate tensor: shape, L2 norm, mean and five values at fixed flat positions."""
    flat = t.flatten().double()
    n = flat.numel()
    idx = [0, n // 4, n // 2, (3 * n) // 4, n - 1]
    return dict(shape=list(t.shape), norm=float(flat.norm()), mean=float(flat.mean()), samples=[float(flat[i]) for i in idx])


def load_ref_focal_loss():
    return focal_loss(gamma=2)


# Small but structurally complete cases (CPU seconds each). Sizes must be keys of the reference's
# spatial lookup table (_xrNmrMcP.py:104-105): 32/64/... for images.
CASES = {
    "XR1Cnn": dict(kw=dict(xr_size=64), batch=3),
    "XR1Cnn_r18": dict(model="XR1Cnn", kw=dict(xr_size=64, xr_arch="resnet18"), batch=2),
    "MR1CnnTrf": dict(kw=dict(mr_size=64, slices=(4,), depth=2), batch=2),
    "MR2CnnTrf": dict(kw=dict(mr_size=32, slices=(3, 2), depth=1), batch=2),
    "XR1MR1CnnTrf": dict(kw=dict(xr_size=64, mr_size=32, slices=(3,), depth=1), batch=2),
    "XR1MR2CnnTrf": dict(kw=dict(xr_size=64, mr_size=32, slices=(3, 2), depth=1), batch=2),
    "XR1MR2C1CnnTrf": dict(kw=dict(xr_size=64, mr_size=32, slices=(3, 2), depth=1), batch=2),
    "MR3CnnTrf": dict(kw=dict(mr_size=32, slices=(3, 2, 2), depth=1), batch=2),
    "XR1MR3C1CnnTrf": dict(kw=dict(xr_size=64, mr_size=32, slices=(3, 2, 2), depth=1), batch=2),
    # the other slicing axes of MR1CnnTrf (_mrN_cnn_trf.py:114-117): the volume is cubic so that every view gives 32 x 32 images
    "MR1CnnTrf_cs": dict(model="MR1CnnTrf", kw=dict(mr_size=32, slices=(32,), depth=1, dims_view="cs"), batch=1),
    "MR1CnnTrf_rs": dict(model="MR1CnnTrf", kw=dict(mr_size=32, slices=(32,), depth=1, dims_view="rs"), batch=1),
    # extractors without the global average pool: one token per spatial position (64 -> 2 x 2, 32 -> 1 x 1). MR1CnnTrf looks
    # up all three dimensions in its size table (_mrN_cnn_trf.py:55-56), so the slice count must be one of its keys
    "MR1CnnTrf_nogap": dict(model="MR1CnnTrf", kw=dict(mr_size=32, slices=(32,), depth=1, with_gap=False), batch=1),
    "XR1MR2C1CnnTrf_nogap": dict(model="XR1MR2C1CnnTrf",
                                 kw=dict(xr_size=64, mr_size=64, slices=(3, 2), depth=1, with_gap=False), batch=2),
}


def run_case(case_name, case):
    name = case.get("model", case_name)
    cfg = ko.make_config(name, **case["kw"])
    torch.manual_seed(778)
    model = build_model(name, cfg)

    spec = ko.model_param_spec(name, cfg)
    ref_sd = model.state_dict()
    assert [k for k, _ in spec] == list(ref_sd.keys()), f"{name}: key order mismatch"
    for k, shp in spec:
        assert tuple(ref_sd[k].shape) == tuple(shp), f"{name}: shape mismatch at {k}"

    seed_w, seed_x = 1000 + len(case_name), 2000 + len(case_name)
    out = dict(case=case_name, model=name, config_kwargs=case["kw"], batch=case["batch"], seed_weights=seed_w,
               seed_inputs=seed_x, num_keys=len(spec))
    inputs, target = ko.make_inputs(name, cfg, case["batch"], seed_x)

    def logits_of(m, ins):
        r = m(*ins)
        return r["main"] if isinstance(r, dict) else r

    # eval mode, with the intermediates the logits hide (at random initialisation the logits barely depend on
    # the input): every extractor output and every transformer's token states, as norm / mean / five sampled values
    model.load_state_dict(ko.make_state_dict(spec, seed_w), strict=True)
    model.eval()
    taps = {}

    def tap(key):
        def hook(_m, _i, o):
            t = o[1] if isinstance(o, tuple) else o          # FeaT returns (outputs, states, attentions)
            taps[key] = tensor_summary(t.detach())
        return hook

    handles = [m.register_forward_hook(tap(k)) for k, m in model.named_children()
               if (k.startswith("_fe") or k.startswith("_agg")) and not k.endswith("_drop") and k != "_agg" or
               (k == "_agg" and not isinstance(m, torch.nn.Sequential))]
    with torch.no_grad():
        out["eval_logits"] = logits_of(model, inputs).tolist()
    for h in handles:
        h.remove()
    out["taps"] = taps
    # sensitised eval
    model.load_state_dict(ko.make_state_dict(spec, seed_w, pos_scale=0.02), strict=True)
    with torch.no_grad():
        out["eval_logits_sensitised"] = logits_of(model, inputs).tolist()
    # one training step (dropout is 0 in these configs)
    model.load_state_dict(ko.make_state_dict(spec, seed_w, pos_scale=0.02), strict=True)
    model.train()
    model.zero_grad()
    logits = logits_of(model, inputs)
    loss = load_ref_focal_loss()(logits, target)
    loss.backward()
    out["train_logits"] = logits.detach().tolist()
    out["train_loss"] = float(loss)
    grads = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            grads[k] = None
        else:
            flat = p.grad.flatten()
            idx = [0, flat.numel() // 2, flat.numel() - 1]
            grads[k] = dict(norm=float(flat.norm()), samples=[float(flat[i]) for i in idx])
            # direction, not only magnitude: the projection of the gradient on a fixed seeded probe of the same shape
            # (a sign or permutation error inside a tensor leaves the norm alone and destroys this number)
            probe = torch.randn(flat.numel(), generator=torch.Generator().manual_seed(flat.numel() % 9973 + 17))
            grads[k]["probe"] = float((flat.double() * probe.double()).sum())
    out["grads"] = grads
    sd_after = model.state_dict()
    bn_keys = [k for k in sd_after if k.endswith("running_mean") or k.endswith("running_var")]
    picks = bn_keys[:2] + bn_keys[-2:]
    out["bn_after"] = {k: dict(sum=float(sd_after[k].sum()), first=float(sd_after[k].flatten()[0])) for k in picks}
    out["num_batches_tracked_after"] = int(sd_after[[k for k in sd_after if k.endswith("num_batches_tracked")][0]])
    return out


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for case_name, case in CASES.items():
        if only and case_name not in only:
            continue
        t0 = time.time()
        res = run_case(case_name, case)
        path = os.path.join(ROOT, "tests", "golden", f"{case_name}.json")
        with open(path, "w") as f:
            json.dump(res, f, indent=1)
        print(f"{case_name}: {time.time() - t0:.1f}s -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
