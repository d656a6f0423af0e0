"""GPU report: CUDA path vs the full-size fixtures of the unmodified reference (tests/golden_full)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oaprogressionmmf_b200.koamodels import dict_models
from oaprogressionmmf_b200.losses import FocalLoss
from oracle import koa_oracle as ko
from tests.util import rel, to_attr

d = os.path.join(ROOT, "tests", "golden_full")
for fn in sorted(os.listdir(d)):
    gold = json.load(open(os.path.join(d, fn)))
    name = gold["model"]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in gold["config_kwargs"].items()}
    cfg = ko.make_config(name, **kw)
    spec = ko.model_param_spec(name, cfg)
    inputs, target = ko.make_inputs(name, cfg, gold["batch"], gold["seed_inputs"], device="cuda")
    model = dict_models[name](to_attr(cfg), None).cuda()
    for tag, ps in (("eval_logits", 1.0), ("eval_logits_sensitised", 0.02)):
        model.load_state_dict(ko.make_state_dict(spec, gold["seed_weights"], pos_scale=ps, device="cuda"))
        model.eval()
        with torch.no_grad():
            lg = model(*inputs)["main"]
        ref = torch.tensor(gold[tag], device="cuda")
        print(f"{fn} {tag}: rel={rel(lg, ref):.3e} argmax_same={bool((lg.argmax(1) == ref.argmax(1)).all())} ref={ref.flatten().tolist()[:4]}", flush=True)
    model.train()
    model.zero_grad(set_to_none=True)
    lg = model(*inputs)["main"]
    loss = FocalLoss(gamma=2)(lg, target)
    loss.backward()
    ref = torch.tensor(gold["train_logits"], device="cuda")
    print(f"{fn} train logits rel={rel(lg, ref):.3e} loss={float(loss):.6f} ref={gold['train_loss']:.6f}")
    rs = []
    for k, p in model.named_parameters():
        g = gold["grads"][k]
        if g is None:
            assert p.grad is None, k
            continue
        rs.append(float(p.grad.norm()) / (g["norm"] + 1e-30))
    r = torch.tensor(rs)
    print(f"{fn} grad-norm ratio: min={float(r.min()):.3f} max={float(r.max()):.3f} median={float(r.median()):.3f} n={len(rs)}", flush=True)
    del model
    torch.cuda.empty_cache()
