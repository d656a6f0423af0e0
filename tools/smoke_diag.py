import sys, os
sys.path.insert(0, "/root/repo")
import torch
from oaprogressionmmf_b200.koamodels import dict_models
from oaprogressionmmf_b200.synthetic import to_attr
from oracle import koa_oracle as ko
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
name = "MR1CnnTrf"
for seed in (1234, 1, 2, 3):
    cfg = ko.make_config(name, mr_size=64, slices=(8,), depth=2)
    spec = ko.model_param_spec(name, cfg)
    sd = ko.make_state_dict(spec, seed, device=dev)
    inputs, target = ko.make_inputs(name, cfg, 2, 4321, device=dev)
    model = dict_models[name](to_attr(cfg), None).to(dev)
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        got = model(*inputs)["main"]
        ref = ko.model_forward(name, cfg, sd, inputs, training=False)
    rel = float((got - ref).norm() / ref.norm())
    print(seed, "rel", rel, "got", got.flatten().tolist(), "ref", ref.flatten().tolist())
