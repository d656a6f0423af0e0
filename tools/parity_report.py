"""GPU parity report: CUDA path vs the oracle (fp32 torch restatement), with per-tensor relative errors.
Run under gpurun. Sections: fe (feature extractor), feat (transformer), models (golden cases).

    python tools/parity_report.py [fe feat models] [--arch resnet50]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oaprogressionmmf_b200 import _lib  # noqa: E402
from oaprogressionmmf_b200.koamodels import FeaT, SliceEncoder, dict_fes, dict_models  # noqa: E402
from oracle import koa_oracle as ko  # noqa: E402

RES_GAIN = 1.0
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


class AD(dict):
    __getattr__ = dict.__getitem__


def to_attr(d):
    if isinstance(d, dict):
        return AD({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return [to_attr(v) for v in d]
    return d


def rel(a, b):
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def section_fe(arch, n_b=2, slices=3, size=64, train=True, xr=False, emulate=False):
    print(f"== FE {arch} B={n_b} S={slices} {size}x{size} train={train} xr={xr} emulate_16bit={emulate}", flush=True)
    dev = "cuda"
    spec = ko.fe_param_spec(arch, "_fe")
    sd = ko.make_state_dict(spec, 11, device=dev, res_gain=RES_GAIN)
    enc = SliceEncoder(dict_fes[arch](pretrained=False), with_gap=True).to(dev)
    enc.load_state_dict({k[len("_fe."):]: v.clone() for k, v in sd.items()})
    enc.train(train)
    g = torch.Generator().manual_seed(5)
    if xr:
        vol = torch.randn(n_b, 1, size, size, generator=g).to(dev)
        imgs = vol.expand(-1, 3, -1, -1)
        tok = enc.encode_image(vol)
    else:
        vol = torch.randn(n_b, 1, size, size, slices, generator=g).to(dev)
        imgs = ko._slices_to_images(vol)
        tok = enc.encode_volume(vol)
    torch.cuda.synchronize()
    print("debug flag", hex(_lib.debug_flag()))
    # oracle
    params = {k: v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    for v in params.values():
        v.requires_grad_(True)
    taps = {}
    ref = ko.fe_forward(sd, "_fe", arch, imgs, train, True, taps, emulate_16bit=emulate).flatten(1)
    got = tok.reshape(-1, tok.shape[-1])
    print(f"features rel={rel(got, ref):.3e}  |ref|={ref.norm():.3f} finite={bool(torch.isfinite(got).all())}")
    # intermediates
    lib = _lib.load()
    n_img = imgs.shape[0]
    desc = _lib.FeDesc(arch=_lib.ARCH_IDS[arch], n_img=n_img, h=size, w=size, slices=0 if xr else slices, with_gap=1,
                       training=1 if train else 0, need_backward=1)
    ws = tok.grad_fn.ws if hasattr(tok.grad_fn, "ws") else None
    fn = tok.grad_fn
    while fn is not None and not hasattr(fn, "ws"):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    ws = fn.ws if fn is not None else None
    if ws is not None:
        off, nb = C.c_size_t(), C.c_size_t()
        _lib.check(lib.koa_fe_debug_offset(C.byref(desc), 4, 0, C.byref(off), C.byref(nb)), "dbg")
        a0 = ws[off.value:off.value + nb.value].view(torch.float16)
        r0 = taps["_fe.stem"].permute(0, 2, 3, 1).reshape(-1)
        print(f"  stem act rel={rel(a0.float(), r0):.3e}")
        for bi, b in enumerate(ko.fe_block_plan(arch)):
            _lib.check(lib.koa_fe_debug_offset(C.byref(desc), 1, bi, C.byref(off), C.byref(nb)), "dbg")
            out = ws[off.value:off.value + nb.value].view(torch.float16)
            r = taps[f"_fe.{b['layer']}.{b['index']}"].permute(0, 2, 3, 1).reshape(-1)
            print(f"  block {b['layer']}.{b['index']} out rel={rel(out.float(), r):.3e}")
    # backward
    gy = torch.randn(ref.shape, generator=g).to(dev)
    (got * gy).sum().backward()
    (ref * gy).sum().backward()
    torch.cuda.synchronize()
    print("debug flag", hex(_lib.debug_flag()))
    worst = []
    mine = dict(enc.named_parameters())
    for k, v in params.items():
        gm = mine[k[len("_fe."):]].grad
        e = rel(gm, v.grad)
        worst.append((e, k, float(v.grad.norm())))
    if os.environ.get("KOA_ALL_GRADS"):
        for e, k, nrm in reversed(worst):
            print(f"  grad {k}: rel={e:.3e} |ref|={nrm:.3e}")
    worst.sort(reverse=True)
    for e, k, nrm in worst[:12]:
        print(f"  grad {k}: rel={e:.3e} |ref|={nrm:.3e}")
    es = torch.tensor([w[0] for w in worst])
    print(f"  grads: max={es.max():.3e} median={es.median():.3e} n={len(worst)}")
    if train:
        msd = enc.state_dict()
        es = [rel(msd[k[len('_fe.'):]], v) for k, v in sd.items() if k.endswith(("running_mean", "running_var"))]
        print(f"  running stats: max rel={max(es):.3e}")


def _ws_view(lib, desc, ws, what, index, dtype):
    off, nb = C.c_size_t(), C.c_size_t()
    _lib.check(lib.koa_fe_debug_offset(C.byref(desc), what, index, C.byref(off), C.byref(nb)), "dbg")
    return ws[off.value:off.value + nb.value].view(dtype)


def _nhwc_bf16(t):
    return t.detach().permute(0, 2, 3, 1).contiguous().half().reshape(-1)


def section_fe_teacher(arch, n_b=2, slices=3, size=64, xr=False):
    """Backward validation with forced-identical forward state: the workspace activations saved by the CUDA
    forward are overwritten with the bf16-emulating oracle's, then koa_fe_backward runs on them. Removes the
    ReLU-mask / rounding-avalanche noise of the forward pass from the gradient comparison."""
    print(f"== FE teacher-forced backward {arch} B={n_b} S={slices} {size}x{size} xr={xr}", flush=True)
    dev = "cuda"
    lib = _lib.load()
    spec = ko.fe_param_spec(arch, "_fe")
    sd = ko.make_state_dict(spec, 11, device=dev, res_gain=RES_GAIN)
    enc = SliceEncoder(dict_fes[arch](pretrained=False), with_gap=True).to(dev)
    enc.load_state_dict({k[len("_fe."):]: v.clone() for k, v in sd.items()})
    enc.train(True)
    g = torch.Generator().manual_seed(5)
    if xr:
        vol = torch.randn(n_b, 1, size, size, generator=g).to(dev)
        imgs = vol.expand(-1, 3, -1, -1)
        tok = enc.encode_image(vol)
    else:
        vol = torch.randn(n_b, 1, size, size, slices, generator=g).to(dev)
        imgs = ko._slices_to_images(vol)
        tok = enc.encode_volume(vol)
    fn = tok.grad_fn
    while fn is not None and not hasattr(fn, "ws"):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    ws, desc = fn.ws, fn.desc
    params = {k: v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    for v in params.values():
        v.requires_grad_(True)
    taps = {}
    ref = ko.fe_forward(sd, "_fe", arch, imgs, True, True, taps, emulate_16bit=True).flatten(1)
    # unit order of the engine: stem, then per block conv1, conv2, [conv3], [downsample]
    ykeys = ["_fe.0.y"]
    for b in ko.fe_block_plan(arch):
        p = f"_fe.{b['layer']}.{b['index']}"
        ykeys += [f"{p}.conv1.y", f"{p}.conv2.y"]
        if b["kind"] == "bottleneck":
            ykeys.append(f"{p}.conv3.y")
        if b["downsample"]:
            ykeys.append(f"{p}.downsample.0.y")
    worst_fwd = 0.0
    for ui, key in enumerate(ykeys):
        y = taps[key]
        dst = _ws_view(lib, desc, ws, 0, ui, torch.float16)
        src = _nhwc_bf16(y)
        worst_fwd = max(worst_fwd, rel(dst.float(), src.float()))
        dst.copy_(src)
        coef = _ws_view(lib, desc, ws, 6, ui, torch.float32).view(7, -1)
        yf = y.detach()
        mean = yf.mean(dim=(0, 2, 3))
        var = yf.var(dim=(0, 2, 3), unbiased=False)
        coef[2].copy_(mean)
        coef[3].copy_(torch.rsqrt(var + 1e-5))
    print(f"  forward raw conv outputs before overwrite: worst rel={worst_fwd:.3e}")
    _ws_view(lib, desc, ws, 4, 0, torch.float16).copy_(_nhwc_bf16(taps["_fe.stem"]))
    for bi, b in enumerate(ko.fe_block_plan(arch)):
        p = f"_fe.{b['layer']}.{b['index']}"
        _ws_view(lib, desc, ws, 1, bi, torch.float16).copy_(_nhwc_bf16(taps[p]))
        _ws_view(lib, desc, ws, 2, bi, torch.float16).copy_(_nhwc_bf16(taps[f"{p}.a1"]))
        if b["kind"] == "bottleneck":
            _ws_view(lib, desc, ws, 3, bi, torch.float16).copy_(_nhwc_bf16(taps[f"{p}.a2"]))
    # pooled activation + argmax from the overwritten stem activation
    a0 = _ws_view(lib, desc, ws, 4, 0, torch.float16)
    p0 = _ws_view(lib, desc, ws, 5, 0, torch.float16)
    idx0 = _ws_view(lib, desc, ws, 7, 0, torch.uint8)
    hs = (size + 6 - 7) // 2 + 1
    _lib.check(lib.koa_maxpool_fwd(a0.data_ptr(), p0.data_ptr(), idx0.data_ptr(), imgs.shape[0], hs, hs, 64,
                                   _lib.current_stream()), "maxpool")
    print(f"  pooled rel vs oracle={rel(p0.float(), _nhwc_bf16(taps['_fe.pool']).float()):.3e}")
    gy = torch.randn(ref.shape, generator=g).to(dev)
    got = tok.reshape(-1, tok.shape[-1])
    (got * gy).sum().backward()
    (ref * gy).sum().backward()
    torch.cuda.synchronize()
    print("  debug flag", hex(_lib.debug_flag()))
    rows = []
    mine = dict(enc.named_parameters())
    for k, v in params.items():
        rows.append((rel(mine[k[len("_fe."):]].grad, v.grad), k, float(v.grad.norm())))
    if os.environ.get("KOA_ALL_GRADS"):
        for e, k, nrm in reversed(rows):
            print(f"  grad {k}: rel={e:.3e} |ref|={nrm:.3e}")
    rows.sort(reverse=True)
    for e, k, nrm in rows[:10]:
        print(f"  grad {k}: rel={e:.3e} |ref|={nrm:.3e}")
    es = torch.tensor([r[0] for r in rows])
    print(f"  grads: max={es.max():.3e} median={es.median():.3e} n={len(rows)}")


def section_floor(arch, n_b=2, slices=3, size=64, train=True):
    """Precision floor with no CUDA-path code involved: fp32 oracle vs bf16-emulating oracle."""
    print(f"== precision floor (oracle fp32 vs oracle bf16-emulated) {arch} B={n_b} S={slices} {size} train={train}", flush=True)
    dev = "cuda"
    spec = ko.fe_param_spec(arch, "_fe")
    g = torch.Generator().manual_seed(5)
    vol = torch.randn(n_b, 1, size, size, slices, generator=g).to(dev)
    imgs = ko._slices_to_images(vol)
    outs = []
    gy = None
    for emu in (False, True):
        sd = ko.make_state_dict(spec, 11, device=dev, res_gain=RES_GAIN)
        params = {k: v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
        for v in params.values():
            v.requires_grad_(True)
        taps = {}
        f = ko.fe_forward(sd, "_fe", arch, imgs, train, True, taps, emulate_16bit=emu).flatten(1)
        if gy is None:
            gy = torch.randn(f.shape, generator=g).to(dev)
        (f * gy).sum().backward()
        outs.append((f.detach(), {k: v.grad for k, v in params.items()}, taps))
    (f0, g0, t0), (f1, g1, t1) = outs
    print(f"  features rel={rel(f1, f0):.3e}")
    for b in ko.fe_block_plan(arch):
        p = f"_fe.{b['layer']}.{b['index']}"
        print(f"  block {b['layer']}.{b['index']} out rel={rel(t1[p], t0[p]):.3e}")
    es = torch.tensor([rel(g1[k], g0[k]) for k in g0])
    print(f"  grads: max={es.max():.3e} median={es.median():.3e}")


def section_feat(b=2, n_p=5, dim=2048, depth=2, heads=8, with_cls=True, head=True):
    print(f"== FeaT B={b} n_p={n_p} D={dim} depth={depth} cls={with_cls} head={head}", flush=True)
    dev = "cuda"
    spec = ko.feat_param_spec("_agg", n_p, dim, depth, dim, 2, with_cls)
    sd = ko.make_state_dict(spec, 21, pos_scale=0.5, device=dev)
    mod = FeaT(n_p, dim, dim, depth, heads, dim, 2, with_cls=with_cls).to(dev)
    mod.load_state_dict({k[len("_agg."):]: v.clone() for k, v in sd.items()})
    mod.train()
    g = torch.Generator().manual_seed(6)
    tok = (torch.randn(b, n_p, dim, generator=g) * 0.7).to(dev).requires_grad_(True)
    tok_ref = tok.detach().clone().requires_grad_(True)
    out, states, attns = mod.run(tok, compute_head=head)
    for v in sd.values():
        v.requires_grad_(True)
    out_ref, states_ref = ko.feat_forward(sd, "_agg", tok_ref, depth, heads, 0.0, 0.0, True)
    torch.cuda.synchronize()
    print("debug flag", hex(_lib.debug_flag()))
    print(f"states rel={rel(states, states_ref):.3e}")
    if head:
        print(f"logits rel={rel(out, out_ref):.3e}  got={out.flatten().tolist()} ref={out_ref.flatten().tolist()}")
    gs = torch.randn(states.shape, generator=g).to(dev)
    go = torch.randn(out.shape, generator=g).to(dev)
    loss = (states * gs).sum() + ((out * go).sum() if head else 0)
    loss_ref = (states_ref * gs).sum() + ((out_ref * go).sum() if head else 0)
    loss.backward()
    loss_ref.backward()
    torch.cuda.synchronize()
    print("debug flag", hex(_lib.debug_flag()))
    print(f"d_tokens rel={rel(tok.grad, tok_ref.grad):.3e}")
    mine = dict(mod.named_parameters())
    rows = []
    for k, v in sd.items():
        gm = mine[k[len("_agg."):]].grad
        if v.grad is None or gm is None:
            rows.append((float("nan"), k, "ref None" if v.grad is None else "mine None"))
            continue
        rows.append((rel(gm, v.grad), k, f"{float(v.grad.norm()):.3e}"))
    for e, k, n in sorted(rows, key=lambda r: -r[0] if r[0] == r[0] else 1):
        print(f"  grad {k}: rel={e:.3e} |ref|={n}")


def section_models(only=None):
    gold_dir = os.path.join(ROOT, "tests", "golden")
    for fn in sorted(os.listdir(gold_dir)):
        case = fn[:-5]
        if only and case not in only:
            continue
        with open(os.path.join(gold_dir, fn)) as f:
            gold = json.load(f)
        name = gold["model"]
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in gold["config_kwargs"].items()}
        cfg = ko.make_config(name, **kw)
        spec = ko.model_param_spec(name, cfg)
        inputs, target = ko.make_inputs(name, cfg, gold["batch"], gold["seed_inputs"], device="cuda")
        model = dict_models[name](to_attr(cfg), None).cuda()
        print(f"== model case {case}", flush=True)
        try:
            model.load_state_dict(ko.make_state_dict(spec, gold["seed_weights"], device="cuda"))
            model.eval()
            with torch.no_grad():
                lg = model(*inputs)["main"]
            ref = torch.tensor(gold["eval_logits"], device="cuda")
            print(f"  eval logits rel={rel(lg, ref):.3e} argmax_same={bool((lg.argmax(1) == ref.argmax(1)).all())}")
            model.load_state_dict(ko.make_state_dict(spec, gold["seed_weights"], pos_scale=0.02, device="cuda"))
            with torch.no_grad():
                lg = model(*inputs)["main"]
            ref = torch.tensor(gold["eval_logits_sensitised"], device="cuda")
            print(f"  eval(sens) logits rel={rel(lg, ref):.3e} argmax_same={bool((lg.argmax(1) == ref.argmax(1)).all())}")
            model.train()
            model.zero_grad()
            lg = model(*inputs)["main"]
            loss = ko.focal_loss(lg, target)
            loss.backward()
            torch.cuda.synchronize()
            ref = torch.tensor(gold["train_logits"], device="cuda")
            print(f"  train logits rel={rel(lg, ref):.3e}  loss={float(loss):.6f} ref={gold['train_loss']:.6f}")
            errs = []
            for k, p in model.named_parameters():
                gref = gold["grads"][k]
                if gref is None:
                    if p.grad is not None:
                        print(f"  !! {k}: reference grad None, ours not")
                    continue
                if p.grad is None:
                    print(f"  !! {k}: our grad None")
                    continue
                errs.append((abs(float(p.grad.norm()) - gref["norm"]) / (gref["norm"] + 1e-30), k, gref["norm"]))
            errs.sort(reverse=True)
            for e, k, n in errs[:6]:
                print(f"  gradnorm {k}: rel={e:.3e} |ref|={n:.3e}")
            es = torch.tensor([e[0] for e in errs])
            print(f"  gradnorms: max={es.max():.3e} median={es.median():.3e} n={len(errs)}")
            print("  debug flag", hex(_lib.debug_flag()))
        except Exception as ex:  # noqa: BLE001
            print(f"  FAILED: {type(ex).__name__}: {ex}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sections", nargs="*", default=["fe", "feat", "models"])
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--cases", nargs="*", default=None)
    ap.add_argument("--res-gain", type=float, default=1.0)
    args = ap.parse_args()
    global RES_GAIN
    RES_GAIN = args.res_gain
    if "fe" in args.sections:
        section_fe(args.arch, train=True)
        section_fe(args.arch, train=False)
    if "fe_emu" in args.sections:
        section_fe(args.arch, train=True, emulate=True)
        section_fe(args.arch, train=False, emulate=True)
        section_fe(args.arch, n_b=4, slices=16, size=128, train=True, emulate=True)
    if "fe_tf" in args.sections:
        section_fe_teacher(args.arch)
        section_fe_teacher(args.arch, n_b=2, slices=8, size=96)
    if "fe_tf_xr" in args.sections:
        section_fe_teacher("resnext50_32x4d", n_b=3, size=64, xr=True)
    if "fe_tf_r18" in args.sections:
        section_fe_teacher("resnet18")
    if "floor" in args.sections:
        section_floor(args.arch, train=True)
        section_floor(args.arch, train=False)
    if "floor2" in args.sections:
        section_floor(args.arch, n_b=4, slices=4, size=64, train=True)
        section_fe(args.arch, n_b=4, slices=4, size=64, train=True)
    if "fe_eval" in args.sections:
        section_fe(args.arch, train=False)
    if "fe_big" in args.sections:
        section_fe(args.arch, n_b=4, slices=16, size=128, train=True)
    if "fe_xr" in args.sections:
        section_fe("resnext50_32x4d", n_b=3, size=64, train=True, xr=True)
    if "fe_r18" in args.sections:
        section_fe("resnet18", train=True)
    if "feat" in args.sections:
        section_feat()
        section_feat(b=3, n_p=7, depth=1, with_cls=False, head=False)
    if "models" in args.sections:
        section_models(args.cases)


if __name__ == "__main__":
    main()
