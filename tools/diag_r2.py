"""Round-2 diagnostics: (1) is koa_fe_backward repeatable on one workspace, whole vs staged; (2) which tensors carry the large
gradient-probe errors of the full-size fixtures."""
import ctypes as C, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oaprogressionmmf_b200 import _lib
from oracle import koa_oracle as ko
from tests.util import rel, to_attr
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_parity as T

lib = _lib.load()
cuda = "cuda"
if "stages" in sys.argv:
    for arch, s in (("resnet18", 2), ("resnet50", 3)):
        sd, enc = T._fe_pair(arch, cuda)
        enc.train()
        x = T._randn(2, 1, 64, 64, s, seed=5)
        tok = enc.encode_volume(x)
        gy = T._randn(*tok.shape, seed=7)
        fn = tok.grad_fn
        while fn is not None and not hasattr(fn, "ws"):
            fn = fn.next_functions[0][0] if fn.next_functions else None
        ws, desc, table = fn.ws, fn.desc, fn.table
        params = enc._trainable()
        nb = lib.koa_fe_num_blocks(C.byref(desc))
        def run(ranges):
            grads, flat = _lib.zeros_like_flat(params)
            gt = _lib.ptr_table(grads)
            d = gy.reshape(-1, gy.shape[-1]).contiguous()
            for b, e, st in ranges:
                _lib.check(lib.koa_fe_backward_range(C.byref(desc), table, gt, ws.data_ptr(), d.data_ptr(), b, e, st, _lib.current_stream()), "r")
            torch.cuda.synchronize()
            return flat.clone(), [g.clone() for g in grads]
        w = [run([(0, -1, 1)]) for _ in range(3)]
        print(arch, "whole vs whole:", rel(w[1][0], w[0][0]), rel(w[2][0], w[0][0]))
        st = [(b, e, s_) for b, e, s_, *_ in enc._stage_bounds(params, *_lib.zeros_like_flat(params))]
        g = run(st)
        print(arch, "staged vs whole:", rel(g[0], w[0][0]), st)
        names = [f"u{i//3}.{'wgb'[i%3]}" for i in range(len(params))]
        worst = sorted(((rel(a, b), n) for a, b, n in zip(g[1], w[0][1], names)), reverse=True)[:6]
        print("  worst tensors staged:", worst)
        worst = sorted(((rel(a, b), n) for a, b, n in zip(w[1][1], w[0][1], names)), reverse=True)[:6]
        print("  worst tensors whole2:", worst)
if "probe" in sys.argv:
    from oaprogressionmmf_b200.losses import FocalLoss
    case = "MR1CnnTrf_full"
    gold_dir = os.path.join(os.path.dirname(T.__file__), "golden_full")
    gold, cfg, model, inputs, target = T._golden_case(case, gold_dir, cuda, pos_scale=0.02)
    model.train()
    loss = FocalLoss(gamma=2)(model(*inputs)["main"], target)
    loss.backward()
    spec = ko.model_param_spec(gold["model"], cfg)
    sd = ko.make_state_dict(spec, gold["seed_weights"], pos_scale=0.02, device=cuda)
    _, _, g_emu = ko.train_step(gold["model"], cfg, sd, inputs, target, emulate_16bit=True)
    sd2 = ko.make_state_dict(spec, gold["seed_weights"], pos_scale=0.02, device=cuda)
    _, _, g_f32 = ko.train_step(gold["model"], cfg, sd2, inputs, target)
    rows = []
    for k, p in model.named_parameters():
        gr = gold["grads"][k]
        if gr is None or gr["norm"] == 0: continue
        pr = T._probe(p.grad.numel(), cuda).double()
        mine = abs(float((p.grad.flatten().double() * pr).sum()) - gr["probe"]) / gr["norm"]
        emu = abs(float((g_emu[k].flatten().double() * pr).sum()) - gr["probe"]) / gr["norm"]
        f32 = abs(float((g_f32[k].flatten().double() * pr).sum()) - gr["probe"]) / gr["norm"]
        rows.append((k, mine, emu, f32, rel(p.grad, g_f32[k]), rel(g_emu[k], g_f32[k]), gr["norm"]))
    print("key | probe err mine | emu | oracle-f32-on-gpu | relL2 mine vs f32 oracle | relL2 emu vs f32 | ref norm")
    for r in rows[:12] + rows[-12:]:
        print(f"{r[0][:46]:46s} {r[1]:.3e} {r[2]:.3e} {r[3]:.3e} {r[4]:.3e} {r[5]:.3e} {r[6]:.3e}")
    import statistics
    print("median relL2 mine", statistics.median(r[4] for r in rows), "emu", statistics.median(r[5] for r in rows))
    bad = sorted(rows, key=lambda r: -(r[1] - r[2]))[:10]
    for r in bad:
        print("BAD", r)
