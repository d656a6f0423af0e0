#!/usr/bin/env python
"""HBM roofline of the kernels next to the hot path (step_ops.cu) on one B200: algorithmic GB/s against the measured
HBM peak of MEASURED_PEAKS.json. CUDA events on the launching stream, warm-up, working sets larger than the 126 MB L2.

    python tools/step_ops_bench.py [--iters 20]

Prints one JSON line per kernel."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oaprogressionmmf_b200 import optim as koptim, preproc  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6536.4
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = json.load(open(path)).get("hbm_gbs", peak)

    def report(name, ms, nbytes, **kw):
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps(dict(kernel=name, ms=ms, algorithmic_bytes=nbytes, achieved_gbs=gbs, peak_gbs=peak,
                              frac=gbs / peak, **kw)), flush=True)

    # Adam on the parameter inventory of the full model: ~1500 tensors, 530 M elements (ResNet-50-like size mix)
    sizes = []
    for _ in range(4):  # four extractors: conv weights + 3 small BatchNorm tensors each
        for cout, cin, k in [(64, 64, 1), (64, 64, 3), (256, 64, 1), (128, 256, 1), (128, 128, 3), (512, 128, 1),
                             (256, 512, 1), (256, 256, 3), (1024, 256, 1), (512, 1024, 1), (512, 512, 3), (2048, 512, 1)] * 4:
            sizes += [cout * cin * k * k, cout, cout]
    for _ in range(4):  # four transformers, depth 4
        sizes += [2048 * 2048, 2048] + [6144 * 2048, 2048 * 2048, 2048, 2048 * 2048, 2048, 2048 * 2048, 2048, 2048, 2048,
                                        2048, 2048] * 4
    params = [torch.nn.Parameter(torch.randn(n, device=dev) * 0.02) for n in sizes]
    for p in params:
        p.grad = torch.randn_like(p) * 1e-3
    opt = koptim.Adam(params, lr=1e-4, weight_decay=1e-4)
    n = sum(sizes)
    ms = timed(opt.step, args.iters)
    report("adam_kernel", ms, 28.0 * n, tensors=len(sizes), elements=n)
    ref = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-4, foreach=True)
    ms_t = timed(ref.step, max(3, args.iters // 4))
    print(json.dumps(dict(kernel="torch.optim.Adam(foreach=True), same tensors", ms=ms_t)), flush=True)
    del params, opt, ref
    torch.cuda.empty_cache()

    # resampling at the sizes of the reference recipes, 16 knees
    for name, shape, factor, dtype in [("dess_u8", (16, 1, 320, 320, 128), (0.5, 0.5, 0.5), torch.uint8),
                                       ("dess_f32", (16, 1, 320, 320, 128), (0.5, 0.5, 0.5), torch.float32),
                                       ("t2_u8", (16, 1, 320, 320, 25), (0.5, 0.5, 1.0), torch.uint8),
                                       ("xr_u16", (16, 1, 700, 700), (0.5, 0.5), torch.int16)]:
        if dtype == torch.float32:
            x = torch.randn(shape, device=dev)
        else:
            x = torch.randint(0, 200, shape, device=dev, dtype=dtype)
        out_elems = x.numel()
        for f in factor:
            out_elems *= f
        nbytes = x.numel() * x.element_size() + 4.0 * out_elems
        ms = timed(lambda: preproc.downscale_x(x, factor), args.iters)
        report(f"resample_linear_kernel[{name}]", ms, nbytes, shape=list(shape), factor=list(factor))
        ms = timed(lambda: preproc.unit_range_affine(x, 0.257, 0.235), args.iters)
        report(f"minmax_kernel[{name}]", ms, float(x.numel() * x.element_size()), shape=list(shape))
        del x

    # the training loader's chain fused with the downscale (koa_augment_resample), 16 knees: plain chain, rotation +
    # gamma on every knee (the worst case: four taps and one powf per stored voxel)
    import random

    for name, stored, crop, factor, dtype in [("dess_u8", (352, 340, 136), (320, 320, 128), (0.5, 0.5, 0.5), torch.uint8),
                                              ("xr_u16", (720, 712), (700, 700), (0.5, 0.5), torch.int16)]:
        x = torch.randint(0, 200, (16, 1) + stored, device=dev, dtype=dtype)
        rng = random.Random(1)
        crop_elems = 16
        for c in crop:
            crop_elems *= c
        out_elems = crop_elems
        for f in factor:
            out_elems *= f
        nbytes = 2.0 * crop_elems * x.element_size() + 4.0 * out_elems   # min / max pass + resampling pass + output
        for label, rp, gp in [("plain", 0.0, 0.0), ("train p=0.5", 0.5, 0.5), ("rotate+gamma", 1.0, 1.0)]:
            states = [preproc.draw_train_state(rng, stored, crop, rotate_prob=rp, gamma_prob=gp) for _ in range(16)]
            ms = timed(lambda: preproc.augment_normalize_downscale(x, crop, states, 0.257, 0.235, factor), args.iters)
            report(f"augment_resample (3 launches)[{name}, {label}]", ms, nbytes, stored=list(stored), crop=list(crop))
        del x


if __name__ == "__main__":
    main()
