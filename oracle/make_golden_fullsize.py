"""Full-size fixtures from the UNMODIFIED reference (build container only; minutes of CPU time per case).

    python oracle/make_golden_fullsize.py [case ...]      # writes tests/golden_full/<case>.json

Same protocol and file format as oracle/make_golden.py, but at the sizes BASELINE.json quotes its metric on
(XR 350x350, DESS 160x160x64, TSE 160x160x32, T2 map 160x160x25, D = 2048, depth 4, 8 heads): this is where the
1e-2 logit tolerance of the north star is defined. Batch 2 keeps the reference's CPU run to a few minutes.
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

CASES = {
    "XR1MR2C1CnnTrf_full": dict(model="XR1MR2C1CnnTrf", kw=dict(), batch=2),   # the reference's full model
    "XR1MR3C1CnnTrf_full": dict(model="XR1MR3C1CnnTrf", kw=dict(), batch=2),   # 3-MRI extension (bench workload)
    "MR1CnnTrf_full": dict(model="MR1CnnTrf", kw=dict(), batch=2),             # BASELINE.json config 2
    "XR1Cnn_full": dict(model="XR1Cnn", kw=dict(), batch=8),                   # BASELINE.json config 1
    "MR2CnnTrf_full": dict(model="MR2CnnTrf", kw=dict(), batch=2),             # BASELINE.json config 3 (strict): DESS + TSE, flat
    "XR1MR1CnnTrf_full": dict(model="XR1MR1CnnTrf", kw=dict(), batch=2),       # XR + DESS, flat
    "XR1MR2CnnTrf_full": dict(model="XR1MR2CnnTrf", kw=dict(slices=(64, 25, 25)), batch=2),  # XR + DESS + T2, no clinical token
    "MR3CnnTrf_full": dict(model="MR3CnnTrf", kw=dict(), batch=2),             # 3-MRI extension without XR / clinical
}


def main():
    out_dir = os.path.join(mg.ROOT, "tests", "golden_full")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for case_name, case in CASES.items():
        if only and case_name not in only:
            continue
        t0 = time.time()
        res = mg.run_case(case_name, case)
        path = os.path.join(out_dir, f"{case_name}.json")
        with open(path, "w") as f:
            json.dump(res, f, indent=1)
        print(f"{case_name}: {time.time() - t0:.1f}s -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


if __name__ == "__main__":
    main()
