#!/usr/bin/env python
"""Benchmark of the koafusion hot path on B200: knees/sec for forward + FocalLoss + backward of the fusion
model (metric of BASELINE.json), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's algorithm on the host cores (oracle port)

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for the definitions of every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "knees/sec fwd+bwd (XR+3MRI+clin fusion) at 1/2/4/8 B200; % tensor-core roofline"

# Algorithmic GFLOP per knee (2*MAC of convolutions, linears and attention matmuls; backward = 2x forward
# minus the stem data gradient), BASELINE.md section 2 / SURVEY.md section 8(d).
FLOPS_FWD_BWD = {"XR1Cnn": 62.06e9, "MR1CnnTrf": 834.5e9, "MR2CnnTrf": 1251.7e9, "XR1MR2CnnTrf": 1279.6e9,
                 "XR1MR2C1CnnTrf": 1280.2e9, "MR3CnnTrf": 1654.5e9, "XR1MR3C1CnnTrf": 1717.8e9}
WORKLOAD_DESC = {
    "XR1MR3C1CnnTrf": "XR 350x350 (ResNeXt-50) + DESS 160x160x64 + TSE 160x160x32 + T2map 160x160x25 (ResNet-50 each) + 9 "
                      "clinical vars, per-sequence + fusion transformers (D 2048, depth 4); 3-MRI extension of the "
                      "reference's XR1MR2C1CnnTrf (BASELINE.json config 4)",
    "XR1MR2C1CnnTrf": "reference full model: XR 350x350 + DESS 160x160x64 + T2map 160x160x25 + clinical (runner.sh:341-363)",
    "MR1CnnTrf": "DESS 160x160x64 only: per-slice ResNet-50 + slice-aggregation transformer (BASELINE.json config 2)",
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p.get("bf16_tflops_sustained", 1386.1), burst=p.get("bf16_tflops", 1663.1),
                    hbm=p.get("hbm_gbs", 6536.4), source="MEASURED_PEAKS.json (sustained bf16 GEMM)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback of B200_PROFILING.md")


def split_by_bound(path, peaks, ms_total):
    """The GEMM family launch by launch (per-shape records of koa_profile_dump): a launch is tensor-bound when its
    algorithmic intensity 2MNK / bytes exceeds the measured ridge (peak TFLOP/s / peak GB/s), HBM-bound otherwise.
    Algorithmic bytes of a fprop / dgrad launch: operands once (a 3x3 im2col operand counts its input once, not nine
    times), output once, 2 B per element of every fused epilogue operand; of a weight-gradient launch: dY and X once,
    dW as fp32."""
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm"] * 1e9)
    acc = {"tensor": [0.0, 0.0, 0.0, 0], "hbm": [0.0, 0.0, 0.0, 0]}  # ms, flops, bytes, launches
    try:
        rows = [l.split() for l in open(path) if not l.startswith("#")]
    except OSError:
        return None
    for r in rows:
        cls, tag, m, n, k, cnt = (int(v) for v in r[:6])
        ms = float(r[6])
        tag &= 255
        if cls == 0:
            taps = 9 if (tag & 1) and k % 9 == 0 else 1
            by = 2.0 * (m * k / taps + n * k + m * n)
            for bit in (4, 8, 128):
                if tag & bit:
                    by += 2.0 * m * n
            if tag & 16:
                by += 4.0 * m * n
            if tag & 32:
                by += 2.0 * m * n
        else:
            taps = 9 if (tag & 1) and n % 9 == 0 else 1
            by = 2.0 * k * (m + n / taps) + 4.0 * m * n
        fl = 2.0 * m * n * k
        a = acc["tensor" if fl / by > ridge else "hbm"]
        a[0] += ms; a[1] += fl * cnt; a[2] += by * cnt; a[3] += cnt
    out = {"ridge_flop_per_byte": ridge}
    t, h = acc["tensor"], acc["hbm"]
    if t[0] > 0:
        out["tensor"] = dict(launches=t[3], share_of_step=t[0] / ms_total, achieved=t[1] / (t[0] * 1e-3) / 1e12,
                             peak=peaks["tflops"], unit="TFLOP/s", frac=t[1] / (t[0] * 1e-3) / 1e12 / peaks["tflops"])
    if h[0] > 0:
        out["hbm"] = dict(launches=h[3], share_of_step=h[0] / ms_total, achieved=h[2] / (h[0] * 1e-3) / 1e9,
                          peak=peaks["hbm"], unit="GB/s (algorithmic bytes)", frac=h[2] / (h[0] * 1e-3) / 1e9 / peaks["hbm"])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def _phase(msg):
    """Progress marker on stderr (the JSON line is the only thing on stdout): tells where a run that never finished stopped."""
    mem = ""
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        mem = f" [reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB, allocated {torch.cuda.memory_allocated() / 2**30:.1f} GiB]"
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}{mem}", file=sys.stderr, flush=True)


class Watchdog:
    """The JSON line must come out even if one of the optional legs after the main measurement (end-to-end pass, CPU
    baseline, full step, inference) never returns: once `arm` has been called with the line measured so far, a daemon thread
    prints it and ends the process when the deadline passes. `disarm` before the normal print."""

    def __init__(self):
        self.line = None
        self.deadline = None
        self.stall_seconds = None  # hunt mode: fire when `beat` has not been called for this long
        self.last_beat = time.time()
        self.lock = threading.Lock()
        self.done = False
        threading.Thread(target=self._run, daemon=True).start()

    def arm(self, line, seconds):
        with self.lock:
            self.line = dict(line)
            self.deadline = time.time() + seconds

    def update(self, **kv):
        with self.lock:
            if self.line is not None:
                self.line.update(kv)

    def beat(self):
        self.last_beat = time.time()

    def disarm(self):
        with self.lock:
            self.done = True

    def _run(self):
        while True:
            time.sleep(1.0)
            with self.lock:
                if self.done:
                    return
                stalled = self.stall_seconds is not None and time.time() - self.last_beat > self.stall_seconds
                if self.deadline is not None and (time.time() > self.deadline or stalled):
                    self.line["watchdog"] = "an optional leg exceeded its time budget: fields measured after `value` may be missing"
                    print(json.dumps(self.line), flush=True)
                    self._post_mortem()
                    os._exit(0)

    @staticmethod
    def _post_mortem():
        """Where every Python thread stands and what the GPU is doing, on stderr (the JSON line is already out)."""
        try:
            import faulthandler

            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            from oaprogressionmmf_b200 import _lib

            latest, first = _lib.debug_flag_peek()
            print(f"[bench watchdog] barrier time-out codes: latest {latest:#x}, first {first:#x}", file=sys.stderr, flush=True)
            buf = C.create_string_buffer(1 << 16)
            n = _lib.load().koa_profile_pending(buf, len(buf))
            print(f"[bench watchdog] {n} profiled launch(es) started but not finished (cls tag m n k):\n"
                  f"{buf.value.decode()}", file=sys.stderr, flush=True)
            q = "utilization.gpu,memory.used,memory.total,clocks.sm"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True,
                                 timeout=10).stdout
            print(f"[bench watchdog] nvidia-smi {q}: {out.strip()}", file=sys.stderr, flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[bench watchdog] post-mortem failed: {e}", file=sys.stderr, flush=True)


def dist_info():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return ws, rank, local


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own implementation of the path on the host cores. oracle/_ref holds the verbatim copy
# of koafusion/models + FocalLoss that oracle/build_ref.py makes in the build container (kind "reference"); without it
# the oracle port (fp32 eager restatement, pinned to the reference by the golden fixtures) is timed (kind "port").
# --------------------------------------------------------------------------------------------------
def cpu_reference_step(workload, knees, steps, warmup, threads, dropout=0.1):
    """Times zero_grad -> forward -> FocalLoss -> backward (koafusion/run/train_prog_fus.py:133-165) of `workload` on
    `knees` synthetic knees; returns (seconds per timed step, kind)."""
    from oracle import koa_oracle as ko

    torch.set_num_threads(threads)  # the reference's OMP_NUM_THREADS=1 is a DataLoader-worker setting (train_prog_fus.py:7-9)
    cfg = ko.make_config(workload, dropout=dropout)
    inputs, target = ko.make_inputs(workload, cfg, knees, 779)
    try:
        from oracle import ref_loader

        torch.manual_seed(778)
        model = ref_loader.build_model(workload, cfg)
        loss_fn = ref_loader.focal_loss(gamma=2)
        model.train()

        def one():
            model.zero_grad(set_to_none=True)
            out = model(*inputs)
            loss_fn(out["main"] if isinstance(out, dict) else out, target).backward()

        kind = "reference"
    except ImportError:
        spec = ko.model_param_spec(workload, cfg)
        sd = ko.make_state_dict(spec, 778)

        def one():
            ko.train_step(workload, cfg, sd, inputs, target)

        kind = "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, kind


def cpu_baseline_record(workload, knees, steps, warmup, dropout):
    threads = os.cpu_count() or 1
    times, kind = cpu_reference_step(workload, knees, steps, warmup, threads, dropout)
    med, best = statistics.median(times), min(times)
    what = ("the unmodified reference (oracle/_ref: koafusion.models + FocalLoss)" if kind == "reference" else
            "the fp32 eager PyTorch restatement of the reference (oracle port)")
    return dict(value=knees / med, unit="knees/s", cores=threads, kind=kind, median_s_per_step=med, min_s_per_step=best,
                best_value=knees / best, steps=len(times), warmup=warmup, knees_per_step=knees, workload=workload,
                sample=f"{warmup} warm-up + {len(times)} timed steps of zero_grad+forward+FocalLoss+backward on {knees} knee(s) "
                       f"of {workload} with {what}, {threads} threads; value = knees / median step time")


def run_reference(args):
    ws, rank, _ = dist_info()
    if rank != 0:
        return
    knees = args.cpu_knees
    warm = max(0, min(args.warmup, 1))
    rec = cpu_baseline_record(args.workload, knees, max(1, args.steps), warm, args.dropout)
    line = dict(metric=METRIC, value=rec["value"], unit="knees/s", impl="reference", n_gpus=args.gpus, steps=rec["steps"],
                warmup=warm, ms_per_step=1e3 * rec["median_s_per_step"], higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=args.workload, description=WORKLOAD_DESC.get(args.workload, args.workload),
                            knees_per_step=knees, device="host CPU", dropout=args.dropout),
                cpu_baseline={k: rec[k] for k in ("value", "unit", "cores", "kind", "sample", "median_s_per_step",
                                                  "min_s_per_step", "best_value")},
                e2e=dict(value=rec["value"], unit="knees/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def measure_full_step(step, model, dev_batches, steps, knees, peaks):
    """zero_grad + forward + FocalLoss + backward + Adam (koa_adam_step) on device-resident batches: knees/s of the full
    training step, and the optimiser update on its own against the HBM roofline (28 B per parameter element: parameter
    and both moments read and written, gradient read)."""
    from oaprogressionmmf_b200.optim import Adam

    opt = Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(2):
        step(*dev_batches[i % 2])
        opt.step()
    torch.cuda.synchronize()
    ev0.record()
    for i in range(steps):
        step(*dev_batches[i % 2])
        marks[i][0].record()
        opt.step()
        marks[i][1].record()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    adam_ms = sum(a.elapsed_time(b) for a, b in marks)
    n_live = sum(p.numel() for g in opt.param_groups for p in g["params"] if p.grad is not None)
    adam_ms /= steps
    gbs = 28.0 * n_live / (adam_ms * 1e-3) / 1e9 if adam_ms > 0 else None
    return dict(value=knees / (ms * 1e-3), unit="knees/s", ms_per_step=ms,
                step="zero_grad + forward + FocalLoss + backward + Adam(lr 1e-4, weight_decay 1e-4)",
                adam=dict(kernel="adam_kernel (multi-tensor, 64 tensors per launch)", ms=adam_ms, elements=n_live,
                          bound="hbm", achieved=gbs, peak=peaks["hbm"], unit="GB/s (28 B per element)",
                          frac=gbs / peaks["hbm"] if gbs else None))


def measure_inference(model, cfg, dev, batch=64, chunk=16, iters=3):
    """Eval-mode forward of `batch` knees in micro-batches of `chunk` (no coupling between knees in eval mode: BatchNorm
    uses its running statistics), inputs from pinned host memory, class predictions read back per micro-batch as the
    reference's eval loop does (koafusion/run/eval_prog_fus.py:286-304)."""
    from oaprogressionmmf_b200.synthetic import synthetic_batch

    was_training = model.training
    model.eval()
    try:
        torch.cuda.empty_cache()
        ins_h, _ = synthetic_batch(cfg, batch, 5, pin=True)

        from oaprogressionmmf_b200.evalpath import predict_batched

        def one():  # copy of micro-batch i + 1 overlaps the compute of micro-batch i; one read-back per batch
            return predict_batched(model, ins_h, dev, micro_batch=chunk)[0]

        with torch.no_grad():
            one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                pred = one()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return dict(value=batch / ms * 1e3, unit="knees/s", batch=batch, micro_batch=chunk, ms_per_batch=ms,
                    h2d_bytes_per_batch=sum(t.numel() * t.element_size() for t in ins_h), predictions=int(pred.numel()),
                    mode="eval, no_grad, H2D (pipelined per micro-batch on a copy stream) + prediction read-back inside the "
                         "timed region, through evalpath.predict_batched")
    finally:
        model.train(was_training)


def run_ours(args):
    import torch.distributed as dist

    from oaprogressionmmf_b200 import _lib
    from oaprogressionmmf_b200.koamodels import dict_models
    from oaprogressionmmf_b200.losses import FocalLoss
    from oaprogressionmmf_b200.synthetic import SyntheticKneeLoader, input_bytes, model_config, to_attr

    ws, rank, local = dist_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path. Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        # NCCL writes its version banner (level VERSION, included in WARN) to stdout; unless the caller asked for NCCL
        # logging, keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    torch.manual_seed(778)
    cfg = model_config(args.workload, dropout=args.dropout)
    model = dict_models[args.workload](to_attr(cfg), None).to(dev)
    model.train()
    n_params = sum(p.numel() for p in model.parameters())
    # knee-wise data parallelism: dead per-sequence heads frozen, one async all-reduce per engine call (dataparallel.py)
    from oaprogressionmmf_b200 import dataparallel

    step_model = dataparallel.wrap(model)
    loss_fn = FocalLoss(gamma=2)
    B = args.batch
    loader = SyntheticKneeLoader(cfg, B, seed=779 + rank, n_distinct=2, pin=True)
    host_batches = loader.batches
    dev_batches = [([t.to(dev) for t in ins], tgt.to(dev)) for ins, tgt in host_batches]
    h2d_bytes = input_bytes(*host_batches[0])

    def step(ins, tgt):
        step_model.zero_grad(set_to_none=True)
        logits = step_model(*ins)["main"]
        loss = loss_fn(logits, tgt)
        loss.backward()
        return loss

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------------------
    _phase("warm-up")
    for i in range(args.warmup):
        step(*dev_batches[i % 2])
    barrier()
    launches0 = lib.koa_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    _phase("timed steps")
    ev0.record()
    for i in range(args.steps):
        loss = step(*dev_batches[i % 2])
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.koa_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = ws * B * args.steps / (ms_total / 1e3)
    # from here on the measured line exists: whatever follows only adds fields to it (Watchdog)
    wd = Watchdog() if rank == 0 else None
    if wd is not None:
        peaks0 = measured_peaks()
        step_tf = value / ws * FLOPS_FWD_BWD[args.workload] / 1e12
        wd.arm(dict(metric=METRIC, value=value, unit="knees/s", n_gpus=ws, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=dict(workload=args.workload, knees_per_gpu=B, global_batch=B * ws),
                    roofline=dict(bound="tensor", kernel="whole training step", achieved=step_tf, peak=peaks0["tflops"],
                                  unit="TFLOP/s", frac=step_tf / peaks0["tflops"], traffic=None),
                    gpu_launches=int(launches), clocks=clocks),
               args.watchdog_seconds or (240 if args.no_cpu_baseline else 420))

    # ---- per-launch pass for the roofline: the modality branches run one after the other here (with concurrent
    # branches the CUDA events around a launch also cover the time it waits for SMs held by another branch's kernel),
    # every tcgen05 launch is bracketed by events on its stream. Same model, same batches, still inside a long step.
    from oaprogressionmmf_b200.koamodels import set_branch_streams

    _phase("per-launch roofline pass")
    prof_steps = 0 if args.no_roofline_pass else max(1, min(3, args.steps))
    set_branch_streams(False)
    if not args.skip_e2e:  # (profiler runs replay every launch: no extra warm-up step for them)
        step(*dev_batches[0])
    barrier()
    lib.koa_profile_enable(1)
    ev0.record()
    for i in range(prof_steps):
        step(*dev_batches[i % 2])
    ev1.record()
    barrier()
    ms_serial_total = ev0.elapsed_time(ev1)
    prof = (C.c_double * 6)()
    dump_path = args.profile_dump or os.path.join("/tmp", f"koa_shapes_{os.getpid()}.txt")
    if rank == 0:
        lib.koa_profile_dump(dump_path.encode())
    lib.koa_profile_read(prof)
    lib.koa_profile_enable(0)
    set_branch_streams(None)
    flag = _lib.debug_flag()

    # ---- end to end: pinned host inputs -> device copy -> step -> loss read back, every step ----------
    # (the loader issues the copy of the next batch on a copy stream while the current step computes; every batch is
    # copied from pinned host memory inside the timed region, the loss of every step is read back to the host)
    from oaprogressionmmf_b200.synthetic import DevicePrefetcher

    def e2e_step(it):
        ins, tgt = next(it)
        return float(step(ins, tgt).item())

    _phase("end-to-end pass")
    if args.skip_e2e:  # profiler runs only (ncu replays every launch): the line then carries no end-to-end number
        last_loss, e2e_value = float(loss.item()), None
        e2e_blocking, e2e_mode = None, None
    else:
        feed = DevicePrefetcher(loader, dev)
        for i in range(min(2, args.warmup)):
            e2e_step(feed)
        barrier()
        _phase("end-to-end pass: warm-up done")
        if args.hunt and wd is not None:
            _lib.debug_flag_peek()  # creates the watchdog's stream and pinned buffer while the device is idle
            wd.stall_seconds = 20
        for rep in range(max(1, args.e2e_repeat)):  # (> 1: soak runs of this leg only, tools/README.md)
            ev0.record()
            for i in range(args.steps):
                if args.hunt_events:  # every tcgen05 launch bracketed by events: koa_profile_pending names a stuck one
                    lib.koa_profile_enable(1)
                if args.hunt and wd is not None:
                    wd.beat()
                last_loss = e2e_step(feed)
            ev1.record()
            barrier()
            if args.e2e_repeat > 1:
                _phase(f"end-to-end pass: repeat {rep}: {ev0.elapsed_time(ev1) / args.steps:.1f} ms per step")
        if args.hunt:
            lib.koa_profile_enable(0)
            if wd is not None:
                wd.stall_seconds = None
        _phase("end-to-end pass: blocking read-back done")
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = ws * B * args.steps / (float(t.item()) / 1e3)
        e2e_blocking, e2e_mode = e2e_value, "blocking: loss.item() after every step, as the reference loop reads it"
        # The same pass with the loss read back the way trainpath.train_epoch leaves the device alone: the loss of step i is
        # copied to pinned host memory asynchronously and waited for while step i + 1 is already issued; all K losses are
        # on the host before the timer stops. Single GPU only (every rank would have to agree on a fallback); guarded: any
        # failure keeps the blocking number.
        if ws == 1:
            try:
                n_warm = min(2, args.warmup)
                host = torch.empty(args.steps + n_warm, dtype=torch.float32).pin_memory()
                marks = []

                def piped_step(i):
                    ins, tgt = next(feed)
                    host[i].copy_(step(ins, tgt).detach(), non_blocking=True)
                    e = torch.cuda.Event()
                    e.record()
                    marks.append(e)
                    if i > 0:
                        marks[i - 1].synchronize()          # the previous step's loss is on the host from here

                for i in range(n_warm):
                    piped_step(i)
                torch.cuda.synchronize()
                for rep in range(1 if args.hunt else max(1, args.e2e_repeat)):
                    ev0.record()
                    for i in range(n_warm, n_warm + args.steps):
                        piped_step(i)
                    marks[-1].synchronize()
                    ev1.record()
                    torch.cuda.synchronize()
                    if args.e2e_repeat > 1:
                        _phase(f"end-to-end pass: pipelined repeat {rep}: {ev0.elapsed_time(ev1) / args.steps:.1f} ms per step")
                        del marks[n_warm:]
                losses = host[n_warm:].tolist()
                if len(losses) == args.steps and all(v == v and abs(v) < 1e6 for v in losses):
                    e2e_value = B * args.steps / (ev0.elapsed_time(ev1) / 1e3)
                    last_loss = losses[-1]
                    e2e_mode = ("pipelined: the loss of step i reaches pinned host memory while step i+1 is issued "
                                "(trainpath.train_epoch); all losses on the host before the timer stops")
            except Exception as e:  # noqa: BLE001
                e2e_mode += f" (pipelined read-back failed: {type(e).__name__}: {e})"[:200]

    # ---- N > 1: what the gradient all-reduce costs. The same K steps with the synchronisation switched off (every rank
    # then trains on its own: timing only); exposed = time with the all-reduce - time without it.
    comm = None
    _phase("communication / inference legs")
    if ws > 1:
        from oaprogressionmmf_b200 import dataparallel as _dp

        _dp._state.enabled = False
        step(*dev_batches[0])
        barrier()
        ev0.record()
        for i in range(args.steps):
            step(*dev_batches[i % 2])
        ev1.record()
        barrier()
        _dp._state.enabled = True
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_nosync = float(t.item()) / args.steps
        n_live = sum(p.numel() for p in model.parameters() if p.requires_grad)
        comm = dict(exposed_ms=ms_step - ms_nosync, ms_per_step_without_allreduce=ms_nosync,
                    allreduce_bytes_per_step=4 * n_live, dtype="f32",
                    schedule="one async NCCL all-reduce per finished backward stage: fusion transformer, per-sequence "
                             "transformers, then layer4 / layer3 / layer2 / layer1+stem of every extractor")

    # ---- batched inference (BASELINE.json config 5), same model in eval mode under no_grad: 64 knees per GPU as two
    # micro-batches of 32, host -> device copy and the read-back of the predictions inside the timed region, no collective
    # (tools/infer_sweep.py is the full sweep). Every rank runs its replica; the slowest rank sets the time.
    inference = None
    if not args.skip_e2e and not args.no_full_step:
        try:
            inference = measure_inference(model, cfg, dev)
            if ws > 1:
                t = torch.tensor([inference["ms_per_batch"]], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                inference["ms_per_batch"] = float(t.item())
                inference["value"] = ws * inference["batch"] / float(t.item()) * 1e3
                inference["batch_per_gpu"] = inference["batch"]
                inference["batch"] = ws * inference["batch"]
                inference["mode"] += f"; {ws} replicas, no collective, time = slowest rank"
        except Exception as e:  # noqa: BLE001
            inference = dict(error=f"{type(e).__name__}: {e}"[:300])
            if ws > 1:
                dist.barrier()

    if rank != 0:
        if ws > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    flops_knee = FLOPS_FWD_BWD[args.workload]
    # dominant kernel family: the tcgen05 GEMM / implicit-GEMM convolution kernels (forward + data gradient),
    # timed per launch with CUDA events on the launching stream over the timed region
    k_ms, k_flops, k_n = prof[0], prof[1], prof[2]
    w_ms, w_flops, w_n = prof[3], prof[4], prof[5]
    achieved = (k_flops / (k_ms / 1e3)) / 1e12 if k_ms > 0 else 0.0
    # DRAM bytes per launch of the same kernel family from the committed ncu capture of this command (profiles/)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if args.workload == "XR1MR3C1CnnTrf" and B == 16 and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/r02_traffic.json (ncu dram__bytes_read+write.sum per launch of the GEMM family, same workload)"
    step_tflops = value / ws * flops_knee / 1e12
    # `frac` is the number the north star targets: algorithmic FLOPs of the WHOLE step (what the
    # reference computes, nothing this implementation adds) / step time / sustained bf16 peak, per GPU. The dominant kernel
    # family and the weight gradients follow with their own per-launch numbers (CUDA events around every launch,
    # algorithmic FLOPs per launch: zero-inserted rows, block-diagonal padding, K padding and the Gram bookkeeping GEMMs
    # are not counted).
    roofline = dict(bound="tensor", kernel="whole training step (forward + FocalLoss + backward), all kernels",
                    achieved=step_tflops, peak=peaks["tflops"], unit="TFLOP/s", frac=step_tflops / peaks["tflops"],
                    flops_per_knee=flops_knee, peak_source=peaks["source"], traffic=traffic, traffic_source=traffic_src,
                    gemm_family=dict(kernel="gemm_conv_kernel + gemm_kmajor_kernel (tcgen05 GEMM / im2col implicit-GEMM conv, "
                                            "forward + data gradient)",
                                     achieved=achieved, frac=achieved / peaks["tflops"], avg_launch_ms=k_ms / max(1.0, k_n),
                                     launches_timed=int(k_n), share_of_step=k_ms / max(ms_serial_total, 1e-9),
                                     dram_gbs=(traffic / (k_ms / max(1.0, k_n) * 1e-3) / 1e9) if traffic and k_ms > 0 else None),
                    wgrad=dict(kernel="gemm_wgrad_kernel (tcgen05 MN-major split-K)",
                               achieved=(w_flops / (w_ms / 1e3)) / 1e12 if w_ms > 0 else 0.0,
                               avg_launch_ms=w_ms / max(1.0, w_n), launches_timed=int(w_n),
                               share_of_step=w_ms / max(ms_serial_total, 1e-9)),
                    timing_pass=dict(steps=prof_steps, ms_per_step=ms_serial_total / max(1, prof_steps),
                                     note="per-launch CUDA events with the modality branches run one after the other; "
                                          "`value` is measured with the branches on concurrent streams"))
    roofline["by_bound"] = split_by_bound(dump_path, peaks, ms_serial_total)
    cpu = None
    _phase("cpu baseline")
    if ws == 1 and not args.no_cpu_baseline:
        # 1 warm-up + 3 timed steps, median and minimum, of the bench workload at 2 knees and of
        # BASELINE.json's config 1 (XR1Cnn, batch 8), on all host cores
        cpu = cpu_baseline_record(args.workload, args.cpu_knees, 3, 1, args.dropout)
        try:
            c1 = cpu_baseline_record("XR1Cnn", 8, 3, 1, args.dropout)
            cpu["config1"] = {k: c1[k] for k in ("value", "best_value", "median_s_per_step", "min_s_per_step", "kind", "sample")}
        except Exception as e:  # noqa: BLE001
            cpu["config1"] = dict(error=f"{type(e).__name__}: {e}"[:200])
    # ---- full training step: the same step followed by the optimiser update of the reference's training
    # configuration (Adam, lr 1e-4, weight decay 1e-4: conf/prog_fus.yaml:47-48) through koa_adam_step. Reported next to
    # `value`, never instead of it; single GPU only and last, so that nothing above depends on it.
    full_step = None
    _phase("cpu baseline done; full step")
    if ws == 1 and not args.skip_e2e and not args.no_full_step:
        try:
            full_step = measure_full_step(step, model, dev_batches, args.steps, B, peaks)
        except Exception as e:  # noqa: BLE001
            full_step = dict(error=f"{type(e).__name__}: {e}"[:300])
    line = dict(metric=METRIC, value=value, unit="knees/s", n_gpus=ws, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=args.workload, description=WORKLOAD_DESC.get(args.workload, args.workload),
                            knees_per_gpu=B, global_batch=B * ws, parallelism=f"dp{ws} (knee-wise, one process per GPU)",
                            streams="modality branches (XR, DESS, TSE, T2 extractor + per-sequence transformer) on 4 concurrent CUDA streams",
                            params=n_params, step="zero_grad + forward + FocalLoss + backward" +
                            (" + NCCL gradient all-reduce (one async collective per engine call, overlapped with backward)" if ws > 1 else ""),
                            dropout=args.dropout, bn="train mode (batch statistics per GPU)",
                            operand_formats="16-bit tensor-core operands, fp32 accumulation: CNN forward fp16, gradients and "
                                            "transformers bf16",
                            l2="working set per step (tens of GB of activations) exceeds the 126 MB L2; two input batches alternate"),
                roofline=roofline, cpu_baseline=cpu,
                e2e=dict(value=e2e_value, unit="knees/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=4,
                         loss_readback=e2e_mode, blocking_value=e2e_blocking),
                gpu_launches=int(launches), clocks=clocks, last_loss=last_loss, debug_flag=flag, full_step=full_step, comm=comm,
                inference=inference, switches={k: v for k, v in sorted(os.environ.items()) if k.startswith("KOA_")})
    if wd is not None:
        wd.disarm()
    print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="XR1MR3C1CnnTrf", choices=sorted(FLOPS_FWD_BWD))
    ap.add_argument("--batch", type=int, default=16, help="knees per GPU (runner.sh:342 trains the full model at 16)")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="fe.*.dropout / agg.emb_dropout / agg.mlp_dropout (authors' recipe: 0.1, runner.sh:352 + conf/model/*.yaml)")
    ap.add_argument("--cpu-knees", type=int, default=2, help="knees per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiler runs: skip the end-to-end pass")
    ap.add_argument("--no-full-step", action="store_true", help="skip the forward+backward+Adam measurement")
    ap.add_argument("--no-roofline-pass", action="store_true", help="profiler runs: skip the per-launch CUDA-event pass")
    ap.add_argument("--e2e-repeat", type=int, default=1, help="soak test: repeat the timed end-to-end loops this many times")
    ap.add_argument("--hunt", action="store_true",
                    help="soak test: the watchdog fires after 20 s without a finished step of the blocking end-to-end loop and "
                         "reports the barrier time-out codes")
    ap.add_argument("--hunt-events", action="store_true",
                    help="with --hunt: every tcgen05 launch bracketed by events, the watchdog also names the launches in flight")
    ap.add_argument("--watchdog-seconds", type=int, default=0, help="override the time budget of the legs after `value`")
    ap.add_argument("--profile-dump", default=None, help="write the per-shape tcgen05 kernel timing table to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
