#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2j_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2j_summary.log; tail -4 gpurun_out/r2j_$name.log | cut -c1-500; }
: > gpurun_out/r2j_summary.log
run new 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "exported or classifier or golden_case or eval_logits"
run full 1800 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size_logits"
run sweep 600 python tools/infer_sweep.py --batches 1,2,4,8,16,32,64,128 --chunk 16
run bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
cat gpurun_out/r2j_summary.log
