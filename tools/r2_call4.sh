#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2c_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2c_summary.log; tail -4 gpurun_out/r2c_$name.log | cut -c1-300; }
: > gpurun_out/r2c_summary.log
run new 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dropout or focal or cs or rs or nogap or linear_small or feat_forward"
run all 1200 python -m pytest tests -x -q -m gpu
run bench 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-step --profile-dump gpurun_out/shapes_r2c.txt
cat gpurun_out/r2c_summary.log
