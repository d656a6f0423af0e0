#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2l_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2l_summary.log; tail -3 gpurun_out/r2l_$name.log | cut -c1-300; }
: > gpurun_out/r2l_summary.log
run tests 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "maxpool or test_fe_ or model_train or model_eval"
run bench 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e --profile-dump gpurun_out/shapes_r2l.txt
cat gpurun_out/r2l_summary.log
