#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2d_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2d_summary.log; tail -4 gpurun_out/r2d_$name.log | cut -c1-300; }
: > gpurun_out/r2d_summary.log
export PYTHONFAULTHANDLER=1
run bench_skip 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e
run bench_e2e -s INT 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-step
run all 1200 python -m pytest tests -x -q -m gpu
cat gpurun_out/r2d_summary.log
