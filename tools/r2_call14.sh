#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2k_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2k_summary.log; tail -3 gpurun_out/r2k_$name.log | cut -c1-400; }
: > gpurun_out/r2k_summary.log
run sweep 400 python tools/infer_sweep.py --batches 1,2,4,8,16,32,64,128,256 --chunk 16
run sweep32 200 python tools/infer_sweep.py --batches 32,64,128 --chunk 32
run bench 600 python bench.py --steps 10 --warmup 3
run steprows 400 python -m pytest tests/test_zz_gpu_step_rows.py -q -m gpu
cat gpurun_out/r2k_summary.log; cat gpurun_out/r2k_sweep.log gpurun_out/r2k_sweep32.log | cut -c1-200; grep "\[bench" gpurun_out/r2k_bench.log
