#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2h_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2h_summary.log; tail -4 gpurun_out/r2h_$name.log | cut -c1-500; }
: > gpurun_out/r2h_summary.log
run diag 300 python tools/diag_r2.py stages
run stages 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "in_stages or teacher or floor"
run full 1800 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size_logits"
cat gpurun_out/r2h_summary.log; head -12 gpurun_out/r2h_diag.log
