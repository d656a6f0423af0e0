#!/bin/bash
# Round 2, GPU call 2: first run of the y-free bottleneck tail (KOA_BN_GRAM) and the fused inference epilogues.
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2b_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2b_summary.log; tail -4 gpurun_out/r2b_$name.log | cut -c1-400; }
: > gpurun_out/r2b_summary.log
run ops 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bn_apply or kcat or gemm_plain or fused_epilogues or conv_backward_epilogue"
run engines 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_fe_"
run rest 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py::test_fe_eval_features
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e"
run bench_gram 400 $B
KOA_BN_GRAM=0 run bench_nogram 400 $B
cat gpurun_out/r2b_summary.log
