#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2f_$name.log 2>&1; echo "rc=$? $name $(grep -o '"value": [0-9.]*' gpurun_out/r2f_$name.log | head -1)" | tee -a gpurun_out/r2f_summary.log; }
: > gpurun_out/r2f_summary.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e --no-roofline-pass"
run attn_tc 120 python tools/attn_bench.py
KOA_ATTN_TC=0 run attn_old 120 python tools/attn_bench.py
run b_f16_tc 200 $B
KOA_FEAT_F16=0 run b_bf16_tc 200 $B
KOA_ATTN_TC=0 run b_f16_old 200 $B
KOA_FEAT_F16=0 KOA_ATTN_TC=0 run b_bf16_old 200 $B
run b_f16_tc2 200 $B
cat gpurun_out/r2f_summary.log; cat gpurun_out/r2f_attn_tc.log gpurun_out/r2f_attn_old.log
