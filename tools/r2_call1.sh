#!/bin/bash
# Round 2, GPU call 1: decide the experiments written blind in round 1 (bench lines only; parity is run for winners later).
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2_summary.log; tail -c 600 gpurun_out/r2_$name.log | grep -o '"value": [0-9.]*' | head -1 | tee -a gpurun_out/r2_summary.log; }
: > gpurun_out/r2_summary.log
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e"
run bench_default 400 $B
run mixed_wgrad 200 python tools/try_mixed_wgrad.py
if grep -q "MIXED OK" gpurun_out/r2_mixed_wgrad.log; then
  KOA_WGRAD_XCVT=3 run bench_xcvt3 400 $B
fi
KOA_WGRAD_BULK_RED=1 run bench_bulkred 400 $B
for lvl in 1 2 3; do KOA_PDL=$lvl run bench_pdl$lvl 400 $B; done
KOA_IDX32=1 run bench_idx32 400 $B
KOA_BRANCH_PRIORITY=1 run bench_prio 400 $B
run bench_default2 400 $B
cat gpurun_out/r2_summary.log
