#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget was spent, in one box visit.
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/round2_first_call.sh'   # ~30-40 min of box time
# Results land in gpurun_out/r2_*.log. Each step has its own timeout and never stops the next one.
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2_summary.log; tail -3 gpurun_out/r2_$name.log; }
: > gpurun_out/r2_summary.log
# 1. the rows next to the hot path, first run on hardware (Adam, resampling, augmentation, predictions)
run step_rows 600 python -m pytest tests/test_zz_gpu_step_rows.py -q -m gpu
# 2. the whole GPU suite
run gpu_suite 900 python -m pytest tests -x -q -m gpu
# 3. baseline bench of the current defaults
run bench_default 600 python bench.py --steps 5 --warmup 3
# 4. experiments written blind (own processes: a rejected instruction must not take anything else down)
run mixed_wgrad 300 python tools/try_mixed_wgrad.py
if grep -q "MIXED OK" gpurun_out/r2_mixed_wgrad.log; then
  KOA_WGRAD_XCVT=3 run bench_xcvt3 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-step
  KOA_WGRAD_XCVT=3 run parity_xcvt3 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu
fi
KOA_WGRAD_BULK_RED=1 run parity_bulkred 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu
KOA_WGRAD_BULK_RED=1 run bench_bulkred 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-step
KOA_PDL=3 run parity_pdl3 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu
for lvl in 1 2 3; do
  KOA_PDL=$lvl run bench_pdl$lvl 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-step
done
KOA_IDX32=1 run parity_idx32 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu
KOA_IDX32=1 run bench_idx32 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-step
KOA_BRANCH_PRIORITY=1 run bench_prio 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-step
run step_ops_bench 300 python tools/step_ops_bench.py
# 5. profiler passes last (numbers printed under ncu are never bench values): launch list + DRAM bytes of one timed step
#    (the warm-up step is skipped), summarised by tools/launch_report.py into profiles/ by hand afterwards
run ncu_launches 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip 1760 --launch-count 1760 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-full-step
cat gpurun_out/r2_summary.log
