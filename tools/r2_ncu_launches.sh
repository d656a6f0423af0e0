#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of the timed step of `bench.py --steps 1 --warmup 1` (default workload): most
# of the warm-up step is skipped, the timed step starts at the fourth-from-last pack_fe_weights_kernel (one per extractor),
# which is where tools/traffic_report.py cuts.   usage: bash tools/r2_ncu_launches.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip 1900 --launch-count 2700 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-full-step --no-roofline-pass > gpurun_out/ncu_$tag.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$tag.log | cut -c1-300; wc -l gpurun_out/launches_$tag.csv; gzip -f gpurun_out/launches_$tag.csv
