#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of one warm-up + one timed step of the default bench workload.
# usage: bash tools/r2_ncu_launches.sh <tag>; the second half of the list is the timed step (tools/traffic_report.py --last-half)
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-full-step --no-roofline-pass > gpurun_out/ncu_$tag.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$tag.log | cut -c1-300; wc -l gpurun_out/launches_$tag.csv; gzip -f gpurun_out/launches_$tag.csv
