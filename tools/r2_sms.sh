#!/bin/bash
# experiment: smaller persistent grids (all GEMMs: KOA_NUM_SMS; HBM-bound convolution launches only: KOA_HBM_GRID_SMS)
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/r2m_$name.log 2>&1; echo "rc=$? $name $(grep -o '"value": [0-9.]*' gpurun_out/r2m_$name.log | head -1)" | tee -a gpurun_out/r2m_summary.log; }
: > gpurun_out/r2m_summary.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e --no-roofline-pass"
run base 200 $B
for n in 128 111 96; do KOA_NUM_SMS=$n run num_$n 200 $B; done
for n in 128 111 96 74; do KOA_HBM_GRID_SMS=$n run hbm_$n 200 $B; done
run base2 200 $B
run tests 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "in_stages or maxpool or feat_forward or test_fe_"
cat gpurun_out/r2m_summary.log
