#!/bin/bash
# `ncu --set full` captures of the two dominant HBM-bound launches of the round-2 step, each after the same command exited 0
# without the profiler: (1) the last 1x1 convolution of a layer-1 bottleneck with BatchNorm + identity + ReLU in the epilogue
# (gemm_conv_kernel MODE 2), (2) the K-concatenated data gradient of the same tail.   usage: bash tools/r2_ncu_full.sh
mkdir -p gpurun_out
run() {  # name M N K
  timeout 300 python tools/prof_gemm.py $2 $3 $4 $1 > gpurun_out/prof_$1.log 2>&1 || { echo "plain run of $1 failed"; tail -5 gpurun_out/prof_$1.log; return 1; }
  cat gpurun_out/prof_$1.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_conv_kernel --launch-skip 2 --launch-count 1 \
    -f -o gpurun_out/r02_ncu_$1 python tools/prof_gemm.py $2 $3 $4 $1 > gpurun_out/ncu_full_$1.log 2>&1
  echo "ncu $1 rc=$?"; ls -la gpurun_out/r02_ncu_$1.ncu-rep
}
run bnapply 1638400 256 64
run kcat 1638400 64 256
