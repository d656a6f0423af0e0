#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29512 tools/dp_check.py > gpurun_out/r2i_dpcheck.log 2>&1; echo "rc=$? dpcheck"; tail -3 gpurun_out/r2i_dpcheck.log | cut -c1-400
timeout 600 $T --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_bench2.log 2>&1; echo "rc=$? bench2"; tail -1 gpurun_out/r2i_bench2.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench1.log 2>&1; echo "rc=$? bench1"; tail -1 gpurun_out/r2i_bench1.log | cut -c1-300
