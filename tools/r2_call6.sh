#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2e_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2e_summary.log; tail -5 gpurun_out/r2e_$name.log | cut -c1-400; }
: > gpurun_out/r2e_summary.log
run attn 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_attention"
KOA_ATTN_LBO=8192 run attn_lbo8k 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_attention and 256"
run feat 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "feat or dropout or focal"
run full 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size or trained"
run all 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py::test_full_size_logits_match_reference --deselect tests/test_gpu_parity.py::test_trained_weights_logits_match_oracle
run bench 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e
KOA_ATTN_TC=0 run bench_noattn 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-step --skip-e2e
cat gpurun_out/r2e_summary.log
