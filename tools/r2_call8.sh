#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/r2g_$name.log 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r2g_summary.log; tail -4 gpurun_out/r2g_$name.log | cut -c1-600; }
: > gpurun_out/r2g_summary.log
run stages 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "in_stages or epoch_loops"
run full 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size"
run all 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py::test_full_size_logits_match_reference
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench 900 python bench.py --steps 10 --warmup 3
cat gpurun_out/r2g_summary.log
