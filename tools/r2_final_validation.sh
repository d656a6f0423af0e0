#!/bin/bash
# Round-end validation on a fresh box: the whole GPU suite, smoke(), the default bench line and the reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/final_bench.json
[ -n "$SKIP_REF" ] || timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"; tail -c 1200 gpurun_out/final_ref.json
