#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 300 python tools/profile_layer_stack.py 2>&1 | tail -8
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_first.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
cat gpurun_out/bench_r01_first.json; tail -5 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_first.json 2>> gpurun_out/bench_err.log; echo "ref exit $?"
cat gpurun_out/bench_ref_first.json
bash tools/gpu_profile.sh r01
