#!/bin/bash
# Round-1 re-entry check: GPU tests, smoke, layer-stack timing for both kernel variants, bench, ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
echo "== variant 1"; PV_ATTN_VARIANT=1 timeout 300 python tools/profile_layer_stack.py 2>&1 | tail -5
echo "== variant 2"; PV_ATTN_VARIANT=2 timeout 300 python tools/profile_layer_stack.py 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
cat gpurun_out/bench_r01_a.json; tail -3 gpurun_out/bench_err.log
bash tools/gpu_profile.sh r01a
