#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_g.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
cat gpurun_out/bench_r01_g.json; tail -3 gpurun_out/bench_err.log
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_r01_g.json 2> gpurun_out/bench_train_err.log; echo "train bench exit $?"
cat gpurun_out/bench_train_r01_g.json; tail -5 gpurun_out/bench_train_err.log
