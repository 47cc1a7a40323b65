#!/bin/bash
# last check of the round: full GPU suite, smoke, default bench line
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
wc -l gpurun_out/bench_r01_final.json; cut -c1-200 gpurun_out/bench_r01_final.json
