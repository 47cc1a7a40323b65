#!/bin/bash
# Round 2: whole GPU suite + smoke + default bench line.  usage: r2_full.sh [tag]
set -u
T=${1:-a}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_$T.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_$T.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke_$T.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_$T.json 2> gpurun_out/bench_r02_$T.err; echo "bench rc=$?"
cat gpurun_out/bench_r02_$T.json; tail -5 gpurun_out/bench_r02_$T.err
