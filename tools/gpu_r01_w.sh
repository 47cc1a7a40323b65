#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "pytest exit $?"
timeout 300 python tools/attn_bench.py 4 6 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_w.json 2> gpurun_out/bench_r01_w.err; echo "bench exit $?"; cat gpurun_out/bench_r01_w.json | cut -c1-600
