#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -12
echo "pytest exit $?"
timeout 300 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_mma.json 2> gpurun_out/bench_train_mma.err; echo "train exit $?"; cut -c1-330 gpurun_out/bench_train_mma.json
