#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -x -k "variants or golden or full_size" > gpurun_out/pytest_b.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_b.log
timeout 600 python tools/attn_bench.py 2 3 2>&1 | tail -8
