#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -q --timeout 200 -x > gpurun_out/pytest_j.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_j.log
timeout 100 python tools/attn_bench.py 4 2>&1 | tail -2
timeout 100 python tools/gemm_bench.py 2>&1 | tail -5
