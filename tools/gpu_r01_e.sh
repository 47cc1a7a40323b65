#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -x -k "variants or golden or full_size or linearity" > gpurun_out/pytest_e.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_e.log
timeout 300 python tools/attn_dbg.py 0 1 2 3 4 2>&1 | tail -5
