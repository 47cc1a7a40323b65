#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -x -k "variants or golden or full_size or linearity or kv_cache" > gpurun_out/pytest_h.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_h.log
timeout 300 python tools/attn_dbg.py 0 1 3 2>&1 | tail -5
timeout 300 python tools/attn_bench.py 3 2>&1 | tail -3
