#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "variants or golden or tall" 2>&1 | tail -5
timeout 300 python tools/attn_bench.py 3 4 5 2>&1 | tail -12
timeout 200 python tools/gemm_bench.py 2>&1 | tail -8
