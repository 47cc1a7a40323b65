#!/bin/bash
# parity of the default attention path + its timing per layer shape
cd /root/repo
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -x -q -m gpu -k "roles or golden or backward or dropout or full_size" 2>&1 | tail -4
timeout 300 python tools/attn_bench.py 6 2>&1 | tail -3
