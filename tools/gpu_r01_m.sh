#!/bin/bash
# variant 5 (key-split softmax groups): parity + timing vs variant 4
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "cta-pair16" 2>&1 | tail -15
echo "pytest exit $?"
timeout 300 python tools/attn_bench.py 4 5 2>&1 | tail -12
