#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "cta-pair16" 2>&1 | tail -15
echo "pytest exit $?"
timeout 300 python tools/attn_bench.py 4 5 2>&1 | tail -12
PV_ATTN_VARIANT=5 PV_TRACE_OUT=gpurun_out/trace_v5p_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -2
