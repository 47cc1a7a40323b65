#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roles" 2>&1 | tail -8
timeout 300 python tools/attn_bench.py 4 6 2>&1 | tail -8
PV_ATTN_VARIANT=6 PV_TRACE_OUT=gpurun_out/trace_v6_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -2
