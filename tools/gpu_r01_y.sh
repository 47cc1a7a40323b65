#!/bin/bash
cd /root/repo
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -x -q -m gpu -k "roles or golden or backward or dropout" 2>&1 | tail -4
timeout 300 python tools/attn_bench.py 6 2>&1 | tail -3
PV_ATTN_VARIANT=6 PV_S=1024 PV_C=640 PV_TRACE_OUT=gpurun_out/trace_v6_B.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -1
