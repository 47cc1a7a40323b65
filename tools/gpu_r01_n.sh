#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
PV_ATTN_VARIANT=4 PV_TRACE_OUT=gpurun_out/trace_v4_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -2
PV_ATTN_VARIANT=5 PV_TRACE_OUT=gpurun_out/trace_v5_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -2
PV_ATTN_VARIANT=5 PV_S=1024 PV_C=640 PV_TRACE_OUT=gpurun_out/trace_v5_B.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -2
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "other_text_lengths" 2>&1 | tail -8
