#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
PV_ATTN_VARIANT=5 PV_TRACE_OUT=gpurun_out/trace_v5q_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -2
