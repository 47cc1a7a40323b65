#!/bin/bash
cd /root/repo
timeout 200 python tools/attn_bench.py 6 2>&1 | tail -1 | cut -c1-240
PV_ATTN_VARIANT=6 PV_S=1024 PV_C=640 PV_TRACE_OUT=gpurun_out/trace_v6_B2.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -1
