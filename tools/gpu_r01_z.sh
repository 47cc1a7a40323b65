#!/bin/bash
cd /root/repo
timeout 300 python tools/attn_bench.py 6 2>&1 | tail -2
PV_ATTN_VARIANT=6 PV_TRACE_OUT=gpurun_out/trace_v6c_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -1
