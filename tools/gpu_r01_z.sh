#!/bin/bash
cd /root/repo
PV_ATTN_VARIANT=6 PV_TRACE_OUT=gpurun_out/trace_v6_pro.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -1
