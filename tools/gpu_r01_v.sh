#!/bin/bash
cd /root/repo
PV_ATTN_VARIANT=6 PV_ONLY_A=1 timeout 200 python tools/attn_dbg.py 0 2 1 2>&1 | tail -3
PV_ATTN_VARIANT=6 PV_DBG=2 PV_TRACE_OUT=gpurun_out/trace_v6s_A.json PV_NEV=5 timeout 120 python tools/attn_trace.py | tail -1
