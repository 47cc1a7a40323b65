#!/bin/bash
cd /root/repo
for blk in 0 2 72 74 146; do
PV_ATTN_VARIANT=6 PV_DBG=$blk PV_TRACE_OUT=gpurun_out/trace_v6_blk$blk.json PV_NEV=3 timeout 120 python tools/attn_trace.py | tail -1
done
