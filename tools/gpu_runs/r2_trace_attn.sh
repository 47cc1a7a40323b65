#!/bin/bash
# Device timeline of the attention phase of one CTA pair (trace build, all events).  usage: r2_trace_attn.sh [S C [block [skip]]]
set -u
mkdir -p gpurun_out
export PV_LIB_PATH=$PWD/photoverse_b200/libphotoverse_b200_trace.so
S=${1:-4096}; C=${2:-320}; BLK=${3:-0}; SKIP=${4:-150}
PV_S=$S PV_C=$C PV_BLOCK=$BLK PV_FUSED=0 PV_NEV=260 PV_SKIP=$SKIP \
  PV_TRACE_OUT=gpurun_out/r2_trace_attn_${S}_${C}_b${BLK}.json timeout 300 python tools/attn_trace.py > gpurun_out/r2_trace_attn_${S}_${C}_b${BLK}.log 2>&1
echo "rc=$?"; tail -262 gpurun_out/r2_trace_attn_${S}_${C}_b${BLK}.log
