#!/bin/bash
# Device timeline of one CTA pair of the fused processor kernel (trace build).  usage: r2_trace.sh [S C [block]]
set -u
mkdir -p gpurun_out
export PV_LIB_PATH=$PWD/photoverse_b200/libphotoverse_b200_trace.so
S=${1:-4096}; C=${2:-320}; BLK=${3:-0}
PV_S=$S PV_C=$C PV_BLOCK=$BLK PV_FUSED=${PV_FUSED:-1} PV_NEV=400 PV_EVENTS=${PV_EVENTS:-1,2,50,51,52,60,61,70,71,72,73,74,75,84,85} \
  PV_TRACE_OUT=gpurun_out/r2_trace_${S}_${C}_b${BLK}.json timeout 300 python tools/attn_trace.py > gpurun_out/r2_trace_${S}_${C}_b${BLK}.log 2>&1
echo "rc=$?"; tail -150 gpurun_out/r2_trace_${S}_${C}_b${BLK}.log
