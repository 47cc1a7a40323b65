#!/bin/bash
mkdir -p gpurun_out
for cfg in "4096 320" "1024 640"; do
  set -- $cfg
  PV_S=$1 PV_C=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent -s 2 -c 1 \
     -f -o gpurun_out/prof_attn3_S$1 python tools/profile_one.py > gpurun_out/prof_attn3_S$1.log 2>&1
  echo "ncu S=$1 exit $?"
done
