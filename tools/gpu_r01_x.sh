#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
PV_ATTN_VARIANT=6 PV_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 1 -s 1 \
  -o gpurun_out/ncu_v6_A -f python tools/profile_one.py > gpurun_out/ncu_v6_A.log 2>&1
echo "ncu v6 exit $?"
