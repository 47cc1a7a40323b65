#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for v in 4 5; do
PV_ATTN_VARIANT=$v PV_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 1 -s 1 \
  -o gpurun_out/ncu_v${v}_A -f python tools/profile_one.py > gpurun_out/ncu_v${v}_A.log 2>&1
echo "ncu v$v exit $?"
done
ls -la gpurun_out/*.ncu-rep
