#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -x -k "variants or golden or full_size" > gpurun_out/pytest_d.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_d.log
timeout 600 python tools/attn_bench.py 2 3 2>&1 | tail -8
for cfg in "4096 320" "1024 640" "256 1280"; do
  set -- $cfg
  PV_S=$1 PV_C=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent -s 2 -c 1 \
     -f -o gpurun_out/prof_attn3_S$1 python tools/profile_one.py > gpurun_out/prof_attn3_S$1.log 2>&1
  echo "ncu S=$1 exit $?"
done
