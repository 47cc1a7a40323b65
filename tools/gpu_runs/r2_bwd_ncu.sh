#!/bin/bash
# ncu --set full capture of the tcgen05 attention backward kernel (C = 320 layer, B = 16)
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_tc -c 1 -f -o gpurun_out/prof_bwdtc_r02 \
  python tools/bwd_bench.py > gpurun_out/prof_bwdtc_r02.log 2>&1
tail -2 gpurun_out/prof_bwdtc_r02.log
