#!/bin/bash
# ncu --set full capture of the self-attention kernel at the S = 4096 / C = 320 shape (first launch, outside any graph)
mkdir -p gpurun_out
for pf in ${PV_POLY_LIST:-2}; do
PV_POLY=$pf PV_SHAPES="${PV_SHAPES:-4096:320}" timeout 600 ncu --set full --import-source on --clock-control none -k regex:self_attn_fwd -c 1 \
  -o gpurun_out/sattn_r02_pf$pf -f python tools/sattn_bench.py > gpurun_out/sattn_ncu_pf$pf.log 2>&1
tail -3 gpurun_out/sattn_ncu_pf$pf.log
done
