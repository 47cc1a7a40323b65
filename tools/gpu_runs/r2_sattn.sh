#!/bin/bash
# attn1 self-attention kernel: parity + timing against torch SDPA.  Small shapes first so that a protocol bug traps early.
set -x
mkdir -p gpurun_out
PV_ROWS=2 PV_SHAPES="256:320,128:320,64:1280,200:640" timeout 120 python tools/sattn_bench.py 2>&1 | tail -8
for mode in ${PV_MODE_LIST:-3}; do
for pf in ${PV_POLY_LIST:-2}; do
  echo "mode $mode poly $pf"
  PV_MODE=$mode PV_POLY=$pf timeout 300 python tools/sattn_bench.py 2>&1 | tail -8
done
done
