#!/bin/bash
# one kernel iteration: A/B timing of the attention kernel, parity of the processor path, device timeline
cd /root/repo
timeout 300 python tools/attn_ab.py ${1:-attn6_prefetch} ${2:-0} ${3:-1} 2>&1 | tail -6
timeout 500 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -4
bash tools/gpu_runs/r2_trace_attn.sh 4096 320 0 0 > /dev/null 2>&1
echo trace rc=$?
