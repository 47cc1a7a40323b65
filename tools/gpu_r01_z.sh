#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "full_size" 2>&1 | tail -4
PV_SWEEP_OUT=gpurun_out/micro_sweep_r01.json timeout 600 python tools/micro_sweep.py 2>&1 | tail -75
