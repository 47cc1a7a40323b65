#!/bin/bash
# Round 2: parity of the single-launch processor (fused out projection) + per-layer timing, fused vs two launches.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2_fuse_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_fuse_tests.log
tail -6 gpurun_out/r2_fuse_tests.log
timeout 300 python tools/attn_bench.py 1 0 > gpurun_out/r2_attn_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/r2_attn_bench.log | tail -12
timeout 300 python tools/gemm_bench.py > gpurun_out/r2_gemm_bench.log 2>&1
tail -6 gpurun_out/r2_gemm_bench.log
