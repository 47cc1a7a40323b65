#!/bin/bash
# Round 2: parity of the single-launch processor (fused out projection) + per-layer timing, fused vs two launches.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "families or other_text or kv_cache or golden or full_size" > gpurun_out/r2_fuse_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_fuse_tests.log
tail -15 gpurun_out/r2_fuse_tests.log
timeout 300 python tools/attn_bench.py 1 0 > gpurun_out/r2_attn_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/r2_attn_bench.log | tail -12
