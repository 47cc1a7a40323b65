#!/bin/bash
# First GPU bring-up: each kernel family in its own process (a trap poisons the CUDA context), bounded by timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for sec in gemm attn_float32 attn_bfloat16; do
  timeout 300 python tools/gpu_diag.py $sec > gpurun_out/diag_$sec.log 2>&1
  echo "diag $sec exit $?" | tee -a gpurun_out/summary.txt
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/pytest_gpu.log
for sec in gemm attn_float32 attn_bfloat16; do echo "== $sec"; grep -v Traceback gpurun_out/diag_$sec.log | tail -25; done
