#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "variants or linear_bf16" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest(variants) exit $?"; tail -12 gpurun_out/pytest_gpu.log
echo "== variant 1, gemm 1cta"; PV_ATTN_VARIANT=1 PV_GEMM_TWO_CTA=0 timeout 300 python tools/profile_layer_stack.py 2>&1 | tail -5
echo "== variant 2, gemm 2cta"; PV_ATTN_VARIANT=2 PV_GEMM_TWO_CTA=1 timeout 300 python tools/profile_layer_stack.py 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest(all) exit $?"; tail -5 gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_v2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_shape_us'], d['roofline']['processor_tflops'], d['roofline']['processor_share_of_step'])"
tail -3 gpurun_out/bench_err.log
