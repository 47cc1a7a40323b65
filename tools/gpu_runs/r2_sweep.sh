#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -m gpu -q -x -k "adapter or group_mean or golden" > gpurun_out/r2_adapter_tests.log 2>&1
echo "adapter tests rc=$?"; tail -3 gpurun_out/r2_adapter_tests.log
timeout 300 python tools/adapter_bench.py > gpurun_out/adapter_roofline_r02.json 2>/dev/null; echo "adapter bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/adapter_roofline_r02.json'))
for k,v in d['shapes'].items():
    print(k, {n: (r.get('GBps') or r.get('TFLOPs'), r.get('frac_of_hbm_peak') or r.get('frac_of_bf16_peak')) for n,r in v.items()})"
PV_SWEEP_OUT=gpurun_out/micro_sweep_r02.json timeout 1500 python tools/micro_sweep.py > gpurun_out/micro_sweep_r02.log 2>&1; echo "sweep rc=$?"
tail -3 gpurun_out/micro_sweep_r02.log; wc -l gpurun_out/micro_sweep_r02.log
