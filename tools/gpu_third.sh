#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_layer_stack.py 2>&1 | tail -6
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cl.json')); print('channels_last', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_shape_us'], d['roofline']['processor_tflops'], d['roofline']['processor_share_of_step'])"
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nchw > gpurun_out/bench_nchw.json 2>> gpurun_out/bench_err.log; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_nchw.json')); print('nchw', d['value'], d['ms_per_step'])"
tail -3 gpurun_out/bench_err.log
