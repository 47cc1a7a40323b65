#!/bin/bash
# LoRA-dropout training path: parity tests + train bench with p = 0 / 0.1
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -15
echo "pytest exit $?"
timeout 300 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_p0_l.json 2> gpurun_out/bench_train_p0_l.err; echo "train p0 exit $?"; cat gpurun_out/bench_train_p0_l.json
timeout 300 python bench.py --workload train --lora-dropout 0.1 --steps 3 --warmup 3 > gpurun_out/bench_train_p01_l.json 2> gpurun_out/bench_train_p01_l.err; echo "train p0.1 exit $?"; cat gpurun_out/bench_train_p01_l.json; tail -5 gpurun_out/bench_train_p01_l.err
