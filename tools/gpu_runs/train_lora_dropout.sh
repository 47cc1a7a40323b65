#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --workload train --lora-dropout 0.1 --steps 3 --warmup 3 > gpurun_out/bench_train_p01_final.json 2> gpurun_out/bench_train_p01_final.err; echo "exit $?"; cut -c1-330 gpurun_out/bench_train_p01_final.json
