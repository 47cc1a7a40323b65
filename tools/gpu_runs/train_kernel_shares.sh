#!/bin/bash
# kernel shares of one training step (config[3]) under ncu (serialised, cold-cache: compare shares only)
cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/launches_bench_r01train.csv python bench.py --workload train --steps 1 --warmup 1 \
    > gpurun_out/bench_train_under_ncu.log 2>&1
echo "train launch list exit $?"; tail -2 gpurun_out/bench_train_under_ncu.log | cut -c1-300
