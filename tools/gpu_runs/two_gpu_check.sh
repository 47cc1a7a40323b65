#!/bin/bash
# 2-GPU check of the data-parallel paths: generation (no collective) and the training step (flat NCCL allreduce).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2_err.log
echo "gen n=2 exit $?"; cat gpurun_out/bench_n2.json | cut -c1-400; tail -3 gpurun_out/bench_n2_err.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2_err.log
echo "train n=2 exit $?"; cat gpurun_out/bench_train_n2.json; tail -3 gpurun_out/bench_train_n2_err.log
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.json 2>> gpurun_out/bench_train_n2_err.log
echo "train n=1 exit $?"; cat gpurun_out/bench_train_n1.json
