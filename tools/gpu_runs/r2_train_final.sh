#!/bin/bash
# Round 2: training step after the tcgen05 attention backward -- pv:: kernel list under ncu and the timed train records.
set -u
mkdir -p gpurun_out
R=r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"pv::" -c 4000 --csv \
    --log-file gpurun_out/launches_pv_train_$R.csv python bench.py --workload train --steps 1 --warmup 0 > gpurun_out/train_ncu_$R.log 2>&1
echo "train pv list exit $?"
timeout 600 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/bench_train_r02_d.json 2> gpurun_out/bench_train_r02_d.err; echo "train bench rc=$?"
cat gpurun_out/bench_train_r02_d.json | cut -c1-1200
timeout 300 ncu --set full --clock-control none -k regex:"attn_bwd_tc|attn_bwd_mma|lora_wgrad_mma" -c 4 \
    -f -o gpurun_out/prof_bwd_$R python tools/bwd_bench.py > gpurun_out/prof_bwd_$R.log 2>&1
echo "bwd full exit $?"
