#!/bin/bash
# Round 2: backward parity + timing.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py -m gpu -q -x > gpurun_out/r2_bwd_tests.log 2>&1
echo "bwd tests rc=$?"; tail -5 gpurun_out/r2_bwd_tests.log
timeout 600 python tools/bwd_bench.py > gpurun_out/r2_bwd_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r2_bwd_bench.log
timeout 900 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r2_train_n1.json 2> gpurun_out/r2_train_n1.err; echo "train rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_train_n1.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','steps')})"
