#!/bin/bash
# Last validation of the round (no ncu): whole GPU suite, smoke, the default bench line, kernel lists of one UNet evaluation
set -u
T=${1:-g}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_$T.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_$T.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_$T.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_$T.json 2> gpurun_out/bench_r02_$T.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_r02_$T.json; tail -3 gpurun_out/bench_r02_$T.err
(echo "== stock PyTorch epilogues (bench.py --stock-epilogues) =="; timeout 300 python tools/unet_profile.py --stock-epilogues --top 30 2>&1 | grep -v Warn | grep -v _warn;
 echo; echo "== pv_backbone.cu epilogues (default) =="; timeout 300 python tools/unet_profile.py --top 30 2>&1 | grep -v Warn | grep -v _warn) > gpurun_out/unet_eval_kernels_r02.txt
echo "unet kernel lists rc=$?"
