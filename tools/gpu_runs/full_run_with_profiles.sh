#!/bin/bash
# Full run: GPU tests, smoke, bench (+ reference arm), ncu evidence of the final kernels.
cd /root/repo
T=${1:-n}        # run tag: outputs gpurun_out/bench_r01_$T.json, *_r01$T.* ; then python tools/summarize_profiles.py r01$T
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_$T.json 2> gpurun_out/bench_err.log; echo "bench exit $?"
cat gpurun_out/bench_r01_$T.json; tail -3 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r01_$T.json 2>> gpurun_out/bench_err.log; echo "ref exit $?"
cat gpurun_out/bench_ref_r01_$T.json
R=r01$T
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_layerstack_$R.csv python tools/profile_layer_stack.py > gpurun_out/layerstack_ncu_$R.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 4 \
    -f -o gpurun_out/prof_attn_$R python tools/profile_layer_stack.py > gpurun_out/prof_attn_$R.log 2>&1
echo "attn full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm3_pair -c 4 \
    -f -o gpurun_out/prof_gemm_$R python tools/profile_layer_stack.py > gpurun_out/prof_gemm_$R.log 2>&1
echo "gemm full exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 1 --warmup 1 --denoise-steps 1 --no-graph \
    --no-cpu-baseline > gpurun_out/bench_under_ncu_$R.log 2>&1
echo "bench launch list exit $?"
