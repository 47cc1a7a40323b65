#!/bin/bash
# Round 2, end of round: whole GPU suite, smoke, both bench arms, refreshed ncu captures of the processor, the
# out-projection GEMM and the attn1 kernel.  Then (on CPU): python tools/summarize_profiles.py r02
set -u
T=${1:-c}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_$T.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_$T.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_$T.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r02_$T.json 2> gpurun_out/bench_ref_r02_$T.err; echo "reference arm rc=$?"
cat gpurun_out/bench_ref_r02_$T.json | cut -c1-600
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_$T.json 2> gpurun_out/bench_r02_$T.err; echo "bench rc=$?"
cat gpurun_out/bench_r02_$T.json; tail -5 gpurun_out/bench_r02_$T.err
R=r02
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 4 \
    -f -o gpurun_out/prof_attn_$R python tools/profile_layer_stack.py > gpurun_out/prof_attn_$R.log 2>&1
echo "attn full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm3_pair -c 3 \
    -f -o gpurun_out/prof_gemm_$R python tools/profile_layer_stack.py > gpurun_out/prof_gemm_$R.log 2>&1
echo "gemm full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_attn_fwd -c 4 \
    -f -o gpurun_out/prof_sattn_$R python tools/sattn_bench.py > gpurun_out/prof_sattn_$R.log 2>&1
echo "sattn full exit $?"
