#!/bin/bash
# Round 2, last sequence: whole GPU suite, smoke, both bench arms, then the ncu evidence of the final kernels
# (launch lists, --set full captures of the attention kernel, the out-projection GEMM and the backbone epilogues).
# Then (on CPU): python tools/summarize_profiles.py r02
set -u
T=${1:-f}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_$T.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_$T.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_$T.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r02_$T.json 2> gpurun_out/bench_ref_r02_$T.err; echo "reference arm rc=$?"
cut -c1-300 gpurun_out/bench_ref_r02_$T.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_$T.json 2> gpurun_out/bench_r02_$T.err; echo "bench rc=$?"
cat gpurun_out/bench_r02_$T.json; tail -5 gpurun_out/bench_r02_$T.err
R=r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_layerstack_$R.csv python tools/profile_layer_stack.py > gpurun_out/layerstack_ncu_$R.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 4 \
    -f -o gpurun_out/prof_attn_$R python tools/profile_layer_stack.py > gpurun_out/prof_attn_$R.log 2>&1
echo "attn full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm3_pair -c 3 \
    -f -o gpurun_out/prof_gemm_$R python tools/profile_layer_stack.py > gpurun_out/prof_gemm_$R.log 2>&1
echo "gemm full exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"gn_stats_nhwc|gn_apply_nhwc|layer_norm_bf16|geglu_kernel|add_bias_nhwc" -s 208 -c 16 \
    -f -o gpurun_out/prof_backbone_$R python tools/unet_profile.py --once > gpurun_out/prof_backbone_$R.log 2>&1
echo "backbone full exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 1 --warmup 1 --denoise-steps 1 --no-graph \
    --no-cpu-baseline --no-train-leg > gpurun_out/bench_under_ncu_$R.log 2>&1
echo "bench launch list exit $?"
