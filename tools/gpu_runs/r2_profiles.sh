#!/bin/bash
# Round 2 evidence: ncu launch lists + `--set full` captures of the final kernels, adapter roofline, training pv:: kernels.
# Then (here, on CPU): python tools/summarize_profiles.py r02
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_layerstack_$R.csv python tools/profile_layer_stack.py > gpurun_out/layerstack_ncu_$R.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 4 \
    -f -o gpurun_out/prof_attn_$R python tools/profile_layer_stack.py > gpurun_out/prof_attn_$R.log 2>&1
echo "attn full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm3_pair -c 3 \
    -f -o gpurun_out/prof_gemm_$R python tools/profile_layer_stack.py > gpurun_out/prof_gemm_$R.log 2>&1
echo "gemm full exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"ln_lrelu_kernel|group_mean_kernel|gemm_bf16_tcgen05" -c 6 \
    -f -o gpurun_out/prof_adapter_$R python tools/adapter_bench.py --once > gpurun_out/prof_adapter_$R.log 2>&1
echo "adapter full exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"attn_bwd_mma|lora_wgrad_mma" -c 4 \
    -f -o gpurun_out/prof_bwd_$R python tools/bwd_bench.py > gpurun_out/prof_bwd_$R.log 2>&1
echo "bwd full exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 1 --warmup 1 --denoise-steps 1 --no-graph \
    --no-cpu-baseline --no-train-leg > gpurun_out/bench_under_ncu_$R.log 2>&1
echo "bench launch list exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"pv::" -c 4000 --csv \
    --log-file gpurun_out/launches_pv_train_$R.csv python bench.py --workload train --steps 1 --warmup 0 > gpurun_out/train_ncu_$R.log 2>&1
echo "train pv list exit $?"
timeout 300 python tools/adapter_bench.py > gpurun_out/adapter_roofline_$R.json 2> gpurun_out/adapter_roofline_$R.err; echo "adapter bench exit $?"
cat gpurun_out/adapter_roofline_$R.json | head -60
