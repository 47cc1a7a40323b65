#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time, (2) --set full capture of the two
# native kernels of a processor call.  One GPU, never under torchrun.
R=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_layerstack_$R.csv python tools/profile_layer_stack.py > gpurun_out/layerstack_ncu_$R.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_attn_fwd -c 4 \
    -f -o gpurun_out/prof_attn_$R python tools/profile_layer_stack.py > gpurun_out/prof_attn_$R.log 2>&1
echo "attn full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 8 -c 4 \
    -f -o gpurun_out/prof_gemm_$R python tools/profile_layer_stack.py > gpurun_out/prof_gemm_$R.log 2>&1
echo "gemm full exit $?"
# one eager UNet evaluation: share of the native kernels in a denoising step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 1 --warmup 1 --denoise-steps 1 --no-graph \
    --no-cpu-baseline > gpurun_out/bench_under_ncu_$R.log 2>&1
echo "bench launch list exit $?"
ls -la gpurun_out | head -40
