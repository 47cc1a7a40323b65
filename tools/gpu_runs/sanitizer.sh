#!/bin/bash
# compute-sanitizer over the GPU kernel / backward tests (memcheck) and a racecheck pass over the adapter / SIMT kernels
cd /root/repo
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 600 \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -x -q -m gpu > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -x -q -m gpu -k "ln_lrelu or adapter_backward or inject or linear_f32 or pack_weight" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -6 gpurun_out/sanitizer_racecheck.log
