#!/bin/bash
# fused GroupNorm(+SiLU) / GEGLU epilogues of the host UNet: parity tests + images/s with and without them
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -x -q -m gpu -s 2>&1 | tail -12
for v in "--stock-epilogues" ""; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-leg $v 2> gpurun_out/backbone_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', d['value'], 'images/s  e2e', d['e2e']['value'], ' ms/step', d['ms_per_step'], ' native/eval', d['native_kernels_per_unet_eval'], d['clocks'])
"
done
tail -3 gpurun_out/backbone_ab.err
