#!/bin/bash
# training-time GroupNorm (+SiLU) kernels: parity tests, then the config[3] training step with the stock ops and with them
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -x -q -m gpu -s 2>&1 | tail -8
for v in "--stock-epilogues" ""; do
  timeout 600 python bench.py --workload train --steps 10 --warmup 3 $v 2> gpurun_out/backbone_train.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('train $v', d['value'], d['unit'], ' ms/step', d['ms_per_step'], 'launches', d.get('gpu_launches'), 'loss', d.get('final_loss'))
"
done
tail -3 gpurun_out/backbone_train.err
timeout 300 python tools/unet_profile.py --top 12 2>&1 | tail -14
