#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -x -q -m gpu -s -k "training or unet or layer_norm" 2>&1 | tail -8
timeout 600 python bench.py --workload train --steps 10 --warmup 3 2> gpurun_out/backbone_train.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('train', d['value'], d['unit'], ' ms/step', d['ms_per_step'], 'launches', d.get('gpu_launches'), 'loss', d.get('final_loss'))
"
tail -3 gpurun_out/backbone_train.err
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
