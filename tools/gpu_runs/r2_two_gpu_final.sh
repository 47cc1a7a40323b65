#!/bin/bash
# 2 GPUs, last check of the round: torchrun parity test (bit-identical shards, allreduced grads == mean), the backbone
# kernel tests, and the default bench line at N = 2 (shorter training sub-records)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -q -x -k "group_norm_nhwc_matches" > gpurun_out/r2_gn_invariance.log 2>&1
echo "gn test rc=$?"; tail -3 gpurun_out/r2_gn_invariance.log
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_multi_test.log 2>&1
echo "multi test rc=$?"; tail -8 gpurun_out/r2_multi_test.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 3 --train-steps 10 > gpurun_out/bench_r02_n2_g.json 2> gpurun_out/bench_r02_n2_g.err
echo "bench n2 rc=$?"; cat gpurun_out/bench_r02_n2_g.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])
print('cfg3', d['cfg3'])
for k,v in d['train'].items(): print(k, {x: v[x] for x in ('value','ms_per_step','allreduce_exposed_ms','buckets','buckets_launched_under_backward_per_step','grad_elements')})
"; tail -5 gpurun_out/bench_r02_n2_g.err
