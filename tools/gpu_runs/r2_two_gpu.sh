#!/bin/bash
# Round 2, 2 GPUs: the torchrun parity test (bit-identical shards, allreduced grads == mean) and the default bench line
# at N = 2 (generation weak scaling + cfg3 strong scaling + training sub-records with the overlapped NCCL allreduce).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_multi_test.log 2>&1
echo "multi test rc=$?"; tail -8 gpurun_out/r2_multi_test.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "bench n2 rc=$?"; cat gpurun_out/bench_r02_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
print('cfg3', d['cfg3'])
for k,v in d['train'].items(): print(k, {x: v[x] for x in ('value','ms_per_step','allreduce_exposed_ms','buckets','buckets_launched_under_backward_per_step','grad_elements')})
"; tail -5 gpurun_out/bench_r02_n2.err
