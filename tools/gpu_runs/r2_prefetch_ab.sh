#!/bin/bash
# A/B of the S-row prefetch in the head_dim 40 / 80 attention kernel + parity of the default path
cd /root/repo
timeout 300 python tools/attn_ab.py attn6_prefetch 0 1 2>&1 | tail -6
timeout 500 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -4
