#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "agrees_with_simt or deterministic" 2>&1 | tail -5
