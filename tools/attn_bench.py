"""Fused attention kernel alone + whole processor, per attn2 layer shape, with the out projection fused into the
attention launch (1) and as a separate GEMM launch (0) (CUDA events, buffers rotated through > L2).
Usage: python tools/attn_bench.py [fuse_out values...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from photoverse_b200 import _lib  # noqa: E402

variants = [int(v) for v in sys.argv[1:]] or [2, 1, 0]
rows = int(os.environ.get("PV_ROWS", "16"))
li = int(os.environ.get("PV_LI", "1"))
dev = torch.device("cuda:0")
for v in variants:
    _lib.set_option("fuse_out", v)
    r = bench.roofline_leg(dev, rows, li, 1.0)
    print(f"fuse_out {v}: attn {r['achieved']} TFLOP/s frac {r['frac']}  processor {r['processor_tflops']} TFLOP/s "
          f"{r['processor_ms_per_unet_eval']} ms/eval")
    print("   ", json.dumps(r["per_shape_us"]))
