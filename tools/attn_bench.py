"""Fused attention kernel alone + whole processor, per attn2 layer shape, for each kernel variant (CUDA events,
buffers rotated through > L2).  Usage: python tools/attn_bench.py [variants...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from photoverse_b200 import _lib  # noqa: E402

variants = [int(v) for v in sys.argv[1:]] or [2, 3]
rows = int(os.environ.get("PV_ROWS", "16"))
li = int(os.environ.get("PV_LI", "1"))
dev = torch.device("cuda:0")
for v in variants:
    _lib.set_option("attn_variant", v)
    _lib.set_option("attn3_wstat", int(os.environ.get("PV_WSTAT", "1")))
    r = bench.roofline_leg(dev, rows, li, 1.0)
    print(f"variant {v}: attn {r['achieved']} TFLOP/s frac {r['frac']}  processor {r['processor_tflops']} TFLOP/s "
          f"{r['processor_ms_per_unet_eval']} ms/eval")
    print("   ", json.dumps(r["per_shape_us"]))
