#!/bin/bash
cd /root/repo
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import torch, bench
from photoverse_b200 import _lib
dev = torch.device("cuda:0")
for pdl in (1, 0, 1):
    _lib.set_option("pdl", pdl)
    r = bench.roofline_leg(dev, 16, 1, 1.0)
    print("pdl", pdl, r["frac"], {k: v["attn_us"] for k, v in r["per_shape_us"].items()}, r["processor_ms_per_unet_eval"])
PY
