"""Profiling target: the 16 attn2 layer shapes of one SD-1.5 UNet evaluation at 16 rows (batch 8, uncond+cond),
each as the ONE native launch of a cached-K/V processor call (PV_FUSE_OUT=0: attention kernel + out-projection GEMM).
Run under ncu (see tools/gpu_profile.sh); prints CUDA-event timings when run plainly."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
ROWS = int(os.environ.get("PV_ROWS", "16"))
LI = int(os.environ.get("PV_LI", "1"))
REPS = int(os.environ.get("PV_REPS", "3"))
_lib.set_option("fuse_out", int(os.environ.get("PV_FUSE_OUT", "1")))
_lib.set_option("gemm_two_cta", int(os.environ.get("PV_GEMM_TWO_CTA", "1")))
SHAPES = [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16
layers = []
for S, C in SHAPES:
    text = torch.randn(ROWS, 77, 768, generator=g).to(dev, dt)
    img = torch.randn(ROWS, LI, 768, generator=g).to(dev, dt)
    wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    wkv_t = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    wkv_i = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    kv = ops.kv_pack(text, img, wkv_t, wkv_i, 8)
    x = torch.randn(ROWS, S, C, generator=g).to(dev, dt)
    layers.append((S, C, x, wq, kv, wo, torch.zeros(C, device=dev)))
torch.cuda.synchronize()
for rep in range(REPS):
    for S, C, x, wq, kv, wo, bo in layers:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y, o, _, _ = ops.dual_attn(x, wq, kv, wo, bo)
        e1.record()
        torch.cuda.synchronize()
        fl = 4 * ROWS * S * C * C + 4 * ROWS * S * C * (77 + LI)
        if rep == REPS - 1:
            print(f"S={S} C={C}: {e0.elapsed_time(e1) * 1e3:.1f} us  {fl / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s (attn + out-proj, cold L2 not enforced)")
print("launches", _lib.launch_count())
