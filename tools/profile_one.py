"""One attn2 layer shape, the fused attention kernel only (ncu target).  env: PV_S, PV_C, PV_ROWS, PV_LI"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
S, C = int(os.environ.get("PV_S", "4096")), int(os.environ.get("PV_C", "320"))
ROWS, LI = int(os.environ.get("PV_ROWS", "16")), int(os.environ.get("PV_LI", "1"))
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16
text = torch.randn(ROWS, 77, 768, generator=g).to(dev, dt)
img = torch.randn(ROWS, LI, 768, generator=g).to(dev, dt)
wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
wkv_t = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
wkv_i = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
kv = ops.kv_pack(text, img, wkv_t, wkv_i, 8)
x = torch.randn(ROWS, S, C, generator=g).to(dev, dt)
o = torch.empty_like(x)
lib = _lib.lib()
for _ in range(int(os.environ.get("PV_REPS", "3"))):
    _lib.check(lib.pv_dual_attn_core_fwd(1, ops._ptr(x), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(o), None,
                                         ROWS, S, C, 8, 77, LI, 1.0, 1.0, ops._stream()))
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
