"""Device timeline of one CTA of the self-attention kernel (issuer warp + warp 0 of both softmax groups).
Needs the trace build: PV_TRACE=1 python -m photoverse_b200.build --force ; PV_LIB_PATH=photoverse_b200/libphotoverse_b200_trace.so
env: PV_S, PV_C, PV_BLOCK, PV_POLY, PV_SKIP, PV_NEV"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
S, C, B, H = int(os.environ.get("PV_S", "4096")), int(os.environ.get("PV_C", "320")), 16, 8
_lib.set_option("trace_block", int(os.environ.get("PV_BLOCK", "0")))
for k in ("POLY",):
    if os.environ.get("PV_" + k):
        _lib.set_option("sattn_" + k.lower(), int(os.environ["PV_" + k]))
lib = _lib.lib()
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, S, 3 * C, generator=g).to(dev, torch.bfloat16)
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
for _ in range(2):
    ops.self_attn(q, k, v, H)
cap = 8192
buf = torch.zeros(8 + 3 * cap, device=dev, dtype=torch.int64)
_lib.check(lib.pv_debug_trace(ops._ptr(buf), cap))
ops.self_attn(q, k, v, H)
torch.cuda.synchronize()
_lib.check(lib.pv_debug_trace(None, 0))
h = buf.cpu().tolist()
per = cap // 8
ev = []
for r in range(3):
    base = 8 + r * per * 3
    for i in range(min(h[r], per)):
        ev.append((h[base + 3 * i + 2], r, h[base + 3 * i], h[base + 3 * i + 1]))
ev.sort()
t0 = ev[0][0]
names = {10: "QK0", 11: "QK1", 12: "PV0", 13: "PV1", 20: "wait s_full", 21: "s_full", 22: "S loaded", 23: "max done", 24: "pv_done ok",
         25: "token ok", 26: "exps done", 27: "p_ready", 28: "unit done"}
skip, nev = int(os.environ.get("PV_SKIP", "300")), int(os.environ.get("PV_NEV", "120"))
for t, r, e, i in ev[skip: skip + nev]:
    print(f"{t - t0:9d}  {'    ' * (3 * r)}{['MMA', 'WG0', 'WG1'][r]} {names.get(e, e)} {i}")
print("events", len(ev), "span", ev[-1][0] - t0)
