"""Timeline of one CTA of the persistent attention kernels (debug trace events) for one layer shape.
Needs a trace build: PV_TRACE=1 python -m photoverse_b200.build --force.  env PV_BLOCK selects the leader CTA."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
S, C = int(os.environ.get("PV_S", "4096")), int(os.environ.get("PV_C", "320"))
ROWS, LI = 16, 1
g = torch.Generator().manual_seed(0)
_lib.set_option("trace_block", int(os.environ.get("PV_BLOCK", "0")))
lib = _lib.lib()
text = torch.randn(ROWS, 77, 768, generator=g).to(dev, dt)
img = torch.randn(ROWS, LI, 768, generator=g).to(dev, dt)
wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
kv = ops.kv_pack(text, img, wkv, wkv, 8)
x = torch.randn(ROWS, S, C, device=dev, dtype=dt)
o = torch.empty_like(x)


def run():
    _lib.check(lib.pv_dual_attn_core_fwd(1, ops._ptr(x), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(o), None,
                                         ROWS, S, C, 8, 77, LI, 1.0, 1.0, ops._stream()))


for _ in range(3):
    run()
cap = 4096
buf = torch.zeros(4 + 3 * cap, device=dev, dtype=torch.int64)
_lib.check(lib.pv_debug_trace(ops._ptr(buf), cap))
run()
torch.cuda.synchronize()
_lib.check(lib.pv_debug_trace(None, 0))
h = buf.cpu().tolist()
per = cap // 4
ev = []
for r in range(4):
    base = 4 + r * per * 3
    for k in range(min(h[r], per)):
        ev.append((h[base + 3 * k + 2], h[base + 3 * k], h[base + 3 * k + 1]))
ev.sort()
n = len(ev)
t0 = ev[0][0]
names = {1: "prologue cycles (entry -> cluster sync) =", 2: "entry -> after griddepcontrol.wait =", 12: " qp full", 13: " qp mma issued", 14: " qp committed", 22: "  qk begin", 23: "  qk mmas issued", 25: "  pv wait p_ready", 26: "  pv p_ready ok", 27: "  pv mmas issued", 10: "QP start", 11: "QP issued", 20: "  QK issued", 21: "  PV issued", 29: "    A wait q_full", 30: "    A q_full", 31: "    A conv done",
         32: "    A slot_free", 33: "    A s_full", 34: "    A p_ready", 35: "    A S loaded", 36: "    A exps done", 37: "    A exchanged", 38: "    A drained",
         45: "        B S loaded", 46: "        B exps done", 47: "        B exchanged", 48: "        B drained", 39: "        B wait q_full", 40: "        B q_full", 41: "        B conv done",
         42: "        B slot_free", 43: "        B s_full", 44: "        B p_ready"}
if os.environ.get("PV_TRACE_OUT"):
    import json
    json.dump([(t - t0, e, i) for t, e, i in ev], open(os.environ["PV_TRACE_OUT"], "w"))
skip = int(os.environ.get("PV_SKIP", "0"))
for t, e, i in ev[skip: skip + int(os.environ.get("PV_NEV", "140"))]:
    print(f"{t - t0:8d}  {names.get(e, e)} {i}")
print("events", n, "span cycles", ev[-1][0] - t0)
