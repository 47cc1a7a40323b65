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
_lib.set_option("fuse_out", 2)
lib = _lib.lib()
text = torch.randn(ROWS, 77, 768, generator=g).to(dev, dt)
img = torch.randn(ROWS, LI, 768, generator=g).to(dev, dt)
wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
kv = ops.kv_pack(text, img, wkv, wkv, 8)
x = torch.randn(ROWS, S, C, device=dev, dtype=dt)
o = torch.empty_like(x)


FUSED = int(os.environ.get("PV_FUSED", "1"))
wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
bo = torch.zeros(C, device=dev)
y = torch.empty_like(x)
sync = torch.zeros(int(lib.pv_dual_attn_sync_words(ROWS, S)), device=dev, dtype=torch.int32)


def run():
    if FUSED:
        _lib.check(lib.pv_dual_attn_fwd(1, ops._ptr(x), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(wo), ops._ptr(bo),
                                        ops._ptr(y), None, ops._ptr(o), None, ops._ptr(sync), ROWS, S, C, 8, 77, LI, 1.0, 1.0,
                                        ops._stream()))
    else:
        _lib.check(lib.pv_dual_attn_core_fwd(1, ops._ptr(x), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(o), None,
                                             ROWS, S, C, 8, 77, LI, 1.0, 1.0, ops._stream()))


for _ in range(3):
    run()
cap = 4096
buf = torch.zeros(8 + 3 * cap, device=dev, dtype=torch.int64)
_lib.check(lib.pv_debug_trace(ops._ptr(buf), cap))
run()
torch.cuda.synchronize()
_lib.check(lib.pv_debug_trace(None, 0))
h = buf.cpu().tolist()
per = cap // 8
ev = []
for r in range(8):
    base = 8 + r * per * 3
    for k in range(min(h[r], per)):
        ev.append((h[base + 3 * k + 2], h[base + 3 * k], h[base + 3 * k + 1]))
ev.sort()
n = len(ev)
t0 = ev[0][0]
names = {1: "prologue cycles (entry -> cluster sync) =", 2: "entry -> after griddepcontrol.wait =", 12: " qp full", 13: " qp mma issued", 14: " qp committed", 22: "  qk begin", 23: "  qk mmas issued", 25: "  pv wait p_ready", 26: "  pv p_ready ok", 27: "  pv mmas issued", 10: "QP start", 11: "QP issued", 20: "  QK issued", 21: "  PV issued", 29: "    A wait q_full", 30: "    A q_full", 31: "    A conv done",
         32: "    A slot_free", 33: "    A head begin", 34: "    A p_ready", 35: "    A next S in registers", 36: "    A exps done", 37: "    A pack begin, prefetch =", 38: "    A drained",
         45: "        B next S in registers", 46: "        B exps done", 47: "        B pack begin, prefetch =", 48: "        B drained", 39: "        B wait q_full", 40: "        B q_full", 41: "        B conv done",
         42: "        B slot_free", 43: "        B head begin", 44: "        B p_ready",
         50: "P2 producer start", 51: "P2 tile ready", 52: "P2 tile loads issued", 53: "P2 tile posted", 60: "  P2 mma first stage full", 61: "  P2 mma tile issued",
         70: "    P2 epi acc_full", 73: "    P2 epi tmem loaded", 74: "    P2 epi staging free", 75: "    P2 epi packed", 71: "    P2 epi store issued", 72: "    P2 epi done",
         80: "            E drain begin", 81: "            E drain end", 82: "            E wait stores", 83: "            E stores done", 86: "            E signalled",
         84: "            E all drained", 85: "            E all signalled"}
only = os.environ.get("PV_EVENTS")
if only:
    keep = {int(v) for v in only.split(",")}
    ev = [e for e in ev if e[1] in keep]
if os.environ.get("PV_TRACE_OUT"):
    import json
    json.dump([(t - t0, e, i) for t, e, i in ev], open(os.environ["PV_TRACE_OUT"], "w"))
skip = int(os.environ.get("PV_SKIP", "0"))
for t, e, i in ev[skip: skip + int(os.environ.get("PV_NEV", "140"))]:
    print(f"{t - t0:8d}  {names.get(e, e)} {i}")
print("events", n, "span cycles", ev[-1][0] - t0)
