"""Per-pair life of one fused processor launch (trace build, trace_block = -1): start, end of the attention phase, first /
last out-projection tile ready, end.  env: PV_S, PV_C, PV_ROWS."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
S, C = int(os.environ.get("PV_S", "4096")), int(os.environ.get("PV_C", "320"))
ROWS, LI = int(os.environ.get("PV_ROWS", "16")), 1
g = torch.Generator().manual_seed(0)
lib = _lib.lib()
_lib.set_option("trace_block", -1)
_lib.set_option("fuse_out", 2)
text = torch.randn(ROWS, 77, 768, generator=g).to(dev, dt)
img = torch.randn(ROWS, LI, 768, generator=g).to(dev, dt)
wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
kv = ops.kv_pack(text, img, wkv, wkv, 8)
x = torch.randn(ROWS, S, C, device=dev, dtype=dt)
o, y = torch.empty_like(x), torch.empty_like(x)
bo = torch.zeros(C, device=dev)
sync = torch.zeros(int(lib.pv_dual_attn_sync_words(ROWS, S)), device=dev, dtype=torch.int32)


def run():
    _lib.check(lib.pv_dual_attn_fwd(1, ops._ptr(x), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(wo), ops._ptr(bo),
                                    ops._ptr(y), None, ops._ptr(o), None, ops._ptr(sync), ROWS, S, C, 8, 77, LI, 1.0, 1.0,
                                    ops._stream()))


for _ in range(3):
    run()
cap = 4096
buf = torch.zeros(8 + 3 * cap, device=dev, dtype=torch.int64)
_lib.check(lib.pv_debug_trace(ops._ptr(buf), cap))
run()
torch.cuda.synchronize()
_lib.check(lib.pv_debug_trace(None, 0))
h = buf.cpu().tolist()
rows = []
for pair in range(74):
    v = h[8 + pair * 8: 8 + pair * 8 + 8]
    if v[0] == 0:
        continue
    rows.append((pair, v))
t0 = min(v[0] for _, v in rows)
print("pair  start  attn_end  probe1  tile0_ready  last_ready  loads_done  end   (us since first start)")
for pair, v in rows:
    f = lambda k: f"{(v[k] - t0) / 1e3:7.2f}" if v[k] else "      -"
    print(f"{pair:3d} {f(0)} {f(1)} {f(5)} {f(2)} {f(6)} {f(3)} {f(4)}")
ends = [v[4] - t0 for _, v in rows if v[4]]
a_ends = [v[1] - t0 for _, v in rows if v[1]]
print(f"attention phase ends: min {min(a_ends) / 1e3:.2f} max {max(a_ends) / 1e3:.2f} us;  kernel ends: min {min(ends) / 1e3:.2f} max {max(ends) / 1e3:.2f} us")
