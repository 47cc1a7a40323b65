"""Timing experiments on the persistent attention kernel: pv_set_option('attn3_dbg', k) removes parts of the softmax
warps' work (results are wrong for k > 0) to expose which pipeline bounds the kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
g = torch.Generator().manual_seed(0)
ROWS, LI = 16, int(os.environ.get("PV_LI", "1"))
_lib.set_option("attn_variant", int(os.environ.get("PV_ATTN_VARIANT", "3")))
_lib.set_option("attn3_wstat", int(os.environ.get("PV_WSTAT", "1")))
_lib.set_option("attn3_prefetch", int(os.environ.get("PV_PF", "0")))
lib = _lib.lib()
SHAPES = [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]
if os.environ.get("PV_ONLY_A"):
    SHAPES = SHAPES[:1]
for S, C in SHAPES:
    text = torch.randn(ROWS, 77, 768, generator=g).to(dev, dt)
    img = torch.randn(ROWS, LI, 768, generator=g).to(dev, dt)
    wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    kv = ops.kv_pack(text, img, wkv, wkv, 8)
    nbuf = 6
    xs = [torch.randn(ROWS, S, C, device=dev, dtype=dt) for _ in range(nbuf)]
    os_ = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    res = []
    for dbg in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]:
        _lib.set_option("attn3_dbg", dbg % 10)
        _lib.set_option("attn3_stages", dbg // 10)

        def run(i):
            _lib.check(lib.pv_dual_attn_core_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp),
                                                 ops._ptr(os_[i]), None, ROWS, S, C, 8, 77, LI, 1.0, 1.0, ops._stream()))
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        # device time only: 20 launches captured in a CUDA graph (the Python/ctypes launch rate is ~15 us per call)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(20):
                run(i % nbuf)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        res.append((dbg, round(e0.elapsed_time(e1) * 1e3 / 20, 1)))
    _lib.set_option("attn3_dbg", 0)
    print(f"S={S} C={C}: " + "  ".join(f"dbg{d}={t}us" for d, t in res))
