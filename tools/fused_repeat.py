"""Stress the single-launch processor kernel: 12 back-to-back calls per shape with rotating buffers (and, with SYNC_EACH=1,\na device synchronisation after every call); the row-block counters must be zero afterwards.  env: SHAPES=BxSxC,..., ROT=0|1"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from photoverse_b200 import _lib, ops
dev = torch.device("cuda:0"); dt = torch.bfloat16
g = torch.Generator().manual_seed(0)
lib = _lib.lib()
_lib.set_option("fuse_out", 2)
shapes = [tuple(int(v) for v in a.split("x")) for a in os.environ.get("SHAPES", "16x4096x320,16x1024x640,16x256x1280").split(",")]
for (B, S, C) in shapes:
    text = torch.randn(B, 77, 768, generator=g).to(dev, dt); img = torch.randn(B, 1, 768, generator=g).to(dev, dt)
    wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt); wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    kv = ops.kv_pack(text, img, wkv, wkv, 8)
    xs = [torch.randn(B, S, C, device=dev, dtype=dt) for _ in range(2)]
    ys = [torch.empty_like(xs[0]) for _ in range(2)]; os_ = [torch.empty_like(xs[0]) for _ in range(2)]
    bo = torch.zeros(C, device=dev)
    sync = torch.zeros(int(lib.pv_dual_attn_sync_words(B, S)), device=dev, dtype=torch.int32)
    for it in range(12):
        i = (it % 2) if os.environ.get("ROT", "1") == "1" else 0
        _lib.check(lib.pv_dual_attn_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(wo), ops._ptr(bo),
                                        ops._ptr(ys[i]), None, ops._ptr(os_[i]), None, ops._ptr(sync), B, S, C, 8, 77, 1, 1.0, 1.0, ops._stream()))
        if os.environ.get("SYNC_EACH"):
            torch.cuda.synchronize()
            print(B, S, C, "call", it, "ok  counters max", int(sync.max()), flush=True)
    torch.cuda.synchronize()
    assert int(sync.abs().max()) == 0, "row-block counters must be back to zero"
    ref = ys[0].clone()
    print("shape", (B, S, C), "12 back-to-back calls ok, counters zero", flush=True)
