"""BASELINE config 5: attention-processor microbench sweep -- latent 32^2 / 64^2 / 96^2 at the four UNet resolutions,
77 text + Li image tokens, head dims 40 / 80 / 160.  One processor call = fused attention kernel + out projection with
cached K/V (CUDA events around a CUDA graph of 10 calls, buffers rotated).  Prints one JSON object per line and, with
PV_SWEEP_OUT set, writes the list to that file.  Usage: python tools/micro_sweep.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
g = torch.Generator().manual_seed(0)
lib = _lib.lib()
PEAK = float(os.environ.get("PV_PEAK_TFLOPS", "1662"))
rows_list = [int(v) for v in os.environ.get("PV_ROWS_LIST", "2,16").split(",")]
out = []
for L in (32, 64, 96):
    for S, C in ((L * L, 320), (L * L // 4, 640), (L * L // 16, 1280), (L * L // 64, 1280)):
        for Li in (1, 5, 16):
            for B in rows_list:
                if B * S * C * 2 * 6 > 3e9:          # keep the rotated buffers bounded
                    continue
                text = torch.randn(B, 77, 768, generator=g).to(dev, dt)
                img = torch.randn(B, Li, 768, generator=g).to(dev, dt)
                wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
                wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
                bo = torch.zeros(C, device=dev)
                wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
                kv = ops.kv_pack(text, img, wkv, wkv, 8)
                nbuf = 3
                xs = [torch.randn(B, S, C, device=dev, dtype=dt) for _ in range(nbuf)]
                ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
                os_ = [torch.empty_like(xs[0]) for _ in range(nbuf)]

                sync = torch.zeros(int(lib.pv_dual_attn_sync_words(B, S)), device=dev, dtype=torch.int32)

                def run(i):
                    _lib.check(lib.pv_dual_attn_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp),
                                                    ops._ptr(wo), ops._ptr(bo), ops._ptr(ys[i]), None, ops._ptr(os_[i]), None,
                                                    ops._ptr(sync), B, S, C, 8, 77, Li, 1.0, 1.0, ops._stream()))
                for i in range(3):
                    run(i % nbuf)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for i in range(10):
                        run(i % nbuf)
                gr.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / 10
                flops = 4 * B * S * C * C + 4 * B * S * C * (77 + Li)        # cached K/V: projections of X and O + both branches
                rec = {"latent": L, "S": S, "C": C, "d": C // 8, "Li": Li, "B": B, "us": round(us, 2),
                       "tflops": round(flops / us / 1e6, 1), "frac_of_peak": round(flops / us / 1e6 / PEAK, 4)}
                out.append(rec)
                print(json.dumps(rec), flush=True)
                del xs, ys, os_
if os.environ.get("PV_SWEEP_OUT"):
    json.dump(out, open(os.environ["PV_SWEEP_OUT"], "w"), indent=0)
