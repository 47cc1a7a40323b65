"""BASELINE config 5: attention-processor microbench sweep -- latent 32^2 / 64^2 / 96^2 at the four UNet resolutions,
77 text + Li image tokens, head dims 40 / 80 / 160.  One processor call = fused attention kernel + out projection with
cached K/V (CUDA events around a CUDA graph of 10 calls).  SURVEY 8d grid: B in {1,2,8,16,64}, Li in {1,4,5,8,16}, bf16
(tensor-core path) and fp32 (the FFMA parity path), warm L2 (3 rotating buffer sets) and cold L2 (buffer sets rotated
through > 2 x the 126 MB L2).  Prints one JSON object per line and, with PV_SWEEP_OUT set, writes the list to that file.
Usage: python tools/micro_sweep.py        env: PV_ROWS_LIST, PV_LI_LIST, PV_DTYPES (bf16,f32)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
lib = _lib.lib()
PEAK = float(os.environ.get("PV_PEAK_TFLOPS", "1662"))
rows_list = [int(v) for v in os.environ.get("PV_ROWS_LIST", "1,2,8,16,64").split(",")]
li_list = [int(v) for v in os.environ.get("PV_LI_LIST", "1,4,5,8,16").split(",")]
dtypes = os.environ.get("PV_DTYPES", "bf16,f32").split(",")
out = []
for dname in dtypes:
  dt = torch.bfloat16 if dname == "bf16" else torch.float32
  code = 1 if dname == "bf16" else 0
  for L in (32, 64, 96):
    for S, C in ((L * L, 320), (L * L // 4, 640), (L * L // 16, 1280), (L * L // 64, 1280)):
        for Li in (li_list if dname == "bf16" else [5]):
            for B in (rows_list if dname == "bf16" else [2, 16]):
              for cold in (False, True):
                esz = 2 if dname == "bf16" else 4
                set_bytes = B * S * C * esz * (3 if dname == "bf16" else 4)
                nbuf = 3 if not cold else max(3, min(24, (256 << 20) // set_bytes + 1))
                if set_bytes * nbuf > 6e9 or (cold and set_bytes * 3 > (200 << 20)):      # bounded memory; big sets are cold anyway
                    continue
                text = torch.randn(B, 77, 768, generator=g).to(dev, dt)
                img = torch.randn(B, Li, 768, generator=g).to(dev, dt)
                wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
                wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
                bo = torch.zeros(C, device=dev)
                wkv = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
                kv = ops.kv_pack(text, img, wkv, wkv, 8)
                xs = [torch.randn(B, S, C, device=dev, dtype=dt) for _ in range(nbuf)]
                ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
                os_ = [torch.empty_like(xs[0]) for _ in range(nbuf)]
                qs = [torch.empty(B, S, C, device=dev) for _ in range(nbuf)] if dname == "f32" else [None] * nbuf

                sync = torch.zeros(int(lib.pv_dual_attn_sync_words(B, S)), device=dev, dtype=torch.int32)

                def run(i):
                    _lib.check(lib.pv_dual_attn_fwd(code, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp),
                                                    ops._ptr(wo), ops._ptr(bo), ops._ptr(ys[i]), ops._ptr(qs[i]), ops._ptr(os_[i]), None,
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
                rec = {"dtype": dname, "l2": "cold" if cold else "warm", "latent": L, "S": S, "C": C, "d": C // 8, "Li": Li, "B": B, "us": round(us, 2),
                       "tflops": round(flops / us / 1e6, 1), "frac_of_peak": round(flops / us / 1e6 / PEAK, 4)}
                out.append(rec)
                print(json.dumps(rec), flush=True)
                del xs, ys, os_, qs
if os.environ.get("PV_SWEEP_OUT"):
    json.dump(out, open(os.environ["PV_SWEEP_OUT"], "w"), indent=0)
