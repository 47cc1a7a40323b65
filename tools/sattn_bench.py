"""attn1 self-attention kernel (pv_self_attn_fwd) against torch SDPA (cuDNN / flash) on the same inputs: max-abs error vs an
fp32 reference and CUDA-event time of graph-replayed launches, at the four SD-1.5 self-attention shapes (16 rows = the
benchmark's doubled batch).  Usage: python tools/sattn_bench.py   env: PV_ROWS (default 16), PV_SHAPES="S:C,S:C"."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("PV_ROWS", "16"))
shapes = [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]
if os.environ.get("PV_SHAPES"):
    shapes = [tuple(int(x) for x in s.split(":")) for s in os.environ["PV_SHAPES"].split(",")]
H = 8
if os.environ.get("PV_POLY"):
    _lib.set_option("sattn_poly", int(os.environ["PV_POLY"]))
g = torch.Generator().manual_seed(0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for S, C in shapes:
    d = C // H
    qkv = torch.randn(B, S, 3 * C, generator=g).to(dev, torch.bfloat16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]

    def heads(t):
        return t.reshape(B, S, H, d).transpose(1, 2)

    def sdpa():
        return F.scaled_dot_product_attention(heads(q), heads(k), heads(v)).transpose(1, 2).reshape(B, S, C)

    ours = ops.self_attn(q, k, v, H)
    torch.cuda.synchronize()
    ref = F.scaled_dot_product_attention(heads(q).float(), heads(k).float(), heads(v).float()).transpose(1, 2).reshape(B, S, C)
    err = (ours.float() - ref).abs().max().item()
    err_sdpa = (sdpa().float() - ref).abs().max().item()
    t_ours = timed(lambda: ops.self_attn(q, k, v, H))
    t_sdpa = timed(sdpa)
    flops = 4.0 * B * H * S * S * d
    print(f"S{S} C{C} d{d}: ours {t_ours:8.1f} us ({flops / t_ours / 1e6:6.1f} TF)  sdpa {t_sdpa:8.1f} us ({flops / t_sdpa / 1e6:6.1f} TF)  "
          f"speedup {t_sdpa / t_ours:5.2f}x   max-abs err ours {err:.2e}  sdpa {err_sdpa:.2e}  finite {bool(torch.isfinite(ours.float()).all())}",
          flush=True)
