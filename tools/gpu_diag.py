"""Diagnostic run for the GPU box: exercises each native kernel family separately and prints error magnitudes
(never asserts), so a single gpurun call localises a failing primitive.  Output: gpurun_out/diag.json"""
import json
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
res = {}
SECTION = sys.argv[1] if len(sys.argv) > 1 else "all"


def rec(name, fn):
    try:
        v = fn()
        torch.cuda.synchronize()
        res[name] = v
    except Exception as e:  # noqa: BLE001
        res[name] = "EXC " + repr(e)[:300]
        traceback.print_exc()
    print(name, res[name], flush=True)


def gemm(M, N, K, bn, swz, out_f32):
    _lib.set_option("epi_swizzle", swz)
    _lib.set_option("force_bn", bn)
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev, torch.bfloat16)
    b = torch.randn(N, generator=g).to(dev)
    y = ops.linear(a, w, b, out_dtype=torch.float32 if out_f32 else torch.bfloat16)
    ref = a.float() @ w.float().t() + b
    d = (y.float() - ref).abs()
    return {"max": d.max().item(), "mean": d.mean().item(), "ref_absmax": ref.abs().max().item()}


for bn in ((64, 128, 160, 256) if SECTION in ("all", "gemm") else ()):
    for swz in (0, 1):
        for of in (True, False):
            rec(f"gemm_bn{bn}_swz{swz}_f32out{int(of)}", lambda: gemm(256, 640, 320, bn, swz, of))
_lib.set_option("epi_swizzle", 1)
_lib.set_option("force_bn", 0)
if SECTION in ("all", "gemm"):
    rec("gemm_ragged", lambda: gemm(300, 328, 776, 0, 1, False))


def attn_core(S, C, Li, B, dtype):
    from oracle import cases
    from oracle.processor_oracle import dual_branch_attention
    from tests.helpers import build_product_layer
    case = cases.ProcCase("d", B=B, S=S, C=C, Li=Li, seed=5)
    attn, proc = build_product_layer(case, dev)
    x, text, img = cases.proc_inputs(case)
    with torch.no_grad():
        y = attn(x.to(dev, dtype), encoder_hidden_states=(text.to(dev, dtype), img.to(dev, dtype)))
        w = cases.proc_weights(case).to(device=dev)
        yr, vr = dual_branch_attention(x.to(dev), text.to(dev), img.to(dev), w)
    d = (y.float() - yr).abs()
    return {"max": d.max().item(), "mean": d.mean().item(), "nan": bool(torch.isnan(y.float()).any().item()),
            "vnorm": (proc.to_v_ip_norm.float() - vr).abs().max().item()}


for dtype in [d for d in (torch.float32, torch.bfloat16) if SECTION in ("all", "attn_" + str(d)[6:])]:
    for (S, C) in ((256, 320), (128, 640), (128, 1280), (64, 1280), (200, 320)):
        rec(f"proc_{str(dtype)[6:]}_S{S}_C{C}", lambda: attn_core(S, C, 5, 2, dtype))

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/diag_{SECTION}.json", "w"), indent=1)
