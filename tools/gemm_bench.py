"""Out-projection GEMM timing (CUDA graph of 20 launches, rotated buffers): single-CTA vs persistent CTA-pair kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
for M, N, K in [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (1024, 1280, 1280), (20480, 1024, 1024)]:
    nbuf = 6
    a = [torch.randn(M, K, device=dev, dtype=dt) for _ in range(nbuf)]
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.zeros(N, device=dev)
    o = [torch.empty(M, N, device=dev, dtype=dt) for _ in range(nbuf)]
    res = []
    for pair in (0, 2):
        _lib.set_option("gemm_persistent", int(pair == 2))
        for i in range(3):
            ops.linear(a[i], w, b, out=o[i])
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(20):
                ops.linear(a[i % nbuf], w, b, out=o[i % nbuf])
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        res.append(f"{("single", "pair", "persistent")[pair]}: {us:.1f} us {2 * M * N * K / us / 1e6:.0f} TFLOP/s")
    _lib.set_option("gemm_persistent", 1)
    print(f"M={M} N={N} K={K}: " + "   ".join(res))
