import os, sys, torch
sys.path.insert(0, "/root/repo")
from photoverse_b200 import _lib, ops
dev = torch.device("cuda:0"); dt = torch.bfloat16
for N, K in [(320, 320), (640, 640), (1280, 1280)]:
    for M in [4096, 16384, 65536, 262144]:
        nbuf = 3 if M > 100000 else 6
        a = [torch.randn(M, K, device=dev, dtype=dt) for _ in range(nbuf)]
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt); b = torch.zeros(N, device=dev)
        o = [torch.empty(M, N, device=dev, dtype=dt) for _ in range(nbuf)]
        for i in range(3): ops.linear(a[i % nbuf], w, b, out=o[i % nbuf])
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(10): ops.linear(a[i % nbuf], w, b, out=o[i % nbuf])
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        print(f"N={N} K={K} M={M}: {us:.1f} us  {2*M*N*K/us/1e6:.0f} TFLOP/s  units/pair={(M/256)*(N/160)/74:.2f}")
