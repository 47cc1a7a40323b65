"""Attention backward core (pv_dual_attn_bwd + pv_kv_pack_bwd) per attn2 layer shape at the config[3] batch (16 per GPU),
bf16, CUDA events around a CUDA graph of 10 calls.  Prints time and the forward kernel's time for comparison."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photoverse_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
B, LI = int(os.environ.get("PV_ROWS", "16")), int(os.environ.get("PV_LI", "5"))
g = torch.Generator().manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for S, C in [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]:
    H = 8
    text = torch.randn(B, 77, 768, generator=g).to(dev, dt)
    img = torch.randn(B, LI, 768, generator=g).to(dev, dt)
    wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    wkv_t = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    wkv_i = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    kv = ops.kv_pack(text, img, wkv_t, wkv_i, H)
    x = torch.randn(B, S, C, device=dev, dtype=dt)
    o, stats = ops.dual_attn_core(x, wq, kv, 1.0, 1.0, want_stats=True)
    q = ops.linear(x.view(B * S, C), wq).view(B, S, C)
    d_o = torch.randn(B, S, C, device=dev, dtype=dt)
    dvn = torch.randn(B, H, LI, device=dev)
    t_fwd = timed(lambda: ops.dual_attn_core(x, wq, kv, 1.0, 1.0, want_stats=True))
    t_bwd = timed(lambda: ops.dual_attn_bwd(d_o, q, kv.kv_text, kv.kv_img, stats, kv.v_ip_norm, dvn, H, 77, LI, 1.0, 1.0))
    fl = 10 * B * H * S * 96 * (C // H)
    print(f"S={S} C={C}: attention backward (core + K/V reduce) {t_bwd:8.1f} us  ({fl / t_bwd / 1e6:6.1f} TFLOP/s padded)   "
          f"forward kernel (Q-proj + attention + stats) {t_fwd:7.1f} us   ratio {t_bwd / t_fwd:4.1f}x")

print("LoRA factor gradients (dA, dB) of one projection, rank 8: fused one-pass kernel vs the GEMM route")
from photoverse_b200 import autograd as ag  # noqa: E402
for M, in_f, out_f in [(B * 4096, 320, 320), (B * 1024, 640, 640), (B * 256, 1280, 1280), (B * 77, 768, 320), (B * 77, 768, 1280)]:
    r, s_ = 8, 0.125
    x = torch.randn(M, in_f, device=dev, dtype=dt)
    gy = torch.randn(M, out_f, device=dev, dtype=dt)
    A = torch.randn(r, in_f, device=dev) / in_f ** 0.5
    Bm = torch.randn(out_f, r, device=dev) * 0.1

    def gemm_route():
        a_c = A.to(dt).contiguous()
        bt_c = ops.transpose(Bm.to(dt).contiguous())
        t = ag._skinny_linear(x, a_c)
        db = ops.linear_bwd_weight(gy, t, alpha=s_)
        u = ag._skinny_linear(gy, bt_c)
        da = ops.linear_bwd_weight(u, x, alpha=s_)
        return da, db

    t_f = timed(lambda: ops.lora_bwd(x, gy, A, Bm, s_))
    t_g = timed(gemm_route)
    print(f"M={M} in={in_f} out={out_f}: fused {t_f:7.1f} us   GEMM route {t_g:7.1f} us   bytes/t = {(M * (in_f + out_f) * 2) / t_f / 1e3:6.0f} GB/s")
