"""Roofline evidence for the adapter kernels (reference models/adapters.py:14-28: Linear -> LayerNorm -> LeakyReLU x2,
Linear; :36,41 patch mean): achieved HBM GB/s of the LN + LeakyReLU epilogue kernel and of the patch-mean kernel against
the measured HBM peak, achieved TFLOP/s of the 1024 x 1024 adapter GEMMs against the measured bf16 peak, at the two
shapes of the path: generation (one token head, batch 8 -> 2 048 patch rows) and training (5 heads, batch 16 -> 5 x 4 096).
CUDA events around a CUDA graph of 20 launches, buffers rotated through > L2.  `--once`: a single pass (ncu target).
Writes one JSON object (stdout)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from photoverse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
once = "--once" in sys.argv
peaks = bench.measured_peaks()
H = 1024
out = {"peaks": peaks, "shapes": {}}
for name, T, B in (("generation: 1 head, batch 8", 1, 8), ("training: 5 heads, batch 16", 5, 16)):
    M = B * 256
    nbuf = max(2, min(12, (256 << 20) // (T * M * H * 6) + 1))
    hs = [torch.randn(T, M, H, device=dev) for _ in range(nbuf)]                    # fp32 pre-norm activations
    as_ = [torch.empty(T, M, H, device=dev, dtype=dt) for _ in range(nbuf)]         # bf16 post-activation
    gam, bet = torch.ones(T, H, device=dev), torch.zeros(T, H, device=dev)
    w = (torch.randn(T, H, H, device=dev) / 32).to(dt)
    bias = torch.zeros(T, H, device=dev)
    pm = [torch.empty(T * B, H, device=dev, dtype=dt) for _ in range(nbuf)]

    def ln(i):
        ops.ln_lrelu(hs[i].view(T * M, H), gam, bet, as_[i].view(T * M, H), rows_per_group=M)

    def gm(i):
        ops.group_mean(as_[i].view(T * B, 256, H), pm[i])

    def gemm(i):
        ops.linear(as_[i], w, bias, out=hs[i])

    rec = {}
    if once:
        for fn in (ln, gm, gemm):
            fn(0)
        torch.cuda.synchronize()
        continue
    t = bench._graph_time_us(ln, nbuf)
    by = T * M * H * (4 + 2)
    rec["ln_lrelu_kernel"] = {"us": round(t, 2), "algorithmic_bytes": by, "GBps": round(by / t / 1e3, 1),
                              "frac_of_hbm_peak": round(by / t / 1e3 / peaks["hbm_gbs"], 3), "bytes_per_element": "4 in (fp32) + 2 out (bf16)"}
    t = bench._graph_time_us(gm, nbuf)
    by = T * M * H * 2 + T * B * H * 2
    rec["group_mean_kernel"] = {"us": round(t, 2), "algorithmic_bytes": by, "GBps": round(by / t / 1e3, 1),
                                "frac_of_hbm_peak": round(by / t / 1e3 / peaks["hbm_gbs"], 3)}
    t = bench._graph_time_us(gemm, nbuf)
    fl = 2 * T * M * H * H
    rec["adapter_gemm_1024x1024 (gemm_bf16_tcgen05_kernel, fp32 out)"] = {
        "us": round(t, 2), "flops": fl, "TFLOPs": round(fl / t / 1e6, 1), "frac_of_bf16_peak": round(fl / t / 1e6 / peaks["bf16_tflops"], 3),
        "hbm_bytes": T * M * H * (2 + 4) + T * H * H * 2, "note": "A bf16 in, fp32 pre-norm out: 6 B per output element -> HBM roofline "
        f"{round(T * M * H * 6 / peaks['hbm_gbs'] / 1e3, 1)} us"}
    out["shapes"][name] = rec
if not once:
    print(json.dumps(out, indent=1))
