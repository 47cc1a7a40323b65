"""Top GPU kernels of ONE UNet evaluation of the benchmarked configuration (batch 16 rows = uncond + cond of batch 8,
latent 64, bf16, channels-last, fused epilogues on unless --stock-epilogues), by torch.profiler CUDA time.
Usage: python tools/unet_profile.py [--stock-epilogues] [--top N] [--once]      (--once: two evaluations, no profiler: the
target of an `ncu -k regex:...` capture)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

stock = "--stock-epilogues" in sys.argv
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
dev = torch.device("cuda:0")
dt = torch.bfloat16
unet, ia, ta = bench.build_models(dev, dt, channels_last=True)
if stock:
    unet.set_fused_epilogues(False)
rows = 16
x = torch.randn(rows, 4, 64, 64, device=dev, dtype=dt)
text = torch.randn(rows, 77, 768, device=dev, dtype=dt)
img = torch.randn(rows, 1, 768, device=dev, dtype=dt)
t = torch.tensor([500], device=dev)
if "--once" in sys.argv:
    with torch.no_grad():
        unet(x, t, (text, img))
        unet(x, t, (text, img))
    torch.cuda.synchronize()
    sys.exit(0)
with torch.no_grad():
    for _ in range(3):
        unet(x, t, (text, img))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        unet(x, t, (text, img))
    e1.record()
    torch.cuda.synchronize()
    print(f"eager UNet evaluation: {e0.elapsed_time(e1) / 5:.2f} ms (launch-bound in eager mode; the engine replays a CUDA graph)")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        unet(x, t, (text, img))
        torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print(f"sum of kernel times {tot / 1e3:.2f} ms over {sum(e.count for e in ev)} launches")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:top]:
    print(f"{e.device_time_total / 1e3:8.3f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count:4d}  {e.key[:120]}")
