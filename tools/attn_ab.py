"""A/B of one library option on the attention kernel (and the whole processor call) at chosen attn2 layer shapes.

    python tools/attn_ab.py <option> <value> [<value> ...]      e.g.  attn6_prefetch 0 1
    PV_SHAPES="4096x320,1024x640"  PV_ROWS=16  PV_LI=1  PV_REPEAT=3

CUDA events around a CUDA graph of 20 launches, buffers rotated through more than the L2 (bench._graph_time_us); every
value is timed PV_REPEAT times in interleaved order so that clock drift does not favour one side.  The results of the
values are also compared bit for bit (an option that only moves work around must not change a single output)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from photoverse_b200 import _lib, ops  # noqa: E402

opt = sys.argv[1]
values = [int(v) for v in sys.argv[2:]]
rows = int(os.environ.get("PV_ROWS", "16"))
Li = int(os.environ.get("PV_LI", "1"))
repeat = int(os.environ.get("PV_REPEAT", "3"))
shapes = [tuple(int(t) for t in s.split("x")) for s in os.environ.get("PV_SHAPES", "4096x320,1024x640").split(",")]
dev = torch.device("cuda:0")
dt = torch.bfloat16
lib = _lib.lib()
g = torch.Generator().manual_seed(1)
for (S, C) in shapes:
    H = 8
    text = torch.randn(rows, bench.LT, 768, generator=g).to(dev, dt)
    img = torch.randn(rows, Li, 768, generator=g).to(dev, dt)
    wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev, dt)
    bo = torch.zeros(C, device=dev)
    wkv_t = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    wkv_i = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(dev, dt)
    kv = ops.kv_pack(text, img, wkv_t, wkv_i, H)
    nbuf = max(2, min(16, (192 << 20) // (3 * rows * S * C * 2) + 1))
    xs = [torch.randn(rows, S, C, device=dev, dtype=dt) for _ in range(nbuf)]
    os_ = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    sync = torch.zeros(int(lib.pv_dual_attn_sync_words(rows, S)), device=dev, dtype=torch.int32)

    def attn_only(i):
        _lib.check(lib.pv_dual_attn_core_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp),
                                             ops._ptr(os_[i]), None, rows, S, C, H, bench.LT, Li, 1.0, 1.0, ops._stream()))

    def full(i):
        _lib.check(lib.pv_dual_attn_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp), ops._ptr(wo),
                                        ops._ptr(bo), ops._ptr(ys[i]), None, ops._ptr(os_[i]), None, ops._ptr(sync),
                                        rows, S, C, H, bench.LT, Li, 1.0, 1.0, ops._stream()))

    t_attn = {v: [] for v in values}
    t_full = {v: [] for v in values}
    outs = {}
    for _ in range(repeat):
        for v in values:
            _lib.set_option(opt, v)
            t_attn[v].append(bench._graph_time_us(attn_only, nbuf))
            t_full[v].append(bench._graph_time_us(full, nbuf))
            attn_only(0)
            full(1)
            torch.cuda.synchronize()
            outs[v] = (os_[0].clone(), ys[1].clone())
    for v in values:
        same = all(torch.equal(outs[v][k], outs[values[0]][k]) for k in (0, 1))
        print(f"S{S} C{C} rows {rows} Li {Li}  {opt}={v}: attn {min(t_attn[v]):.2f} us (runs {[round(t, 2) for t in t_attn[v]]})  "
              f"processor {min(t_full[v]):.2f} us (runs {[round(t, 2) for t in t_full[v]]})  identical_to_first={same}")
    del xs, os_, ys
