"""Turn the raw ncu outputs in gpurun_out/ into the small, tracked summaries under profiles/.
    python tools/summarize_profiles.py <tag>      (e.g. r01a)
Reads  gpurun_out/launches_layerstack_<tag>.csv, launches_bench_<tag>.csv, prof_attn_<tag>.ncu-rep, prof_gemm_<tag>.ncu-rep
Writes profiles/<tag>_launches_layerstack.csv, <tag>_bench_kernel_shares.csv, <tag>_ncu_{attn,gemm,adapter,bwd,sattn}.csv"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def read_launch_csv(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us":
                v *= 1e3
            rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], v))
    return rows


def short(name):
    name = name.replace("void ", "")
    return name[:110]


p = os.path.join(G, f"launches_layerstack_{tag}.csv")
if os.path.exists(p):
    rows = read_launch_csv(p)
    with open(os.path.join(P, f"{tag}_launches_layerstack.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["idx", "kernel", "grid", "block", "duration_ns"])
        for i, (k, g, b, v) in enumerate(rows):
            w.writerow([i, short(k), g, b, int(v)])

p = os.path.join(G, f"launches_bench_{tag}.csv")
if os.path.exists(p):
    rows = read_launch_csv(p)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, g, b, v in rows:
        a = agg[short(k)]
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_bench_kernel_shares.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ns", "share_of_profiled_time"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            w.writerow([k, a[0], int(a[1]), round(a[1] / tot, 4)])
        ours = sum(a[1] for k, a in agg.items() if k.startswith("pv::"))
        w.writerow(["TOTAL pv:: kernels", sum(a[0] for k, a in agg.items() if k.startswith("pv::")), int(ours), round(ours / tot, 4)])
        w.writerow(["TOTAL all kernels", sum(a[0] for a in agg.values()), int(tot), 1.0])

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for what in ("attn", "gemm", "adapter", "bwd", "sattn", "backbone"):
    rep = os.path.join(G, f"prof_{what}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in ("Kernel Name", "Grid Size", "Block Size") or h in KEYS]
    with open(os.path.join(P, f"{tag}_ncu_{what}.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i][:100] for i in cols])
print("profiles/:", sorted(os.listdir(P)))

# pv:: kernels of one training step (tools/gpu_runs/r2_profiles.sh)
p = os.path.join(G, f"launches_pv_train_{tag}.csv")
if os.path.exists(p):
    rows = read_launch_csv(p)
    agg = collections.OrderedDict()
    for k, g, b, v in rows:
        a = agg.setdefault(short(k)[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_train_pv_kernels.txt"), "w") as f:
        f.write("pv:: kernels of ONE training step (config 3, batch 16, LoRA r = 8, bf16) under ncu gpu__time_duration (serialised, cold)\n")
        f.write(f"total pv ms {tot / 1e6:.2f}   launches {sum(a[0] for a in agg.values())}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"{k:72s} n={a[0]:4d} total={a[1] / 1e6:7.2f} ms avg={a[1] / a[0] / 1e3:8.1f} us\n")
