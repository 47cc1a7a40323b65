#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config 2 of BASELINE.json): 50-step DDIM generation at 512^2 (latent 64^2), batch 8 per GPU, bf16,
guidance 1.0, SD-1.5-shaped UNet + PhotoVerse adapters, random-init weights, synthetic encoder outputs.
One bench "step" = ONE full generation of the per-GPU batch: image/text adapters once, K/V projection of the 16
attn2 layers once, then 50 denoising steps, each evaluating the UNet on the unconditional AND the conditional
branch as the reference does (models/infer.py:103-114) -- here as one doubled-batch call.

Reported on one JSON line (rank 0):
  value      images/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST (pinned) inputs: H2D of every input + D2H of the latents
             inside the timed region
  roofline   the fused Q-projection + dual-branch attention kernel (the dominant native kernel), timed alone with
             CUDA events over the 16 attn2 layer shapes of one UNet evaluation, against the measured bf16 peak
  cpu_baseline  the oracle port of the reference path, host UNet in fp32 on the host cores, bounded sample
  --impl reference : the CPU arm as the whole job (same metric / config keys), rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line: NCCL's own messages (e.g. the "NCCL version ..." banner that NCCL_DEBUG=VERSION
# prints on the first communicator) go to stderr.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @50 steps 512^2; dual-branch cross-attn TFLOP/s as % of B200 peak"
LAYER_SHAPES = [(4096, 320)] * 5 + [(1024, 640)] * 5 + [(256, 1280)] * 5 + [(64, 1280)]   # (S, C) of the 16 attn2 layers
LT = 77


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per generation")
    ap.add_argument("--denoise-steps", type=int, default=50)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--token-index", default="0", help="adapter head used at inference (reference default 0) or 'full'")
    ap.add_argument("--mode", default="batched", choices=["batched", "two_call", "cond_only"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="keep the host UNet in NCHW (default: channels_last)")
    ap.add_argument("--workload", default="generate", choices=["generate", "train"],
                    help="generate: BASELINE config[1] (the headline line);  train: config[3] training step")
    ap.add_argument("--train-batch", type=int, default=16, help="samples per GPU per training step (config[3])")
    ap.add_argument("--lora-rank", type=int, default=8)
    ap.add_argument("--lora-dropout", type=float, default=0.0,
                    help="train workload: peft lora_dropout (reference default 0.1; 0 = SURVEY config 4 parity setting)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": float(d["bf16_tflops"]), "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [x for x in sm if x >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# model construction (identical seeds on both arms)
# ------------------------------------------------------------------------------------------------------------
def build_models(device, dtype, T=5, channels_last=True):
    import photoverse_b200 as pv
    torch.backends.cudnn.benchmark = True
    from photoverse_b200.host.unet_sd15 import UNetSD15
    torch.manual_seed(0)
    unet = UNetSD15()
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(T,))
    image_adapter = pv.PhotoVerseAdapter(num_tokens=T)
    text_adapter = pv.PhotoVerseAdapter(num_tokens=T)
    for m in (unet, image_adapter, text_adapter):
        m.requires_grad_(False)
        m.eval()
        m.to(device=device, dtype=dtype)
    if channels_last:      # host-model plumbing: NHWC activations for the cuDNN convolutions of the backbone
        unet.to(memory_format=torch.channels_last)
    return unet, image_adapter, text_adapter


def pinned_inputs(batch, latent, seed, dtype):
    from photoverse_b200.host.pipeline import GenInputs, synthetic_inputs
    h = synthetic_inputs(batch, latent, seed=seed, dtype=dtype)
    pin = lambda t: t.contiguous().pin_memory()
    return GenInputs([pin(t) for t in h.clip_hidden], [pin(t) for t in h.clip_hidden_uncond], pin(h.text),
                     pin(h.text_uncond), pin(h.noise))


# ------------------------------------------------------------------------------------------------------------
# roofline leg: the fused attention kernel alone, per attn2 layer shape, CUDA-event timed
# ------------------------------------------------------------------------------------------------------------
def attn_kernel_flops(rows_b, S, C, Li):
    return 2 * rows_b * S * C * C + 4 * rows_b * S * C * (LT + Li)      # Q projection + QK^T + PV (SURVEY 8d terms)


def processor_flops_cached(rows_b, S, C, Li):
    return 4 * rows_b * S * C * C + 4 * rows_b * S * C * (LT + Li)      # + out projection; K/V cached across steps


def roofline_leg(device, rows_b, Li, scale):
    """Returns roofline dict for the fused attention kernel + per-family detail + whole-processor numbers."""
    from photoverse_b200 import _lib, ops
    peaks = measured_peaks()
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(1)
    fam = {}
    l2_bytes = 192 << 20
    for (S, C) in sorted(set(LAYER_SHAPES), reverse=True):
        H = 8
        text = torch.randn(rows_b, LT, 768, generator=g).to(device, dt)
        img = torch.randn(rows_b, Li, 768, generator=g).to(device, dt)
        wq = (torch.randn(C, C, generator=g) / C ** 0.5).to(device, dt)
        wo = (torch.randn(C, C, generator=g) / C ** 0.5).to(device, dt)
        bo = torch.zeros(C, device=device)
        wkv_t = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(device, dt)
        wkv_i = (torch.randn(2 * C, 768, generator=g) / 768 ** 0.5).to(device, dt)
        kv = ops.kv_pack(text, img, wkv_t, wkv_i, H)
        xbytes = rows_b * S * C * 2
        nbuf = max(2, min(16, l2_bytes // max(1, 3 * xbytes) + 1))     # rotate X/O/Y sets larger than L2 in total
        xs = [torch.randn(rows_b, S, C, device=device, dtype=dt) for _ in range(nbuf)]
        os_ = [torch.empty_like(xs[0]) for _ in range(nbuf)]
        ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
        lib = _lib.lib()
        sync = torch.zeros(int(lib.pv_dual_attn_sync_words(rows_b, S)), device=device, dtype=torch.int32)

        def attn_only(i):
            _lib.check(lib.pv_dual_attn_core_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp),
                                                 ops._ptr(os_[i]), None, rows_b, S, C, H, LT, Li, 1.0, 1.0, ops._stream()))

        def full(i):
            _lib.check(lib.pv_dual_attn_fwd(1, ops._ptr(xs[i]), ops._ptr(wq), ops._ptr(kv.Kp), ops._ptr(kv.Vp),
                                            ops._ptr(wo), ops._ptr(bo), ops._ptr(ys[i]), None, ops._ptr(os_[i]), None,
                                            ops._ptr(sync), rows_b, S, C, H, LT, Li, 1.0, 1.0, ops._stream()))

        out = {}
        for name, fn in (("attn", attn_only), ("proc", full)):
            for i in range(3):
                fn(i % nbuf)
            reps = 20
            torch.cuda.synchronize()
            # device time: the launches are captured in a CUDA graph (as the generation engine runs them) so that the
            # Python / ctypes launch rate (~15 us per call) does not bound kernels that are shorter than that
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for i in range(reps):
                    fn(i % nbuf)
            gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            out[name + "_us"] = e0.elapsed_time(e1) * 1e3 / reps
        fam[f"S{S}_C{C}"] = out
        del xs, os_, ys
    # aggregate over the 16 layers of one UNet evaluation
    t_attn = sum(fam[f"S{S}_C{C}"]["attn_us"] for S, C in LAYER_SHAPES) * 1e-6
    t_proc = sum(fam[f"S{S}_C{C}"]["proc_us"] for S, C in LAYER_SHAPES) * 1e-6
    f_attn = sum(attn_kernel_flops(rows_b, S, C, Li) for S, C in LAYER_SHAPES)
    f_proc = sum(processor_flops_cached(rows_b, S, C, Li) for S, C in LAYER_SHAPES)
    ach = f_attn / t_attn / 1e12
    roof = {"bound": "tensor", "kernel": "dual_attn_fwd_pair_roles_kernel (head_dim 40 / 80 layers) / dual_attn_fwd_pair_kernel (head_dim 160): fused Q-proj + dual-branch attention on cta_group::2 CTA pairs; single-CTA persistent kernel for S <= 128",
            "achieved": round(ach, 2), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": round(ach / peaks["bf16_tflops"], 4), "peak_source": peaks["source"] + " cuBLAS bf16 burst",
            # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant (S4096, C320) instance from the ncu
            # --set full capture of tools/profile_layer_stack.py (profiles/r01n_ncu_attn.csv): 44.7 MB + 4.6 MB;
            # algorithmic bytes of that launch: X 41.9 MB in (+ 0.2 MB Wq, 2.4 MB K/V tiles), O 41.9 MB out (stays in L2)
            "traffic": 49.3e6, "traffic_unit": "bytes per launch, S4096_C320 layer (ncu r01n)",
            "how": f"CUDA events around a CUDA graph of 20 launches per attn2 layer shape at {rows_b} rows (uncond+cond), "
                   f"inputs rotated through > L2 of buffers; aggregated over the 16 layers of one UNet evaluation; "
                   f"algorithmic FLOPs = 2 rows C^2 + 4 rows C (77 + Li) per launch",
            "per_shape_us": {k: {kk: round(vv, 2) for kk, vv in v.items()} for k, v in fam.items()},
            "processor_tflops": round(f_proc / t_proc / 1e12, 2),
            "processor_frac": round(f_proc / t_proc / 1e12 / peaks["bf16_tflops"], 4),
            "processor_ms_per_unet_eval": round(t_proc * 1e3, 3)}
    return roof


# ------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, denoise_steps, latent, token_index, T=5):
    """Each step = bounded sample: 1 image, adapters once + ONE denoise step (uncond + cond UNet evaluation, fp32)."""
    from oracle.host_reference import clone_adapter_as_oracle, clone_with_oracle_processors
    from photoverse_b200.host.ddim import make_ddim_schedule
    from photoverse_b200.host.pipeline import synthetic_inputs
    import photoverse_b200 as pv
    from photoverse_b200.host.unet_sd15 import UNetSD15
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    unet = UNetSD15()
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(T,))
    unet = clone_with_oracle_processors(unet).eval()
    torch.manual_seed(1)
    image_adapter = clone_adapter_as_oracle(pv.PhotoVerseAdapter(num_tokens=T)).eval()
    text_adapter = clone_adapter_as_oracle(pv.PhotoVerseAdapter(num_tokens=T)).eval()
    inp = synthetic_inputs(1, latent, seed=0, dtype=torch.float32)
    sched = make_ddim_schedule(denoise_steps)
    t_ad, t_step = [], []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            text_adapter(inp.clip_hidden, token_index=token_index)
            img = image_adapter(inp.clip_hidden, token_index=token_index)
            img_u = image_adapter(inp.clip_hidden_uncond, token_index=token_index)
            t1 = time.perf_counter()
            lat = inp.noise
            t = torch.tensor([float(sched.timesteps[0])])
            eu = unet(lat, t, encoder_hidden_states=(inp.text_uncond, img_u)).sample     # infer.py:103-107
            ec = unet(lat, t, encoder_hidden_states=(inp.text, img)).sample              # infer.py:110-114
            eps = eu + 1.0 * (ec - eu)
            lat = sched.c_x[0] * lat + sched.c_eps[0] * eps
            t2 = time.perf_counter()
            if it >= warmup:
                t_ad.append(t1 - t0)
                t_step.append(t2 - t1)
    ad, st = statistics.mean(t_ad), statistics.mean(t_step)
    per_image_s = ad + denoise_steps * st
    return {"images_per_s": 1.0 / per_image_s, "cores": cores, "adapter_s": ad, "denoise_step_s": st,
            "sample": f"1 image: adapters once + 1 of {denoise_steps} DDIM steps (uncond+cond UNet evaluation), fp32, "
                      f"extrapolated as adapters + {denoise_steps} x step; mean of {steps} after {warmup} warm-up"}


def run_train(args, rank, world, local_rank):
    """BASELINE config[3]: one training step = adapters + UNet forward in grad mode (16 processors, stochastic fusion),
    backward into the trainable set {adapters, to_k_ip/to_v_ip, LoRA A/B}, ONE flat-buffer NCCL allreduce, per-group
    clipping, AdamW.  bf16 backbone, fp32 masters for the trainable set.  Weak scaling: batch per GPU fixed."""
    import torch.distributed as dist
    import photoverse_b200 as pv
    from photoverse_b200 import _lib
    from photoverse_b200.host.train_step import Trainer, synthetic_train_batch
    from photoverse_b200.host.unet_sd15 import UNetSD15
    from photoverse_b200.lora import inject_lora
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    unet = UNetSD15()
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(5,))
    ia, ta = pv.PhotoVerseAdapter(num_tokens=5), pv.PhotoVerseAdapter(num_tokens=5)
    unet.requires_grad_(False)
    inject_lora(unet, r=args.lora_rank, lora_dropout=args.lora_dropout)
    unet.to(device=device, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    for m in (ia, ta):
        m.to(device)
    for n, p in unet.named_parameters():
        if "to_k_ip" in n or "to_v_ip" in n or "lora_" in n:
            p.data = p.data.float()                       # fp32 masters for the trainable set
            p.requires_grad_(True)
    unet.eval()
    if args.lora_dropout > 0:
        pv.unet.set_cross_attention_layers_to_train(unet)     # train.py:462 (activates the LoRA dropout)
    tr = Trainer(unet, ia, ta)
    b = synthetic_train_batch(args.train_batch, args.latent, seed=100 + rank, device=device, dtype=torch.bfloat16)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(1000 + rank)
    for _ in range(args.warmup):
        tr.step(b)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = tr.step(b)
    e1.record()
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    if rank == 0:
        line = {"metric": "training samples/s, adapter + LoRA + to_k_ip/to_v_ip step through dual-branch attention (config[3])",
                "value": round(args.train_batch * world * args.steps / (ms * 1e-3), 3), "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"config[3]: training step, batch {args.train_batch}/GPU, latent {args.latent}^2, Li=5, LoRA r="
                                       f"{args.lora_rank} (dropout {args.lora_dropout}) on attn2.to_q/k/v, fwd+bwd through 16 processors + 2 adapters, flat-buffer "
                                       f"NCCL allreduce of {tr.buf.numel()} fp32 gradients, per-group clip, AdamW",
                           "batch_per_gpu": args.train_batch, "grad_elements": tr.buf.numel()},
                "gpu_launches": int(_lib.launch_count() - n0), "clocks": clocks, "final_loss": round(float(loss), 5)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    token_index = args.token_index if args.token_index == "full" else int(args.token_index)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    Li = 5 if token_index == "full" else 1
    config = {"workload": f"config[1]: {args.denoise_steps}-step DDIM generation, batch {args.batch}/GPU, latent "
                          f"{args.latent}^2 (512^2 px), guidance 1.0, SD-1.5-shaped UNet (random init) + PhotoVerse "
                          f"processors on 16 attn2 layers + image/text adapters",
              "batch_per_gpu": args.batch, "denoise_steps": args.denoise_steps, "latent": args.latent,
              "unet_evals_per_step": "uncond+cond (reference infer.py:103-114)", "image_tokens": Li,
              "token_index": token_index, "l2": "step working set >> L2 (126 MB); roofline leg rotates buffers > L2"}

    if args.workload == "train" and args.impl == "ours":
        return run_train(args, rank, world, local_rank)
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(args.steps, args.warmup, args.denoise_steps, args.latent, token_index)
        line = {"impl": "reference", "metric": METRIC, "value": round(r["images_per_s"], 6), "unit": "images/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round((r["adapter_s"] + r["denoise_step_s"]) * 1e3, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": round(r["images_per_s"], 6), "unit": "images/s", "cores": r["cores"],
                                 "kind": "port", "sample": r["sample"]},
                "e2e": {"value": round(r["images_per_s"], 6), "unit": "images/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch.distributed as dist
    from photoverse_b200 import _lib
    from photoverse_b200.host.pipeline import GenerationEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: photoverse_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- CPU baseline (rank 0, N == 1 only) before the GPU section ----
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(2, 1, args.denoise_steps, args.latent, token_index)
        cpu_base = {"value": round(r["images_per_s"], 6), "unit": "images/s", "cores": r["cores"], "kind": "port",
                    "sample": r["sample"], "denoise_step_s": round(r["denoise_step_s"], 3)}

    dtype = torch.bfloat16
    unet, image_adapter, text_adapter = build_models(device, dtype, channels_last=not args.nchw)
    eng = GenerationEngine(unet, image_adapter, text_adapter, args.batch, args.latent, args.denoise_steps, 1.0,
                           token_index, args.mode, dtype, device, use_cuda_graph=not args.no_graph)
    host = pinned_inputs(args.batch, args.latent, seed=100 + rank, dtype=dtype)
    h2d = host.nbytes()
    out_host = torch.empty(args.batch, 4, args.latent, args.latent, dtype=dtype).pin_memory()
    d2h = out_host.numel() * out_host.element_size()

    # ---- device-resident leg ("value") ----
    eng.load_inputs(host)
    for _ in range(args.warmup):
        lat = eng.generate()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0, r0 = _lib.launch_count(), eng.replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lat = eng.generate()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = (_lib.launch_count() - n0) + (eng.replays - r0) * eng.launches_per_eval
    finite = bool(torch.isfinite(lat.float()).all().item())

    # ---- end-to-end leg: host buffers in, host latents out, copies inside the timed region ----
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        eng.load_inputs(host)
        lat = eng.generate()
        out_host.copy_(lat, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    images = args.batch * world * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)

    roof = None
    if rank == 0:
        rows_b = 2 * args.batch if args.mode == "batched" else args.batch
        roof = roofline_leg(device, rows_b, Li, 1.0)
        # share of the generation spent in the native processor kernels (Amdahl note)
        per_gen_proc_ms = roof["processor_ms_per_unet_eval"] * args.denoise_steps * (2 if args.mode == "two_call" else 1)
        roof["processor_share_of_step"] = round(per_gen_proc_ms / (ms / args.steps), 4)
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 4), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "roofline": roof, "cpu_baseline": cpu_base,
                "e2e": {"value": round(e2e_value, 4), "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / args.steps, 3)},
                "gpu_launches": int(launches), "clocks": clocks, "finite_output": finite,
                "native_kernels_per_unet_eval": eng.launches_per_eval}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
