"""Generation loop around the hot path (counterpart of the reference's ``models/infer.py:run_inference``).

Scope note: this is the *caller* of the B200 path, written in plain PyTorch.  CLIP encoders, the
tokenizer and the VAE are outside the path -- their outputs (CLIP ViT-L/14 hidden states, text-encoder
states, initial noise) are the inputs here, and the final latents are the output.

What it keeps from the reference (infer.py):
  :89-91   text_adapter / image_adapter (cond and zero-image "uncond") run once per generation, token_index=0 default
  :98-119  per step: eps_uncond and eps_cond from the UNet with (text, image) conditioning tuples, classifier-free
           guidance combine, scheduler step
What it does the B200 way:
  * the two UNet calls of a step are ONE doubled-batch call (mode "batched"); "two_call" reproduces the reference's
    call pattern exactly; "cond_only" skips the uncond branch, valid only for guidance_scale == 1
  * K/V projections of all 16 attn2 layers are computed once per generation and cached
  * one UNet evaluation is captured in a CUDA graph and replayed for the remaining steps (no host sync in the loop)
  * DDIM (BASELINE.json) instead of DPM-Solver++ -- same scheduler on both arms of every comparison
"""
from dataclasses import dataclass
from typing import List, Optional

import torch

from ..unet import set_kv_cache
from .ddim import make_ddim_schedule


@dataclass
class GenInputs:
    clip_hidden: List[torch.Tensor]           # T x [B,257,1024]: [last_hidden_state, hs[4], hs[8], hs[12], hs[16]]
    clip_hidden_uncond: List[torch.Tensor]    # same for the all-zero image (infer.py:77-78)
    text: torch.Tensor                        # [B,77,768] text-encoder states (concept tokens injected upstream)
    text_uncond: torch.Tensor                 # [B,77,768] for the empty prompt
    noise: torch.Tensor                       # [B,4,h,w]

    def to(self, device=None, dtype=None, non_blocking=False):
        def cv(t):
            return t.to(device=device, dtype=dtype, non_blocking=non_blocking)
        return GenInputs([cv(t) for t in self.clip_hidden], [cv(t) for t in self.clip_hidden_uncond],
                         cv(self.text), cv(self.text_uncond), cv(self.noise))

    def nbytes(self) -> int:
        ts = self.clip_hidden + self.clip_hidden_uncond + [self.text, self.text_uncond, self.noise]
        return sum(t.numel() * t.element_size() for t in ts)


class _GraphedUNet:
    """One UNet evaluation captured in a CUDA graph (static input buffers, replayed every step)."""

    def __init__(self, unet, latent_shape, ctx, dtype, device):
        self.unet = unet
        self.ctx = ctx
        self.x = torch.zeros(latent_shape, device=device, dtype=dtype)
        self.t = torch.zeros(1, device=device, dtype=torch.float32)
        s = torch.cuda.Stream(device=device)
        s.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(s):
            for _ in range(2):       # warm-up outside capture: packs weights, fills the K/V cache, cuDNN autotune
                unet(self.x, self.t, encoder_hidden_states=ctx)
        torch.cuda.current_stream(device).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = unet(self.x, self.t, encoder_hidden_states=ctx).sample

    def __call__(self, x, t_dev):
        self.x.copy_(x)
        self.t.copy_(t_dev)
        self.graph.replay()
        return self.out


@torch.no_grad()
def run_generation(unet, image_adapter, text_adapter, inputs: GenInputs, num_steps: int = 50,
                   guidance_scale: float = 1.0, token_index=0, mode: str = "batched", use_cuda_graph: bool = True,
                   kv_cache: bool = True, return_aux: bool = False):
    """Returns the final latents [B,4,h,w] (and the adapter outputs if ``return_aux``)."""
    assert mode in ("batched", "two_call", "cond_only")
    if mode == "cond_only" and guidance_scale != 1.0:
        raise ValueError("mode='cond_only' drops the unconditional branch: only valid for guidance_scale == 1")
    dev, dtype = inputs.noise.device, inputs.noise.dtype
    sched = make_ddim_schedule(num_steps)

    # ---- adapters: once per generation (infer.py:89-91) ----
    concept_text = text_adapter(inputs.clip_hidden, token_index=token_index) if text_adapter is not None else None
    img_tokens = image_adapter(inputs.clip_hidden, token_index=token_index)
    img_tokens_uncond = image_adapter(inputs.clip_hidden_uncond, token_index=token_index)

    latents = inputs.noise * sched.init_noise_sigma
    B = latents.shape[0]
    if mode == "batched":
        ctx = (torch.cat([inputs.text_uncond, inputs.text]).contiguous(),
               torch.cat([img_tokens_uncond, img_tokens]).contiguous())
        ctxs = [ctx]
    elif mode == "two_call":
        ctxs = [(inputs.text_uncond.contiguous(), img_tokens_uncond.contiguous()),
                (inputs.text.contiguous(), img_tokens.contiguous())]
    else:
        ctxs = [(inputs.text.contiguous(), img_tokens.contiguous())]

    set_kv_cache(unet, kv_cache)
    ts_dev = torch.tensor(sched.timesteps, device=dev, dtype=torch.float32)
    rows = 2 * B if mode == "batched" else B
    lat_shape = (rows,) + tuple(latents.shape[1:])
    evals = None
    if use_cuda_graph and dev.type == "cuda":
        evals = [_GraphedUNet(unet, lat_shape, c, dtype, dev) for c in ctxs]
    try:
        for i in range(num_steps):
            t_dev = ts_dev[i:i + 1]
            x_in = torch.cat([latents, latents]) if mode == "batched" else latents
            outs = []
            for j, c in enumerate(ctxs):
                if evals is not None:
                    outs.append(evals[j](x_in, t_dev))
                else:
                    outs.append(unet(x_in, t_dev, encoder_hidden_states=c).sample)
            if mode == "batched":
                eps_u, eps_c = outs[0].chunk(2)
            elif mode == "two_call":
                eps_u, eps_c = outs
            else:
                eps_u = eps_c = outs[0]
            if mode == "cond_only":
                eps = eps_c
            else:
                eps = eps_u + guidance_scale * (eps_c - eps_u)          # infer.py:116
            latents = sched.c_x[i] * latents + sched.c_eps[i] * eps     # DDIM step (eta = 0)
    finally:
        set_kv_cache(unet, False)
    if return_aux:
        return latents, {"concept_text": concept_text, "img_tokens": img_tokens, "img_tokens_uncond": img_tokens_uncond}
    return latents


def synthetic_inputs(batch: int, latent: int = 64, tokens: int = 257, T: int = 5, seed: int = 0,
                     device="cpu", dtype=torch.float32, pin: bool = False) -> GenInputs:
    """Synthetic stand-ins for the encoder outputs (there is no network for real weights / images):
    N(0,1) CLIP hidden states, text states and noise; the "uncond" image is a different fixed draw."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape):
        t = torch.randn(*shape, generator=g, dtype=torch.float32).to(dtype)
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        return t.to(device) if str(device) != "cpu" else t

    clip = [rn(batch, tokens, 1024) for _ in range(T)]
    clip_u = [rn(1, tokens, 1024).expand(batch, tokens, 1024).contiguous() for _ in range(T)]
    if pin and torch.cuda.is_available() and str(device) == "cpu":
        clip_u = [t.pin_memory() for t in clip_u]
    return GenInputs(clip, clip_u, rn(batch, 77, 768), rn(1, 77, 768).expand(batch, 77, 768).contiguous(),
                     rn(batch, 4, latent, latent))
