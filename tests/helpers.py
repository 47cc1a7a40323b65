"""Shared test helpers: build product-side layers from the oracle's deterministic case table."""
import json
import os

import numpy as np
import torch

from oracle import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def manifest():
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
        return json.load(f)


def build_product_layer(case: cases.ProcCase, device, master_dtype=torch.float32):
    """photoverse_b200 Attention module + PhotoVerse processor holding the case's weights."""
    from photoverse_b200.attention_processor import PhotoVerseAttnProcessor2_0
    from photoverse_b200.host.unet_sd15 import Attention
    from photoverse_b200.lora import LoraLinear

    w = cases.proc_weights(case, torch.float32)
    attn = Attention(case.C, case.Dc, case.H)
    proc = PhotoVerseAttnProcessor2_0(hidden_size=case.C, cross_attention_dim=case.Dc, num_tokens=(5,))
    with torch.no_grad():
        attn.to_q.weight.copy_(w.to_q)
        attn.to_k.weight.copy_(w.to_k)
        attn.to_v.weight.copy_(w.to_v)
        attn.to_out[0].weight.copy_(w.to_out_w)
        attn.to_out[0].bias.copy_(w.to_out_b)
        proc.to_k_ip[0].weight.copy_(w.to_k_ip)
        proc.to_v_ip[0].weight.copy_(w.to_v_ip)
    for name, lw in w.lora.items():
        r = lw.A.shape[0]
        wrapped = LoraLinear(getattr(attn, name), r=r, lora_alpha=lw.scaling * r)
        with torch.no_grad():
            wrapped.lora_A["default"].weight.copy_(lw.A)
            wrapped.lora_B["default"].weight.copy_(lw.B)
        setattr(attn, name, wrapped)
    attn.set_processor(proc)
    attn.to(device=device, dtype=master_dtype)
    return attn, proc


def force_fusion_seed(w_text, w_img):
    """Seed the global CPU RNG so that the processor's single torch.rand(1) selects the wanted branch."""
    want = 0.5 if (w_text, w_img) == (1.0, 1.0) else (0.1 if w_img == 0.0 else 0.9)
    for seed in range(10000):
        torch.manual_seed(seed)
        u = torch.rand(1).item()
        if abs(u - want) < 0.1:
            torch.manual_seed(seed)
            return u
    raise RuntimeError("no seed found")
