"""PhotoVerse checkpoint wire format (counterpart of reference ``models/modeling_utils.py:13-50``).

One ``torch.save`` dict:
    "image_adapter" / "text_adapter"   adapter state dicts (``mapping_{i}.{0,1,3,4,6}.{weight,bias}``)
    "cross_attention_adapter"          every unet key that contains "attn2" AND one of "processor", "to_q", "to_k",
                                       "to_v" -- i.e. ``...attn2.processor.to_k_ip.0.weight`` plus, with LoRA, the peft
                                       keys ``...attn2.to_q.base_layer.weight / lora_A.default.weight / lora_B.default.weight``
    "optimizer" (optional), "lora_config" (optional dict: r, lora_alpha, lora_dropout, target_modules)
Files written by the reference load here and vice versa (host-side dict handling only; no arithmetic).
"""
import os
from typing import Optional

import torch

from .lora import DEFAULT_TARGETS, LoraLinear, inject_lora


def cross_attention_state_dict(unet) -> dict:
    out = {}
    for key, value in unet.state_dict().items():
        if "attn2" in key and ("processor" in key or "to_q" in key or "to_k" in key or "to_v" in key):
            out[key] = value
    return out


def save_progress(image_adapter, text_adapter, unet, output_path: str, step: Optional[int] = None,
                  lora_config: Optional[dict] = None, optimizer=None) -> str:
    final = {"image_adapter": image_adapter.state_dict(), "text_adapter": text_adapter.state_dict(),
             "cross_attention_adapter": cross_attention_state_dict(unet)}
    if optimizer is not None:
        final["optimizer"] = optimizer.state_dict()
    if lora_config is not None:
        final["lora_config"] = dict(lora_config)
    name = f"photoverse_{str(step).zfill(6)}.pt" if step is not None else "photoverse.pt"
    path = os.path.join(output_path, name)
    torch.save(final, path)
    return path


def load_photoverse_model(path: str, image_adapter, text_adapter, unet):
    """Returns (image_adapter, text_adapter, unet, lora_config) like the reference; injects LoRA wrappers first when
    the checkpoint carries a ``lora_config`` and the unet has none yet."""
    sd = torch.load(path, map_location="cpu")
    lora_config = sd.get("lora_config")
    if lora_config is not None and not any(isinstance(m, LoraLinear) for m in unet.modules()):
        inject_lora(unet, r=int(lora_config.get("r", 8)), lora_alpha=float(lora_config.get("lora_alpha", 1.0)),
                    lora_dropout=float(lora_config.get("lora_dropout", 0.0)),
                    target_modules=tuple(lora_config.get("target_modules") or DEFAULT_TARGETS))
    if "image_adapter" in sd:
        image_adapter.load_state_dict(sd["image_adapter"])
    if "text_adapter" in sd:
        text_adapter.load_state_dict(sd["text_adapter"])
    if "cross_attention_adapter" in sd:
        unet.load_state_dict(sd["cross_attention_adapter"], strict=False)
    return image_adapter, text_adapter, unet, lora_config
