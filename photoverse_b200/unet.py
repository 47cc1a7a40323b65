"""Counterparts of the reference's ``models/unet.py`` helpers, installing the B200 processor.

``set_visual_cross_attention_adapter`` (unet.py:8-35), ``get_visual_cross_attention_values_norm`` (:38-47) and
``set_cross_attention_layers_to_train`` (:50-53) keep the reference's names, arguments and behaviour; they work on
any UNet exposing diffusers' ``attn_processors`` / ``set_attn_processor`` / ``config`` protocol -- a real
``UNet2DConditionModel`` or the SD-1.5-shaped host model in ``photoverse_b200.host.unet_sd15``.
"""
import re

import torch

from .attention_processor import PhotoVerseAttnProcessor2_0


def _default_self_attn_processor():
    try:  # a real diffusers install: use its stock processor exactly as the reference does (unet.py:20-24)
        from diffusers.models.attention_processor import AttnProcessor2_0  # type: ignore
        return AttnProcessor2_0()
    except Exception:
        from .host.unet_sd15 import AttnProcessor2_0
        return AttnProcessor2_0()


_BLOCK_RE = re.compile(r"^(down_blocks|up_blocks)\.(\d+)\.|^(mid_block)\.")


def _layer_width(processor_key: str, widths) -> int:
    """Channel width C of the transformer block a processor key belongs to: down block i -> widths[i], up block i ->
    widths[-1 - i], mid block -> widths[-1] (what reference unet.py:12-19 derives from ``block_out_channels``)."""
    m = _BLOCK_RE.match(processor_key)
    if m is None:
        raise ValueError(f"unexpected attention processor name {processor_key!r}")
    if m.group(3):
        return widths[-1]
    idx = int(m.group(2))
    return widths[idx] if m.group(1) == "down_blocks" else widths[len(widths) - 1 - idx]


def _is_self_attention(processor_key: str) -> bool:
    return processor_key.endswith("attn1.processor")


def set_visual_cross_attention_adapter(unet, num_tokens=(5,)):
    """Install the B200 dual-branch processor on every cross-attention (``attn2``) module and a stock SDPA processor on
    every self-attention (``attn1``) module -- the contract of reference unet.py:8-35 (same name / arguments / result)."""
    widths = tuple(unet.config.block_out_channels)
    context_dim = unet.config.cross_attention_dim
    unet.set_attn_processor({
        key: (_default_self_attn_processor() if _is_self_attention(key) else
              PhotoVerseAttnProcessor2_0(hidden_size=_layer_width(key, widths), cross_attention_dim=context_dim,
                                         num_tokens=num_tokens))
        for key in unet.attn_processors})
    return unet


def get_visual_cross_attention_values_norm(unet):
    """The regulariser input of train.py:512-513: every attn2 processor's ``to_v_ip_norm`` side output [B, H, Li, 1],
    layers in ``attn_processors`` order -> [B, n_layers * H * Li] (reference unet.py:38-47)."""
    norms = [proc.to_v_ip_norm for key, proc in unet.attn_processors.items() if not _is_self_attention(key)]
    return torch.stack(norms, dim=1).flatten(1)


def set_cross_attention_layers_to_train(unet):
    """``.train()`` on every module under an ``attn2`` (activates peft's LoRA dropout; reference unet.py:50-53)."""
    for module in (m for n, m in unet.named_modules() if "attn2" in n):
        module.train()


def set_kv_cache(unet, enabled: bool, static: bool = False):
    """Enable / clear / disable K/V-projection caching on every PhotoVerse processor of ``unet`` (used by the
    denoise loop: encoder_hidden_states are constant across steps, models/infer.py:89-98)."""
    for proc in unet.attn_processors.values():
        if isinstance(proc, PhotoVerseAttnProcessor2_0):
            proc.enable_kv_cache(enabled, static=static)


def prepare_kv(unet, text, img):
    """Project + pack K/V of every attn2 layer for the given (text, image) contexts (once per generation)."""
    for m in unet.modules():
        proc = getattr(m, "processor", None)
        if isinstance(proc, PhotoVerseAttnProcessor2_0):
            proc.prepare_kv(m, text, img)
