"""Counterparts of the reference's ``models/unet.py`` helpers, installing the B200 processor.

``set_visual_cross_attention_adapter`` (unet.py:8-35), ``get_visual_cross_attention_values_norm`` (:38-47) and
``set_cross_attention_layers_to_train`` (:50-53) keep the reference's names, arguments and behaviour; they work on
any UNet exposing diffusers' ``attn_processors`` / ``set_attn_processor`` / ``config`` protocol -- a real
``UNet2DConditionModel`` or the SD-1.5-shaped host model in ``photoverse_b200.host.unet_sd15``.
"""
import torch

from .attention_processor import PhotoVerseAttnProcessor2_0


def _default_self_attn_processor():
    try:  # a real diffusers install: use its stock processor exactly as the reference does (unet.py:20-24)
        from diffusers.models.attention_processor import AttnProcessor2_0  # type: ignore
        return AttnProcessor2_0()
    except Exception:
        from .host.unet_sd15 import AttnProcessor2_0
        return AttnProcessor2_0()


def set_visual_cross_attention_adapter(unet, num_tokens=(5,)):
    attn_procs = {}
    for name in unet.attn_processors.keys():
        cross_attention_dim = None if name.endswith("attn1.processor") else unet.config.cross_attention_dim
        if name.startswith("mid_block"):
            hidden_size = unet.config.block_out_channels[-1]
        elif name.startswith("up_blocks"):
            block_id = int(name[len("up_blocks.")])
            hidden_size = list(reversed(unet.config.block_out_channels))[block_id]
        elif name.startswith("down_blocks"):
            block_id = int(name[len("down_blocks.")])
            hidden_size = unet.config.block_out_channels[block_id]
        else:
            raise ValueError(f"unexpected attention processor name {name!r}")
        if cross_attention_dim is None:
            attn_procs[name] = _default_self_attn_processor()
        else:
            attn_procs[name] = PhotoVerseAttnProcessor2_0(
                cross_attention_dim=cross_attention_dim, hidden_size=hidden_size, num_tokens=num_tokens)
    unet.set_attn_processor(attn_procs)
    return unet


def get_visual_cross_attention_values_norm(unet):
    """Stack the ``to_v_ip_norm`` side outputs of the attn2 processors -> [B, n_layers * H * Li] (unet.py:38-47)."""
    attn_values = []
    for name, attn_processor in unet.attn_processors.items():
        if name.endswith("attn1.processor"):
            continue
        attn_values.append(attn_processor.to_v_ip_norm)
    cross_attn_values_norm = torch.stack(attn_values, dim=1)
    bsz = cross_attn_values_norm.shape[0]
    return cross_attn_values_norm.view(bsz, -1)


def set_cross_attention_layers_to_train(unet):
    for name, module in unet.named_modules():
        if "attn2" in name:
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
