"""TEST INFRASTRUCTURE ONLY -- the oracle plugged into a host UNet: the "reference arm".

``OracleAttnProcessor`` implements the diffusers AttnProcessor call protocol of the reference
(attention_processor.py:245-254) by evaluating ``oracle.processor_oracle.dual_branch_attention`` with the
weights it finds on the ``attn`` module -- plain torch on whatever device / dtype the tensors live on.  It is the
checker of the end-to-end parity tests (final-latent cosine) and the thing timed by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs (the reference itself cannot travel to the GPU box and needs
diffusers/peft, which are not installable).  Never imported by the product package.
"""
import copy
from typing import Optional

import torch
from torch import nn

from . import adapter_oracle
from .processor_oracle import LoraWeights, ProcessorWeights, dual_branch_attention, fusion_weights


def _parts(mod):
    if hasattr(mod, "base_layer"):
        return (mod.base_layer.weight, LoraWeights(mod.lora_A["default"].weight, mod.lora_B["default"].weight,
                                                   float(mod.scaling["default"])))
    return mod.weight, None


class OracleAttnProcessor(nn.Module):
    """Same parameters / names as the reference processor (to_k_ip.0.weight, to_v_ip.0.weight)."""

    def __init__(self, hidden_size, cross_attention_dim=None, num_tokens=(5,), scale=2.0, fusion_rules=(1 / 3, 2 / 3)):
        super().__init__()
        self.scale = [scale]
        self.fusion_rule1, self.fusion_rule2 = fusion_rules
        self.to_k_ip = nn.ModuleList([nn.Linear(cross_attention_dim, hidden_size, bias=False)])
        self.to_v_ip = nn.ModuleList([nn.Linear(cross_attention_dim, hidden_size, bias=False)])
        self.to_v_ip_norm = None
        self.compute_dtype: Optional[torch.dtype] = None     # e.g. torch.float32 to evaluate in fp32 on bf16 inputs

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=2.0,
                 ip_adapter_masks=None):
        text, img = encoder_hidden_states
        if isinstance(img, list):
            img = img[0]
        cd = self.compute_dtype or hidden_states.dtype
        lora = {}
        wq, l = _parts(attn.to_q)
        if l: lora["to_q"] = l
        wk, l = _parts(attn.to_k)
        if l: lora["to_k"] = l
        wv, l = _parts(attn.to_v)
        if l: lora["to_v"] = l
        w = ProcessorWeights(wq, wk, wv, attn.to_out[0].weight, attn.to_out[0].bias, self.to_k_ip[0].weight,
                             self.to_v_ip[0].weight, attn.heads, lora).to(dtype=cd)
        grad = torch.is_grad_enabled()
        u = torch.rand(1).item() if grad else None
        wt, wi = fusion_weights(grad, u, self.scale[0], self.fusion_rule1, self.fusion_rule2)
        y, vn = dual_branch_attention(hidden_states.to(cd), text.to(cd), img.to(cd), w, wt, wi)
        self.to_v_ip_norm = vn.to(hidden_states.dtype)
        return y.to(hidden_states.dtype)


class OracleAdapter(nn.Module):
    """PhotoVerseAdapter-compatible module evaluating oracle.adapter_oracle (same state-dict keys)."""

    def __init__(self, num_tokens=5, clip_embedding_dim=1024, cross_attention_dim=768):
        super().__init__()
        for i in range(num_tokens):
            for name in (f"mapping_{i}", f"mapping_patch_{i}"):
                setattr(self, name, nn.Sequential(
                    nn.Linear(clip_embedding_dim, 1024), nn.LayerNorm(1024), nn.LeakyReLU(),
                    nn.Linear(1024, 1024), nn.LayerNorm(1024), nn.LeakyReLU(), nn.Linear(1024, cross_attention_dim)))

    def forward(self, embs, token_index=None):
        sd = {k: v for k, v in self.state_dict().items()}
        dt = embs[0].dtype
        return adapter_oracle.adapter_forward([e.float() for e in embs], {k: v.float() for k, v in sd.items()},
                                              token_index).to(dt)


def clone_with_oracle_processors(unet, device=None, dtype=None, compute_dtype=None):
    """Deep-copy a host UNet whose attn2 layers carry product processors and swap in oracle processors holding the
    same weights.  ``device`` / ``dtype`` move the copy (e.g. to the CPU in fp32 for the reference arm)."""
    ref = copy.deepcopy(unet)
    for m in ref.modules():
        proc = getattr(m, "processor", None)
        if isinstance(proc, nn.Module) and hasattr(proc, "to_k_ip"):
            C, Dc = proc.to_k_ip[0].weight.shape
            op = OracleAttnProcessor(C, Dc)
            op.to_k_ip[0].weight = proc.to_k_ip[0].weight
            op.to_v_ip[0].weight = proc.to_v_ip[0].weight
            op.compute_dtype = compute_dtype
            m.set_processor(op)
    if hasattr(ref, "set_fused_epilogues"):
        ref.set_fused_epilogues(False)           # the reference arm runs stock PyTorch ops only
    if device is not None or dtype is not None:
        ref.to(device=device, dtype=dtype)
    return ref


def clone_adapter_as_oracle(adapter, device=None, dtype=None):
    ref = OracleAdapter(adapter.num_tokens, adapter.clip_embedding_dim, adapter.cross_attention_dim)
    ref.load_state_dict(adapter.state_dict(), strict=True)
    if device is not None or dtype is not None:
        ref.to(device=device, dtype=dtype)
    return ref
