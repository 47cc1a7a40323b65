"""Drop-in processor for the self-attention (``attn1``) layers of the transformer blocks the PhotoVerse processors live
in.

The reference installs diffusers' stock ``AttnProcessor2_0`` there (reference models/unet.py:20-24):
``softmax(Q K^T / sqrt(d)) V`` through ``F.scaled_dot_product_attention``.  :class:`SelfAttnProcessor` has the same call
protocol and produces the same result (bf16, within 2e-2 max-abs of an fp32 evaluation) with the attention itself on
``pv_self_attn_fwd`` (photoverse_b200/csrc/pv_sattn.cu); the four projections stay what they are in the reference --
plain ``nn.Linear`` GEMMs -- with q / k / v evaluated as ONE fused ``[3C, C]`` GEMM whose three column blocks the kernel
reads in place.

Opt-in (``install_self_attention(unet)``): measured on B200 the kernel is at 0.55-1.0x of cuDNN's sm100 flash-attention
forward, so the benchmark line keeps the stock processor, like the reference.  Inference only: with
gradients enabled, or for inputs the kernel does not take (fp32, head_dim not in {40, 80, 160}, masks, cross-attention),
the call goes to the stock SDPA path of the reference -- these layers are not part of the hot path and are never
trained by PhotoVerse.
"""
import torch
import torch.nn.functional as F

from . import ops


class SelfAttnProcessor:
    def __init__(self):
        self._wqkv = None
        self._key = None

    def _fused_weight(self, attn):
        ws = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight)
        key = tuple((w.data_ptr(), w._version, w.dtype, w.device) for w in ws)
        if key != self._key:
            self._wqkv = torch.cat([w.detach() for w in ws], dim=0).contiguous()
            self._key = key
        return self._wqkv

    @staticmethod
    def _stock(attn, hidden_states):
        B, S, C = hidden_states.shape
        h = attn.heads
        q = attn.to_q(hidden_states).view(B, S, h, C // h).transpose(1, 2)
        k = attn.to_k(hidden_states).view(B, S, h, C // h).transpose(1, 2)
        v = attn.to_v(hidden_states).view(B, S, h, C // h).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        return o.transpose(1, 2).reshape(B, S, C).to(q.dtype)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kwargs):
        if encoder_hidden_states is not None or attention_mask is not None:
            raise NotImplementedError("SelfAttnProcessor handles the unmasked self-attention (attn1) layers only")
        if hidden_states.dim() != 3:
            raise NotImplementedError("SelfAttnProcessor expects [batch, tokens, channels] hidden states")
        B, S, C = hidden_states.shape
        d = C // attn.heads
        needs_grad = torch.is_grad_enabled() and (hidden_states.requires_grad or attn.to_q.weight.requires_grad)
        if (needs_grad or hidden_states.dtype != torch.bfloat16 or not hidden_states.is_cuda or d not in (40, 80, 160)
                or getattr(attn.to_q, "bias", None) is not None):
            o = self._stock(attn, hidden_states)
        else:
            qkv = F.linear(hidden_states, self._fused_weight(attn))          # [B, S, 3C], one GEMM
            o = ops.self_attn(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], attn.heads)
        return attn.to_out[1](attn.to_out[0](o))


def install_self_attention(unet):
    """Put :class:`SelfAttnProcessor` on every ``attn1`` module of ``unet`` (the attn2 processors are left alone)."""
    procs = dict(unet.attn_processors)
    for name in procs:
        if name.endswith("attn1.processor"):
            procs[name] = SelfAttnProcessor()
    unet.set_attn_processor(procs)
    return unet
