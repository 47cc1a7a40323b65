"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference dual-branch cross-attention.

Follows /root/reference/models/attention_processor.py (PhotoVerseAttnProcessor2_0) line by line for
the only configuration the PhotoVerse callers ever exercise (3-D hidden states, tuple
``encoder_hidden_states``, no masks / norms / residual):

    :297        Q   = to_q(X)                                   (LoRA-wrapped when injected)
    :304-305    K_t = to_k(text) ; V_t = to_v(text)            (LoRA-wrapped when injected)
    :307-313    split heads, head_dim = C // heads
    :317-319    O_t = softmax(Q K_t^T / sqrt(d)) V_t            (SDPA, default scale)
    :392-396    K_i = to_k_ip(img) ; V_i = to_v_ip(img)
    :397        to_v_ip_norm = ||V_i||_2 over head_dim, keepdim  (side output, every call)
    :400-407    O_i = softmax(Q K_i^T / sqrt(d)) V_i
    :411-420    no-grad: O = O_t + O_i ; grad: u~U[0,1): u<r1 -> scale*O_t ; u>r2 -> scale*O_i ; else sum
    :423-425    Y = to_out[0](O)  (+bias) ; dropout(p=0)
    :433        Y / rescale_output_factor (== 1.0 for SD-1.5 attn2)

LoRA (peft==0.10.0 ``lora.Linear.forward``; dependency not vendored in the reference, restated from
its published algorithm):  y = W x + scaling * B(A(dropout(x))),  scaling = lora_alpha / r.

Everything is plain differentiable torch, so ``torch.autograd`` through these functions is the
oracle for the backward kernels too.  Parity status: pinned against the verbatim reference import
(see oracle/make_golden.py); the reference itself ships no test vectors.
"""
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch


@dataclass
class LoraWeights:
    A: torch.Tensor          # [r, in]
    B: torch.Tensor          # [out, r]
    scaling: float           # lora_alpha / r


@dataclass
class ProcessorWeights:
    """Parameters one attn2 layer + its PhotoVerse processor own (names as in the reference)."""
    to_q: torch.Tensor       # [C, C]     attn.to_q.weight   (bias-free)
    to_k: torch.Tensor       # [C, 768]   attn.to_k.weight
    to_v: torch.Tensor       # [C, 768]   attn.to_v.weight
    to_out_w: torch.Tensor   # [C, C]     attn.to_out[0].weight
    to_out_b: torch.Tensor   # [C]        attn.to_out[0].bias
    to_k_ip: torch.Tensor    # [C, 768]   processor.to_k_ip[0].weight (attention_processor.py:51-53)
    to_v_ip: torch.Tensor    # [C, 768]   processor.to_v_ip[0].weight (attention_processor.py:54-56)
    heads: int = 8
    lora: Dict[str, LoraWeights] = field(default_factory=dict)   # keys among 'to_q','to_k','to_v'

    def to(self, dtype=None, device=None):
        def cv(t):
            return t.to(dtype=dtype, device=device)
        return ProcessorWeights(
            cv(self.to_q), cv(self.to_k), cv(self.to_v), cv(self.to_out_w), cv(self.to_out_b),
            cv(self.to_k_ip), cv(self.to_v_ip), self.heads,
            {k: LoraWeights(cv(v.A), cv(v.B), v.scaling) for k, v in self.lora.items()})


def lora_linear(x: torch.Tensor, w: torch.Tensor, lora: Optional[LoraWeights],
                drop: Optional[Tuple[torch.Tensor, float]] = None) -> torch.Tensor:
    """peft 0.10.0 ``lora.Linear.forward``: ``result = base_layer(x); result += lora_B(lora_A(dropout(x))) * scaling``.
    ``drop`` = (keep_mask, p) makes the ``nn.Dropout(p)`` of training mode explicit (train.py:264-269 default p = 0.1):
    dropout(x) = x * keep_mask / (1 - p); None = eval mode / p = 0 (nn.Identity)."""
    y = x @ w.t()
    if lora is not None:
        xd = x
        if drop is not None:
            mask, p = drop
            xd = x * mask.to(x.dtype).view_as(x) / (1.0 - p)
        y = y + (xd @ lora.A.t()) @ lora.B.t() * lora.scaling
    return y


def fusion_weights(grad_enabled: bool, u: Optional[float], scale: float = 2.0,
                   rule1: float = 1 / 3, rule2: float = 2 / 3) -> Tuple[float, float]:
    """(w_text, w_image) of attention_processor.py:411-420."""
    if not grad_enabled:
        return 1.0, 1.0
    assert u is not None
    if u < rule1:
        return float(scale), 0.0
    if u > rule2:
        return 0.0, float(scale)
    return 1.0, 1.0


def _sdpa(q, k, v):
    # F.scaled_dot_product_attention(q,k,v) with default scale 1/sqrt(head_dim), no mask, no dropout
    d = q.shape[-1]
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    return torch.softmax(s, dim=-1) @ v


def dual_branch_attention(x: torch.Tensor, text: torch.Tensor, img: torch.Tensor,
                          w: ProcessorWeights, w_text: float = 1.0, w_img: float = 1.0,
                          drop: Optional[Dict[str, Tuple[torch.Tensor, float]]] = None):
    """Returns (Y [B,S,C], to_v_ip_norm [B,H,Li,1]).  ``drop``: per-projection (keep_mask, p) of the LoRA dropout."""
    B, S, C = x.shape
    H = w.heads
    d = C // H
    drop = drop or {}

    q = lora_linear(x, w.to_q, w.lora.get("to_q"), drop.get("to_q"))       # :297
    k = lora_linear(text, w.to_k, w.lora.get("to_k"), drop.get("to_k"))    # :304
    v = lora_linear(text, w.to_v, w.lora.get("to_v"), drop.get("to_v"))    # :305

    def split(t):                                                       # :310-313
        return t.view(B, -1, H, d).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    o_text = _sdpa(q, k, v).transpose(1, 2).reshape(B, -1, C)           # :317-322

    ik = split(img @ w.to_k_ip.t())                                     # :392,:395
    iv = split(img @ w.to_v_ip.t())                                     # :393,:396
    v_ip_norm = torch.norm(iv, dim=-1, keepdim=True)                    # :397
    o_img = _sdpa(q, ik, iv).transpose(1, 2).reshape(B, -1, C)          # :400-407

    if w_text == 1.0 and w_img == 1.0:
        o = o_text + o_img                                              # :412 / :420
    elif w_img == 0.0:
        o = w_text * o_text                                             # :416
    elif w_text == 0.0:
        o = w_img * o_img                                               # :418
    else:  # not reachable from the reference; kept for kernel tests of arbitrary weights
        o = w_text * o_text + w_img * o_img

    y = o @ w.to_out_w.t() + w.to_out_b                                 # :423
    return y, v_ip_norm


def segment_softmax_form(x, text, img, w: ProcessorWeights, w_text: float = 1.0, w_img: float = 1.0):
    """The algebraically identical form the CUDA kernel uses: one QK^T over the
    concatenated keys, per-segment normalisation folded into P, ONE PV contraction.  Kept here so
    ``tests/`` can assert identity with :func:`dual_branch_attention` on the CPU."""
    B, S, C = x.shape
    H = w.heads
    d = C // H
    Lt = text.shape[1]
    q = lora_linear(x, w.to_q, w.lora.get("to_q")).view(B, S, H, d).transpose(1, 2)
    kc = torch.cat([lora_linear(text, w.to_k, w.lora.get("to_k")), img @ w.to_k_ip.t()], 1)
    vc = torch.cat([lora_linear(text, w.to_v, w.lora.get("to_v")), img @ w.to_v_ip.t()], 1)
    kc = kc.view(B, -1, H, d).transpose(1, 2)
    vc = vc.view(B, -1, H, d).transpose(1, 2)
    s = (q @ kc.transpose(-1, -2)) * (d ** -0.5)
    p = torch.cat([w_text * torch.softmax(s[..., :Lt], -1), w_img * torch.softmax(s[..., Lt:], -1)], -1)
    o = (p @ vc).transpose(1, 2).reshape(B, S, C)
    y = o @ w.to_out_w.t() + w.to_out_b
    v_ip_norm = torch.norm(vc[:, :, Lt:], dim=-1, keepdim=True)
    return y, v_ip_norm


def gather_values_norm(norms):
    """models/unet.py:38-47 -- stack the per-layer side outputs -> [B, n_layers*H*Li]."""
    stacked = torch.stack(list(norms), dim=1)
    return stacked.view(stacked.shape[0], -1)


def processor_flops(B, S, C, Lt, Li, Dc=768, r=0):
    """Algorithmic FLOPs of one processor call, SURVEY.md §8(d) / BASELINE.md §3."""
    f = 4 * B * S * C * C + 4 * B * S * C * (Lt + Li) + 4 * B * (Lt + Li) * Dc * C
    if r:
        f += 4 * B * S * C * r + 4 * B * Lt * r * (Dc + C)
    return f
