"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference image/text adapters.

Follows /root/reference/models/adapters.py:

    :13-28   per token head i two MLPs ``mapping_i`` (CLS token) and ``mapping_patch_i`` (patch tokens):
             Linear(1024,1024) -> LayerNorm(1024) -> LeakyReLU(0.01) -> Linear(1024,1024) -> LayerNorm(1024)
             -> LeakyReLU(0.01) -> Linear(1024,768); all Linear with bias, LayerNorm eps=1e-5 affine.
    :32-37   token_index int : out = mapping_ti(e[:, :1]) + mapping_patch_ti(e[:, 1:]).mean(dim=1, keepdim=True)
    :39-44   token_index None / 'full' : same for every head i on embs[i]; cat over dim=1 -> [B,T,768]

Weights are passed as a flat dict using the reference's own state_dict key names
(``mapping_0.0.weight`` ... ``mapping_patch_4.6.bias``).
"""
from typing import Dict, List, Optional, Union

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.01   # nn.LeakyReLU() default, adapters.py:16
LN_EPS = 1e-5        # nn.LayerNorm default, adapters.py:15


def _mlp(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    h = x @ sd[f"{prefix}.0.weight"].t() + sd[f"{prefix}.0.bias"]
    h = F.layer_norm(h, (h.shape[-1],), sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"], LN_EPS)
    h = F.leaky_relu(h, LRELU_SLOPE)
    h = h @ sd[f"{prefix}.3.weight"].t() + sd[f"{prefix}.3.bias"]
    h = F.layer_norm(h, (h.shape[-1],), sd[f"{prefix}.4.weight"], sd[f"{prefix}.4.bias"], LN_EPS)
    h = F.leaky_relu(h, LRELU_SLOPE)
    return h @ sd[f"{prefix}.6.weight"].t() + sd[f"{prefix}.6.bias"]


def _head(e: torch.Tensor, sd, i: int) -> torch.Tensor:
    return _mlp(e[:, :1], sd, f"mapping_{i}") + _mlp(e[:, 1:], sd, f"mapping_patch_{i}").mean(dim=1, keepdim=True)


def adapter_forward(embs: List[torch.Tensor], sd: Dict[str, torch.Tensor],
                    token_index: Optional[Union[int, str]] = None) -> torch.Tensor:
    if token_index is not None and token_index != "full":      # adapters.py:32-37
        ti = int(token_index)
        return _head(embs[ti], sd, ti)
    return torch.cat([_head(e, sd, i) for i, e in enumerate(embs)], dim=1)   # adapters.py:39-44


def adapter_flops(B: int, T: int, tokens: int = 257, hid: int = 1024, out: int = 768) -> int:
    """Algorithmic FLOPs as the reference computes it."""
    return 2 * B * tokens * T * (2 * hid * hid + hid * out)


def make_state_dict(num_tokens: int, seed: int, clip_dim: int = 1024, out_dim: int = 768,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic adapter weights (default-init ranges; LayerNorm affine perturbed off (1,0)
    so that gamma/beta handling is actually exercised)."""
    from . import detgen
    sd = {}
    s = seed * 1000
    for i in range(num_tokens):
        for name in (f"mapping_{i}", f"mapping_patch_{i}"):
            dims = [(1024, clip_dim), (1024, 1024), (out_dim, 1024)]
            for li, (o, k) in zip((0, 3, 6), dims):
                s += 1
                sd[f"{name}.{li}.weight"] = detgen.uniform_linear(o, k, s, dtype)
                s += 1
                sd[f"{name}.{li}.bias"] = detgen.uniform_bias(o, k, s, dtype)
            for li in (1, 4):
                s += 1
                sd[f"{name}.{li}.weight"] = 1.0 + 0.1 * detgen.uniform((1024,), s, dtype=dtype)
                s += 1
                sd[f"{name}.{li}.bias"] = 0.1 * detgen.uniform((1024,), s, dtype=dtype)
    return sd
