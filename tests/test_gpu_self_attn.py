"""GPU (B200): the attn1 self-attention kernel (pv_self_attn_fwd; reference models/unet.py:20-24 -> diffusers
AttnProcessor2_0 -> F.scaled_dot_product_attention) against an fp32 evaluation of the same bf16 inputs.
Tolerance: 2e-2 max-abs (north_star's bf16 attention tolerance); cuDNN's own bf16 SDPA is 1e-3 .. 4e-3 on these inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(q, k, v, H):
    B, S, C = q.shape
    f = lambda t: t.float().reshape(B, S, H, C // H).transpose(1, 2)
    return F.scaled_dot_product_attention(f(q), f(k), f(v)).transpose(1, 2).reshape(B, S, C)


@pytest.mark.parametrize("B,S,C", [
    (2, 4096, 320), (16, 1024, 640), (16, 256, 1280), (16, 64, 1280),       # the four SD-1.5 shapes (latent 64^2)
    (1, 2304, 320), (2, 576, 640), (3, 144, 1280), (2, 36, 1280),           # latent 48^2: ragged tiles, odd tile counts
    (1, 1, 320), (2, 129, 320), (1, 200, 640), (1, 65, 1280),               # one row / one row past a tile
], ids=lambda v: str(v))
def test_self_attention_matches_fp32_reference(cuda_device, B, S, C):
    from photoverse_b200 import ops
    H = 8
    g = torch.Generator().manual_seed(S * 7 + C)
    qkv = (torch.randn(B, S, 3 * C, generator=g) * 1.5).to(cuda_device, torch.bfloat16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    out = ops.self_attn(q, k, v, H)
    ref = _ref(q, k, v, H)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    print(f"B{B} S{S} C{C}: max-abs {err:.3e}")
    assert err <= 2e-2


def test_self_attention_peaked_rows_and_separate_tensors(cuda_device):
    """Large logits (the running maximum moves by far more than the lazy-rescaling threshold from tile to tile, and most
    exponentials underflow) and q / k / v as three separate contiguous tensors."""
    from photoverse_b200 import ops
    B, S, C, H = 2, 1024, 320, 8
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(B, S, C, generator=g) * 6).to(cuda_device, torch.bfloat16)
    k = (torch.randn(B, S, C, generator=g) * 6).to(cuda_device, torch.bfloat16)
    k[:, 700:] *= 3.0                                   # later tiles hold much larger maxima
    v = torch.randn(B, S, C, generator=g).to(cuda_device, torch.bfloat16)
    out = ops.self_attn(q, k, v, H)
    ref = _ref(q, k, v, H)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 3e-2


def test_self_attention_is_deterministic_and_option_independent(cuda_device):
    from photoverse_b200 import _lib, ops
    B, S, C, H = 2, 1024, 640, 8
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B, S, 3 * C, generator=g).to(cuda_device, torch.bfloat16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    a = ops.self_attn(q, k, v, H)
    b = ops.self_attn(q, k, v, H)
    assert torch.equal(a, b)
    ref = _ref(q, k, v, H)
    try:
        for pf in (0, 4):
            _lib.set_option("sattn_poly", pf)
            assert (ops.self_attn(q, k, v, H).float() - ref).abs().max().item() <= 2e-2
    finally:
        _lib.set_option("sattn_poly", 2)


def test_self_attention_processor_matches_stock(cuda_device):
    """The opt-in attn1 processor against the stock SDPA processor on the same Attention module."""
    from photoverse_b200.host.unet_sd15 import Attention, AttnProcessor2_0
    from photoverse_b200.self_attention import SelfAttnProcessor
    torch.manual_seed(0)
    attn = Attention(640, None, 8).to(cuda_device, torch.bfloat16).eval().requires_grad_(False)
    x = torch.randn(2, 1024, 640, device=cuda_device, dtype=torch.bfloat16)
    with torch.no_grad():
        want = AttnProcessor2_0()(attn, x)
        attn.set_processor(SelfAttnProcessor())
        got = attn(x)
    assert (got.float() - want.float()).abs().max().item() <= 2e-2
    # gradients requested -> the stock path, bit-identical to it
    xg = x.clone().requires_grad_(True)
    assert torch.equal(attn(xg), AttnProcessor2_0()(attn, xg))


def test_self_attention_rejects_what_it_cannot_do(cuda_device):
    from photoverse_b200 import _lib, ops
    x = torch.randn(1, 64, 3 * 512, device=cuda_device, dtype=torch.bfloat16)
    with pytest.raises(_lib.PhotoverseB200Error):
        ops.self_attn(x[..., :512], x[..., 512:1024], x[..., 1024:], 8)          # head_dim 64
    y = torch.randn(1, 64, 320, device=cuda_device)
    with pytest.raises(_lib.PhotoverseB200Error):
        ops.self_attn(y, y, y, 8)                                               # fp32
