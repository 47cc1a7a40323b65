"""GPU (B200): the inference-time epilogue kernels of the host UNet (csrc/pv_backbone.cu) against the
stock PyTorch ops they replace -- GroupNorm (+ SiLU) on channels-last activations and the GEGLU product -- through the
C ABI, and the UNet with / without them."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, C, H, W): every GroupNorm width of the SD-1.5 UNet (320 ... 2560 with the skip concatenations), ragged pixel counts
GN_SHAPES = [(2, 320, 64, 64), (2, 640, 32, 32), (2, 960, 32, 32), (1, 1280, 16, 16), (1, 1920, 16, 16), (2, 2560, 8, 8),
             (3, 1280, 8, 8), (2, 320, 7, 9), (1, 640, 1, 3), (16, 640, 64, 64)]


@pytest.mark.parametrize("shape", GN_SHAPES, ids=[f"B{b}_C{c}_{h}x{w}" for b, c, h, w in GN_SHAPES])
@pytest.mark.parametrize("silu,with_add", [(True, False), (False, False), (True, True)], ids=["silu", "plain", "silu-addend"])
def test_group_norm_nhwc_matches_torch(cuda_device, shape, silu, with_add):
    from photoverse_b200 import ops
    B, C, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + C + H)
    # per-channel offsets far from zero: the statistics must survive |mean| >> std
    x = (torch.randn(B, C, H, W, generator=g) * 0.7 + torch.randn(1, C, 1, 1, generator=g) * 4.0 + 3.0)
    x = x.to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gamma = (1.0 + 0.3 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    eps = 1e-5 if silu else 1e-6
    # ResnetBlock2D: conv1 bias + time-embedding projection, one value per (sample, channel), added before the statistics
    add = (torch.randn(B, C, generator=g) * 1.5).to(cuda_device) if with_add else None
    y = ops.group_norm_nhwc(x, gamma, beta, 32, eps, silu, add)
    assert y.shape == x.shape and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    xin = x.float() if add is None else x.float() + add[:, :, None, None]
    ref = F.group_norm(xin, 32, gamma, beta, eps)              # fp32 torch reference on the same bf16 input
    if silu:
        ref = F.silu(ref)
    err = (y.float() - ref).abs()
    tol = 1e-2 + 1e-2 * ref.abs()                              # bf16 output: 2^-8 relative + the affine cancellation
    assert bool((err <= tol).all()), f"max err {err.max().item():.4e} at |ref| {ref.abs().flatten()[err.argmax()].item():.3f}"
    # bit-reproducible (fixed summation order) ...
    assert torch.equal(y, ops.group_norm_nhwc(x, gamma, beta, 32, eps, silu, add))
    # ... and a sample's result does not depend on the batch it is normalised in (generation shards by sample)
    if B > 1:
        y1 = ops.group_norm_nhwc(x[1:2].contiguous(memory_format=torch.channels_last), gamma, beta, 32, eps, silu,
                                 None if add is None else add[1:2].contiguous())
        assert torch.equal(y1[0], y[1])


BWD_SHAPES = [(2, 320, 64, 64), (2, 640, 32, 32), (1, 1920, 16, 16), (2, 2560, 8, 8), (3, 1280, 5, 7), (16, 320, 64, 64)]


@pytest.mark.parametrize("shape", BWD_SHAPES, ids=[f"B{b}_C{c}_{h}x{w}" for b, c, h, w in BWD_SHAPES])
@pytest.mark.parametrize("silu", [True, False], ids=["silu", "plain"])
def test_group_norm_nhwc_backward_matches_autograd(cuda_device, shape, silu):
    """Input gradient of the training-time GroupNorm (+ SiLU) against torch.autograd through the fp32 torch ops on the same
    bf16 input and upstream gradient (the affine is frozen on the PhotoVerse path: no gamma / beta gradients)."""
    from photoverse_b200 import ops
    B, C, H, W = shape
    g = torch.Generator().manual_seed(C + H + B)
    x = (torch.randn(B, C, H, W, generator=g) * 0.8 + torch.randn(1, C, 1, 1, generator=g) * 2.0 + 1.0)
    x = x.to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, C, H, W, generator=g).to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gamma = (1.0 + 0.3 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    y, stats = ops.group_norm_nhwc(x, gamma, beta, 32, 1e-5, silu, None, save_stats=True)
    assert torch.equal(y, ops.group_norm_nhwc(x, gamma, beta, 32, 1e-5, silu))          # saving statistics changes nothing
    dx = ops.group_norm_nhwc_bwd(x, dy, stats, gamma, beta, 32, silu)
    assert dx.dtype == torch.bfloat16 and dx.is_contiguous(memory_format=torch.channels_last)
    xr = x.float().requires_grad_(True)
    ref = F.group_norm(xr, 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref.backward(dy.float())
    err = (dx.float() - xr.grad).abs()
    scale = xr.grad.abs().max().item()
    assert err.max().item() <= 1e-2 * scale + 1e-2 * 0, f"max err {err.max().item():.3e} vs gradient scale {scale:.3e}"
    rel = ((dx.float() - xr.grad).norm() / xr.grad.norm()).item()
    assert rel <= 6e-3, rel
    assert torch.equal(dx, ops.group_norm_nhwc_bwd(x, dy, stats, gamma, beta, 32, silu))      # fixed summation order
    m = x.float().view(B, 32, C // 32, H * W).mean(dim=(2, 3))
    assert torch.allclose(stats[..., 0], m, atol=2e-3, rtol=1e-3)


def test_training_group_norm_function_in_the_unet(cuda_device):
    """Grad mode, channels-last bf16 UNet: the GroupNorms run the fused forward / backward pair (launch count) and the
    gradients that reach the trainable set agree with the stock ops."""
    import photoverse_b200 as pv
    from photoverse_b200 import _lib
    from photoverse_b200.host.unet_sd15 import UNetSD15
    torch.manual_seed(4)
    unet = UNetSD15()
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(5,))
    unet.requires_grad_(False).eval().to(device=cuda_device, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    trainable = [p for n, p in unet.named_parameters() if "to_k_ip" in n or "to_v_ip" in n]
    for p in trainable:
        p.requires_grad_(True)
    x = torch.randn(2, 4, 32, 32, device=cuda_device, dtype=torch.bfloat16)
    text = torch.randn(2, 77, 768, device=cuda_device, dtype=torch.bfloat16)
    img = torch.randn(2, 5, 768, device=cuda_device, dtype=torch.bfloat16)
    t = torch.tensor([500], device=cuda_device)
    target = torch.randn(2, 4, 32, 32, device=cuda_device, dtype=torch.bfloat16)

    def grads(fused):
        unet.set_fused_epilogues(fused)
        for p in trainable:
            p.grad = None
        torch.manual_seed(11)                       # the fusion rule draws one random number per layer in grad mode
        n0 = _lib.launch_count()
        out = unet(x, t, (text, img)).sample
        F.mse_loss(out.float(), target.float()).backward()
        # a branch the fusion rule dropped leaves its parameters without a gradient (same draw on both arms: same seed)
        flat = [(p.grad if p.grad is not None else torch.zeros_like(p)).float().flatten() for p in trainable]
        return torch.cat(flat), _lib.launch_count() - n0, out.detach()

    grads(False)                                    # first call packs the processors' weights: keep it out of the counts
    g_stock, n_stock, o_stock = grads(False)
    g_fused, n_fused, o_fused = grads(True)
    # 61 GroupNorms x 2 launches forward, 2 more backward for those downstream of the first trainable layer, 22 residual
    # sums, 48 LayerNorms forward + backward
    assert n_fused >= n_stock + 2 * 61 + 2 * 40 + 22 + 48 + 40, (n_stock, n_fused)
    cos = F.cosine_similarity(g_stock.double(), g_fused.double(), dim=0).item()
    print(f"training UNet, fused vs stock GroupNorm: gradient cosine {cos:.5f}, output cosine "
          f"{F.cosine_similarity(o_stock.double().flatten(), o_fused.double().flatten(), dim=0).item():.6f}; native launches {n_stock} -> {n_fused}")
    assert cos >= 0.99


@pytest.mark.parametrize("B,HW,C", [(2, 1, 64), (3, 2, 256), (1, 5, 32)])
def test_group_norm_nhwc_blocks_narrower_than_the_group_count(cuda_device, B, HW, C):
    """Tiny activations ([B, HW, C] form): the block has C / 8 x min(HW, 16) threads, fewer than the 32 groups."""
    from photoverse_b200 import ops
    g = torch.Generator().manual_seed(C + HW)
    x = (torch.randn(B, HW, C, generator=g) + 2.0).to(cuda_device, torch.bfloat16)
    dy = torch.randn(B, HW, C, generator=g).to(cuda_device, torch.bfloat16)
    gamma = (1.0 + 0.3 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    y, stats = ops.group_norm_nhwc(x, gamma, beta, 32, 1e-5, True, None, save_stats=True)
    dx = ops.group_norm_nhwc_bwd(x, dy, stats, gamma, beta, 32, True)
    xr = x.float().permute(0, 2, 1).contiguous().requires_grad_(True)          # [B, C, HW] for F.group_norm
    ref = F.silu(F.group_norm(xr, 32, gamma, beta, 1e-5))
    ref.backward(dy.float().permute(0, 2, 1))
    assert bool(((y.float() - ref.detach().permute(0, 2, 1)).abs() <= 1e-2 + 1e-2 * ref.detach().permute(0, 2, 1).abs()).all())
    gref = xr.grad.permute(0, 2, 1)
    assert ((dx.float() - gref).norm() / gref.norm().clamp_min(1e-6)).item() <= 1e-2


def test_group_norm_nhwc_rejects_what_it_cannot_do(cuda_device):
    from photoverse_b200 import _lib, ops
    x = torch.randn(2, 320, 8, 8, device=cuda_device, dtype=torch.bfloat16)
    w = torch.ones(320, device=cuda_device)
    with pytest.raises(_lib.PhotoverseB200Error):
        ops.group_norm_nhwc(x, w, w, 32, 1e-5, True)                       # NCHW-contiguous
    with pytest.raises(_lib.PhotoverseB200Error):
        ops.group_norm_nhwc(x.float().contiguous(memory_format=torch.channels_last), w, w, 32, 1e-5, True)   # fp32
    assert _lib.lib().pv_group_norm_nhwc_ws_bytes(2, 64, 324, 32) == -1      # C % 8, C % groups


@pytest.mark.parametrize("M,N", [(16 * 4096, 1280), (16 * 1024, 2560), (4096, 5120), (1024, 5120), (7, 16)])
def test_geglu_matches_torch(cuda_device, M, N):
    from photoverse_b200 import ops
    g = torch.Generator().manual_seed(N + M)
    h = (torch.randn(M, 2 * N, generator=g) * 1.5).to(cuda_device, torch.bfloat16)
    y = ops.geglu(h)
    a, gate = h.chunk(2, dim=-1)
    ref = a * F.gelu(gate)                                     # the stock two-kernel bf16 sequence
    assert y.shape == ref.shape
    # same rounding points (gelu rounded to bf16, then the product): differences are last-place flips of erf
    err = (y.float() - ref.float()).abs()
    assert bool((err <= 2e-3 + 8e-3 * ref.float().abs()).all()), err.max().item()
    ref32 = a.float() * F.gelu(gate.float())
    assert bool(((y.float() - ref32).abs() <= 2e-3 + 1.2e-2 * ref32.abs()).all())


@pytest.mark.parametrize("rows,C", [(16 * 4096, 320), (16 * 1024, 640), (4096, 1280), (77, 1280), (3, 8), (5, 328)])
def test_layer_norm_matches_torch(cuda_device, rows, C):
    from photoverse_b200 import ops
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 2.0 + torch.randn(rows, 1, generator=g) * 3.0).to(cuda_device, torch.bfloat16)
    gamma = (1.0 + 0.3 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    y = ops.layer_norm(x, gamma, beta, 1e-5)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    err = (y.float() - ref).abs()
    assert y.dtype == torch.bfloat16 and bool((err <= 1e-2 + 1e-2 * ref.abs()).all()), err.max().item()
    stock = F.layer_norm(x, (C,), gamma.to(torch.bfloat16), beta.to(torch.bfloat16), 1e-5)
    assert (y.float() - ref).abs().max() <= (stock.float() - ref).abs().max() + 1e-2     # no worse than the op it replaces
    # residual sum in the same pass: the sum is the bf16 sum, the norm is the norm of that sum (bit for bit)
    r = torch.randn(rows, C, generator=g).to(cuda_device, torch.bfloat16)
    s2, y2 = ops.add_layer_norm(x, r, gamma, beta, 1e-5)
    assert torch.equal(s2, x + r) and torch.equal(y2, ops.layer_norm(x + r, gamma, beta, 1e-5))


@pytest.mark.parametrize("rows,C", [(16 * 4096, 320), (16 * 1024, 640), (4096, 1280), (5, 328), (3, 8)])
def test_layer_norm_backward_matches_autograd(cuda_device, rows, C):
    from photoverse_b200 import ops
    g = torch.Generator().manual_seed(rows * 3 + C)
    x = (torch.randn(rows, C, generator=g) * 2.0 + torch.randn(rows, 1, generator=g) * 3.0).to(cuda_device, torch.bfloat16)
    dy = torch.randn(rows, C, generator=g).to(cuda_device, torch.bfloat16)
    gamma = (1.0 + 0.3 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    dx = ops.layer_norm_bwd(x, dy, gamma, 1e-5)
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (C,), gamma, beta, 1e-5).backward(dy.float())
    rel = ((dx.float() - xr.grad).norm() / xr.grad.norm()).item()
    assert dx.dtype == torch.bfloat16 and rel <= 5e-3, rel
    assert (dx.float() - xr.grad).abs().max().item() <= 1e-2 * xr.grad.abs().max().item() + 1e-3


@pytest.mark.parametrize("shape", [(16, 320, 64, 64), (2, 1280, 8, 8), (3, 640, 5, 7)])
def test_add_bias_nhwc_matches_torch(cuda_device, shape):
    from photoverse_b200 import ops
    g = torch.Generator().manual_seed(shape[1])
    a = torch.randn(*shape, generator=g).to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(*shape, generator=g).to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(shape[1], generator=g).to(cuda_device)
    y = ops.add_bias_nhwc(a, b, bias)
    ref = a.float() + b.float() + bias[None, :, None, None]
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert bool(((y.float() - ref).abs() <= 4e-3 * ref.abs() + 1e-6).all())            # one bf16 rounding of the exact sum
    d = torch.randn(7, 33, 64, generator=g).to(cuda_device, torch.bfloat16)
    assert torch.equal(ops.add_bias_nhwc(d, d, bias[:64].contiguous()).float(),
                       (2 * d.float() + bias[:64]).to(torch.bfloat16).float())


@pytest.mark.parametrize("M,N", [(16 * 4096, 1280), (4096, 5120), (7, 16)])
def test_geglu_backward_matches_autograd(cuda_device, M, N):
    from photoverse_b200 import ops
    g = torch.Generator().manual_seed(N * 7 + M)
    h = (torch.randn(M, 2 * N, generator=g) * 1.5).to(cuda_device, torch.bfloat16)
    dy = torch.randn(M, N, generator=g).to(cuda_device, torch.bfloat16)
    dh = ops.geglu_bwd(h, dy)
    hr = h.float().requires_grad_(True)
    a, gate = hr.chunk(2, dim=-1)
    (a * F.gelu(gate)).backward(dy.float())
    assert dh.shape == h.shape and dh.dtype == torch.bfloat16
    assert bool(((dh.float() - hr.grad).abs() <= 2e-3 + 1e-2 * hr.grad.abs()).all())


def test_unet_with_fused_epilogues_matches_stock(cuda_device):
    """One UNet evaluation (channels-last bf16, random init, batch 2, latent 32) with the fused epilogues against the same
    model with the stock ops; and the fused kernels really run (native launch count)."""
    import photoverse_b200 as pv
    from photoverse_b200 import _lib
    from photoverse_b200.host.unet_sd15 import UNetSD15
    torch.manual_seed(3)
    unet = UNetSD15()
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(5,))
    unet.requires_grad_(False).eval().to(device=cuda_device, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(2, 4, 32, 32, device=cuda_device, dtype=torch.bfloat16)
    text = torch.randn(2, 77, 768, device=cuda_device, dtype=torch.bfloat16)
    img = torch.randn(2, 5, 768, device=cuda_device, dtype=torch.bfloat16)
    t = torch.tensor([500], device=cuda_device)
    with torch.no_grad():
        unet(x, t, (text, img))                     # first call packs the processors' weights: keep it out of the counts
        unet.set_fused_epilogues(False)
        n0 = _lib.launch_count()
        stock = unet(x, t, (text, img)).sample
        n_stock = _lib.launch_count() - n0
        unet.set_fused_epilogues(True)
        n0 = _lib.launch_count()
        fused = unet(x, t, (text, img)).sample
        n_fused = _lib.launch_count() - n0
    # 61 GroupNorms x 2 launches + 16 GEGLUs + 48 LayerNorms (32 of them with the residual sum) + 22 resnet residual sums
    assert n_fused == n_stock + 2 * 61 + 16 + 48 + 22, (n_stock, n_fused)
    cos = F.cosine_similarity(stock.double().flatten(1), fused.double().flatten(1), dim=1).min().item()
    rel = ((stock.float() - fused.float()).norm() / stock.float().norm()).item()
    print(f"UNet eval, fused vs stock epilogues: cosine {cos:.6f} relative L2 {rel:.4e}; native launches {n_stock} -> {n_fused}")
    assert cos >= 0.9995 and rel <= 3e-2
    # autograd on: a frozen norm runs the training pair (forward keeps the statistics), a trainable one the stock op
    from photoverse_b200.host.unet_sd15 import group_norm_act
    norm = unet.conv_norm_out
    h = torch.randn(2, 320, 8, 8, device=cuda_device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    n0 = _lib.launch_count()
    y_train = group_norm_act(norm, h, True)
    assert _lib.launch_count() == n0 + 2
    with torch.no_grad():
        y_inf = group_norm_act(norm, h, True)
    assert _lib.launch_count() == n0 + 4 and torch.equal(y_train, y_inf)
    norm.weight.requires_grad_(True)
    group_norm_act(norm, h, True)
    assert _lib.launch_count() == n0 + 4
    norm.weight.requires_grad_(False)
