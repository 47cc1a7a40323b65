"""GPU (B200): one data-parallel-ready training step (BASELINE config 4, reference train.py:495-549) on a reduced
SD-1.5-shaped UNet -- gradients of the whole trainable set through the CUDA path vs plain torch autograd through the
oracle processors / adapters with the same weights, inputs and fusion-rule RNG stream."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(device, dtype, seed=0, lora_dropout=0.0):
    import photoverse_b200 as pv
    from photoverse_b200.host.unet_sd15 import UNetSD15
    from photoverse_b200.lora import inject_lora
    torch.manual_seed(seed)
    unet = UNetSD15(block_out_channels=(320, 640), layers_per_block=1)
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(5,))
    ia, ta = pv.PhotoVerseAdapter(num_tokens=5), pv.PhotoVerseAdapter(num_tokens=5)
    unet.requires_grad_(False)
    inject_lora(unet, r=8, lora_dropout=lora_dropout)
    g = torch.Generator().manual_seed(seed + 1)
    for n, p in unet.named_parameters():
        if "to_k_ip" in n or "to_v_ip" in n:
            p.requires_grad_(True)
        if "lora_B" in n:                       # peft initialises B = 0: perturb so that dA is exercised
            with torch.no_grad():
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    for m in (unet, ia, ta):
        m.to(device=device, dtype=torch.float32)
    unet.to(dtype)                              # bf16 run: bf16 backbone (adapters keep fp32 masters)
    unet.eval()
    if lora_dropout > 0:
        pv.unet.set_cross_attention_layers_to_train(unet)      # train.py:462: activates the LoRA dropout
    return unet, ia, ta


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("concept_injection", [False, True], ids=["plain-text", "concept-injection"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_train_step_gradients_match_oracle_arm(cuda_device, dtype, concept_injection):
    """``concept_injection``: the text adapter's concept embeddings enter a (2-layer) CLIP text tower at the placeholder
    position (train.py:495-499, models/clip.py:17-24 -> pv_inject_concept_fwd/_bwd), so the text adapter's gradient comes
    from the denoising loss through the text branch of every processor; the oracle arm runs the same tower with the CPU
    restatement of the injection."""
    from oracle import adapter_oracle, clip_oracle
    from photoverse_b200.host import text_encoder as te_mod
    from oracle.host_reference import clone_with_oracle_processors
    from photoverse_b200.host.parallel import trainable_named_parameters
    from photoverse_b200.host.train_step import Trainer, synthetic_train_batch
    from photoverse_b200.unet import get_visual_cross_attention_values_norm
    unet, ia, ta = _build(cuda_device, dtype)
    ref_unet = clone_with_oracle_processors(unet).float()
    b = synthetic_train_batch(2, latent=16, seed=3, device=cuda_device, dtype=dtype)
    te = None
    if concept_injection:
        torch.manual_seed(11)
        te = te_mod.ConceptTextEncoder(layers=2).to(cuda_device).requires_grad_(False)
    # ---- product arm ----
    tr = Trainer(unet, ia, ta, text_encoder=te)
    torch.manual_seed(0)                        # fusion-rule RNG stream (one draw per attn2 layer): all 3 branches
    with torch.enable_grad():
        loss, parts = tr.loss(b)
    loss.backward()
    named = trainable_named_parameters(unet, ia, ta)
    got = {n: p.grad.detach().clone() for n, p in named if p.grad is not None}
    fusions = [p.last_fusion for p in unet.attn_processors.values() if hasattr(p, "last_fusion")]
    # ---- oracle arm: same weights, fp32 torch autograd ----
    ref_named = trainable_named_parameters(ref_unet, ia, ta)
    for _, p in ref_named:
        p.grad = None
    torch.manual_seed(0)
    f32 = lambda t: t.float()
    with torch.enable_grad():
        sd_t = {k: v for k, v in ta.named_parameters()}
        sd_i = {k: v for k, v in ia.named_parameters()}
        emb = [f32(e) for e in b.clip_hidden]
        concept = adapter_oracle.adapter_forward(emb, sd_t, None)
        img_tokens = adapter_oracle.adapter_forward(emb, sd_i, None)
        text = f32(b.text)
        if concept_injection:
            product_inject = te_mod.inject_concept_embeddings
            te_mod.inject_concept_embeddings = clip_oracle.inject_concept_embeddings      # the checker's gather
            try:
                text = te({"text_input_ids": b.text_ids, "concept_text_embeddings": concept,
                           "concept_placeholder_idx": b.placeholder_idx})[0]
            finally:
                te_mod.inject_concept_embeddings = product_inject
        pred = ref_unet(f32(b.noisy_latents), b.timesteps.float(), encoder_hidden_states=(text, img_tokens)).sample
        l_vis = get_visual_cross_attention_values_norm(ref_unet).mean()
        ref_loss = torch.nn.functional.mse_loss(pred, f32(b.noise)) + 0.01 * concept.abs().mean() + 0.001 * l_vis
    ref_loss.backward()
    assert len(fusions) == 4 and len(set(fusions)) == 3, fusions      # the seed exercises all three fusion branches
    assert abs(loss.item() - ref_loss.item()) <= (1e-4 if dtype == torch.float32 else 3e-2) * abs(ref_loss.item())
    # bf16 arm: bf16 backbone + bf16 tensor-core kernels against an fp32 autograd run; gradients of single parameters are
    # sums over a few thousand bf16-rounded terms
    tol = 2e-3 if dtype == torch.float32 else 1.0e-1
    checked = 0
    for (n, p_prod), (n_ref, p_ref) in zip(named, ref_named):
        assert n == n_ref
        r = p_ref.grad
        if r is None or r.abs().max() == 0:
            assert n not in got or got[n].abs().max().item() <= 1e-6, n
            continue
        e = _rel(got[n].float(), r.float())
        assert e <= tol, f"{n}: relative error {e:.3e} > {tol}"
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("lora_dropout", [0.0, 0.1], ids=["p0", "p0.1"])
def test_trainer_step_updates_only_the_trainable_set(cuda_device, lora_dropout):
    from photoverse_b200.host.train_step import Trainer, synthetic_train_batch
    unet, ia, ta = _build(cuda_device, torch.float32, seed=5, lora_dropout=lora_dropout)
    frozen_before = {n: p.detach().clone() for n, p in unet.named_parameters() if not p.requires_grad}
    train_before = {n: p.detach().clone() for n, p in unet.named_parameters() if p.requires_grad}
    tr = Trainer(unet, ia, ta, lr=1e-3)
    b = synthetic_train_batch(2, latent=16, seed=4, device=cuda_device, dtype=torch.float32)
    torch.manual_seed(7)
    l0, _ = tr.step(b)
    torch.manual_seed(7)
    l1, _ = tr.step(b)
    assert torch.isfinite(l0) and torch.isfinite(l1)
    for n, p in unet.named_parameters():
        if n in frozen_before:
            assert torch.equal(p, frozen_before[n]), n
    assert any(not torch.equal(p, train_before[n]) for n, p in unet.named_parameters() if n in train_before)
    assert tr.buf.numel() == sum(p.numel() for _, p in tr.named)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_fused_training_loss_matches_torch(cuda_device, dtype):
    """pv_train_loss_fwd / _bwd (train.py:509-535 in one reduction) vs the reference's torch expression, values and all
    three gradients; deterministic."""
    from photoverse_b200.loss import train_loss
    g = torch.Generator().manual_seed(5)
    pred = torch.randn(4, 4, 64, 64, generator=g).to(cuda_device, dtype).requires_grad_(True)
    noise = torch.randn(4, 4, 64, 64, generator=g).to(cuda_device, dtype)
    concept = torch.randn(4, 5, 768, generator=g).to(cuda_device, dtype).requires_grad_(True)
    vn = torch.rand(4, 16 * 8 * 5, generator=g).to(cuda_device, dtype).requires_grad_(True)
    loss, (l_mse, l_text, l_vis) = train_loss(pred, noise, concept, vn)
    (3.0 * loss).backward()
    pr, cr, vr = (t.detach().double().requires_grad_(True) for t in (pred, concept, vn))
    ref = torch.nn.functional.mse_loss(pr, noise.double()) + 0.01 * cr.abs().mean() + 0.001 * vr.mean()
    (3.0 * ref).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert abs(l_mse.item() - torch.nn.functional.mse_loss(pr, noise.double()).item()) <= 1e-5
    assert abs(l_text.item() - cr.abs().mean().item()) <= 1e-5 and abs(l_vis.item() - vr.mean().item()) <= 1e-5
    tol = 1e-6 if dtype == torch.float32 else 8e-3          # bf16: gradients are stored in bf16
    for got, want in ((pred.grad, pr.grad), (concept.grad, cr.grad), (vn.grad, vr.grad)):
        assert _rel(got.float(), want.float()) <= tol
    loss2, _ = train_loss(pred.detach(), noise, concept.detach(), vn.detach())
    assert torch.equal(loss2, loss.detach())
