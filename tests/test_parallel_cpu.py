"""CPU, world_size 2, gloo: the multi-GPU host logic -- generation sharding with rank-count-invariant
per-sample seeds, and the flat-buffer gradient allreduce with missing gradients contributing zeros."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from photoverse_b200.host.parallel import FlatGradBuffer, sample_seeds, shard_range, trainable_named_parameters


def test_shard_range_covers_batch_exactly():
    for gb in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(gb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_sample_seeds_do_not_depend_on_world_size():
    ref = sample_seeds(1234, 64, 1, 0)
    for world in (2, 4, 8):
        got = sum((sample_seeds(1234, 64, world, r) for r in range(world)), [])
        assert got == ref


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin_a, lin_b = torch.nn.Linear(6, 4), torch.nn.Linear(4, 3, bias=False)
        named = [("image_adapter.a.weight", lin_a.weight), ("image_adapter.a.bias", lin_a.bias), ("unet.b.weight", lin_b.weight)]
        buf = FlatGradBuffer(named)
        # rank-dependent gradients; rank 1 has NO gradient for unet.b.weight (dropped fusion branch)
        lin_a.weight.grad = torch.full_like(lin_a.weight, float(rank + 1))
        lin_a.bias.grad = torch.arange(4, dtype=torch.float32) * (rank + 1)
        if rank == 0:
            lin_b.weight.grad = torch.full_like(lin_b.weight, 4.0)
        buf.pack()
        buf.allreduce_mean()
        norms = buf.clip_groups_(("image_adapter.", "unet."), max_norm=1.0)
        buf.unpack()
        q.put((rank, lin_a.weight.grad.clone(), lin_a.bias.grad.clone(), lin_b.weight.grad.clone(),
               {k: float(v) for k, v in norms.items()}))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_mean_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: mean over ranks, missing gradient == zeros, then per-group clipping to norm 1
    a_w = torch.full((4, 6), 1.5)
    a_b = torch.arange(4, dtype=torch.float32) * 1.5
    b_w = torch.full((3, 4), 2.0)
    n_img = torch.sqrt(a_w.pow(2).sum() + a_b.pow(2).sum())
    n_unet = b_w.norm()
    for rank, gw, gb, gu, norms in res:
        assert torch.allclose(gw, a_w / (n_img + 1e-6), atol=1e-6)
        assert torch.allclose(gb, a_b / (n_img + 1e-6), atol=1e-6)
        assert torch.allclose(gu, b_w / (n_unet + 1e-6), atol=1e-6)
        assert abs(norms["image_adapter."] - float(n_img)) < 1e-5 and abs(norms["unet."] - float(n_unet)) < 1e-5
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][3], res[1][3])      # replicas stay identical


def test_trainable_set_matches_survey_counts():
    """adapters 2 x 28 904 960 + to_k_ip/to_v_ip 19 169 280 + LoRA r=8 595 968."""
    import photoverse_b200 as pv
    from photoverse_b200.host.unet_sd15 import UNetSD15
    from photoverse_b200.lora import inject_lora
    unet = UNetSD15()
    pv.set_visual_cross_attention_adapter(unet, num_tokens=(5,))
    ia, ta = pv.PhotoVerseAdapter(num_tokens=5), pv.PhotoVerseAdapter(num_tokens=5)
    unet.requires_grad_(False)
    inject_lora(unet, r=8)
    for n, p in unet.named_parameters():
        if "to_k_ip" in n or "to_v_ip" in n:
            p.requires_grad_(True)                      # explicit trainable set
    named = trainable_named_parameters(unet, ia, ta)
    total = sum(p.numel() for _, p in named)
    assert total == 2 * 28_904_960 + 19_169_280 + 595_968
    assert [n for n, _ in named] == [n for n, _ in trainable_named_parameters(unet, ia, ta)]
