"""CPU, world_size 2, gloo: the multi-GPU host logic -- generation sharding with rank-count-invariant
per-sample seeds, and the flat-buffer gradient allreduce with missing gradients contributing zeros."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from photoverse_b200.host.parallel import (FlatGradBuffer, OverlappedGradReducer, sample_seeds, shard_range,
                                           trainable_named_parameters)


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
        # plain lists: a tensor in an mp.Queue travels as a shared-memory handle that dies with the worker
        q.put((rank, lin_a.weight.grad.tolist(), lin_a.bias.grad.tolist(), lin_b.weight.grad.tolist(),
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
    res = [(r, torch.tensor(gw), torch.tensor(gb), torch.tensor(gu), norms) for r, gw, gb, gu, norms in res]
    for rank, gw, gb, gu, norms in res:
        assert torch.allclose(gw, a_w / (n_img + 1e-6), atol=1e-6)
        assert torch.allclose(gb, a_b / (n_img + 1e-6), atol=1e-6)
        assert torch.allclose(gu, b_w / (n_unet + 1e-6), atol=1e-6)
        assert abs(norms["image_adapter."] - float(n_img)) < 1e-5 and abs(norms["unet."] - float(n_unet)) < 1e-5
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][3], res[1][3])      # replicas stay identical


def _worker_overlap(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # identical replicas
        up, down, ad = torch.nn.Linear(5, 4), torch.nn.Linear(4, 5), torch.nn.Linear(3, 4)
        dead = torch.nn.Linear(2, 2, bias=False)   # never used: no rank produces a gradient
        named = [("image_adapter.l.weight", ad.weight), ("image_adapter.l.bias", ad.bias),
                 ("unet.down_blocks.0.w", down.weight), ("unet.down_blocks.0.b", down.bias),
                 ("unet.mid_block.dead", dead.weight),
                 ("unet.up_blocks.0.w", up.weight), ("unet.up_blocks.0.b", up.bias)]
        buf = FlatGradBuffer(named)
        red = OverlappedGradReducer(buf)
        assert [len(b) for b in red.buckets] == [2, 2, 1, 2]
        x = torch.full((2, 3), float(rank + 1))
        h = down(ad(x))
        y = up(h) if rank == 0 else h.sum(1, keepdim=True).expand(-1, 4)      # rank 1: the up block gets NO gradient
        # rank 1 knows after its "forward" that the up block is not in its graph; nobody expects the dead parameter
        expected = [True, True, True, True, False, rank == 0, rank == 0]
        red.begin(expected)
        y.sum().backward()
        early = red.finish()
        buf.unpack()
        q.put((rank, early, [None if p.grad is None else p.grad.tolist() for _, p in named]))
    finally:
        dist.destroy_process_group()


def test_overlapped_reducer_matches_mean_of_rank_gradients_world2_gloo():
    """Bucketed allreduce launched from autograd hooks: result == mean of the per-rank gradients, a parameter only one
    rank touched gets half of that rank's gradient, a parameter nobody touched keeps grad None; buckets whose gradients
    all arrived are launched before backward() returns."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: the same two losses, gradients averaged by hand
    torch.manual_seed(0)
    up, down, ad = torch.nn.Linear(5, 4), torch.nn.Linear(4, 5), torch.nn.Linear(3, 4)
    params = [ad.weight, ad.bias, down.weight, down.bias, None, up.weight, up.bias]
    grads = []
    for rank in range(2):
        for p_ in (up, down, ad):
            p_.zero_grad()
        x = torch.full((2, 3), float(rank + 1))
        h = down(ad(x))
        y = up(h) if rank == 0 else h.sum(1, keepdim=True).expand(-1, 4)
        y.sum().backward()
        grads.append([None if (p_ is None or p_.grad is None) else p_.grad.clone() for p_ in params])
    res = [(r, e, [None if g is None else torch.tensor(g) for g in got]) for r, e, got in res]
    for rank, early, got in res:
        for k, g in enumerate(got):
            parts = [gr[k] for gr in grads if gr[k] is not None]
            if not parts:
                assert g is None, f"parameter {k}: nobody produced a gradient, grad must stay None"
            else:
                assert torch.allclose(g, sum(parts) / 2, atol=1e-6), f"parameter {k} on rank {rank}"
    assert res[0][1] >= 3 and res[1][1] >= 3    # up, mid (nothing expected) and down buckets went out during backward
    assert all(torch.equal(a, b) if a is not None else b is None for a, b in zip(res[0][2], res[1][2]))


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
