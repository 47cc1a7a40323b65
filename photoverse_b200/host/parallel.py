"""One-process-per-GPU data parallelism for the two PhotoVerse workloads.

Generation: samples (and their CFG branches) are independent through the adapters, the processors and the whole
denoise loop (models/infer.py:98-119 has no cross-sample op) -> the global batch is split contiguously across ranks,
weights are replicated, per-sample seeds make results independent of the rank count, and there is NO collective in
the loop.

Training: data parallel; one allreduce (sum, then / world) per step over ONE flat fp32 buffer holding the gradients of
the trainable set {image_adapter, text_adapter, to_k_ip / to_v_ip, LoRA A/B} (train.py:348-370, 412-419).  A parameter
that received no gradient this step (the stochastic fusion rule drops a branch in ~1/3 of the layers,
attention_processor.py:413-420) contributes zeros -- the situation stock DDP would reject.
The reference's three per-module ``clip_grad_norm_`` calls (train.py:541-544) are applied to the reduced gradients.
"""
from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the global batch owned by ``rank`` (remainder spread over the first ranks)."""
    if not (0 <= rank < world) or global_batch < 0:
        raise ValueError(f"bad shard query batch={global_batch} world={world} rank={rank}")
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def sample_seeds(base_seed: int, global_batch: int, world: int, rank: int) -> List[int]:
    """Per-sample RNG seeds of this rank's shard: a sample's noise depends on its GLOBAL index only, so the generated
    latents are identical for any number of ranks."""
    b, e = shard_range(global_batch, world, rank)
    return [base_seed + i for i in range(b, e)]


def trainable_named_parameters(unet, image_adapter, text_adapter) -> List[Tuple[str, torch.nn.Parameter]]:
    """The trainable set in a deterministic order (identical on every rank): adapters, then the unet's requires_grad
    parameters (to_k_ip / to_v_ip and / or LoRA factors, depending on how the caller froze the model)."""
    out = [(f"image_adapter.{n}", p) for n, p in image_adapter.named_parameters() if p.requires_grad]
    out += [(f"text_adapter.{n}", p) for n, p in text_adapter.named_parameters() if p.requires_grad]
    out += [(f"unet.{n}", p) for n, p in unet.named_parameters() if p.requires_grad]
    return out


class FlatGradBuffer:
    """Gradients of a fixed parameter list packed into one contiguous fp32 buffer for a single allreduce."""

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]], device=None):
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + s)
        dev = device if device is not None else (self.params[0].device if self.params else "cpu")
        self.flat = torch.zeros(self.offsets[-1], device=dev, dtype=torch.float32)
        self.views = [self.flat[o:o + s].view(p.shape) for o, s, p in zip(self.offsets, self.sizes, self.params)]

    def numel(self) -> int:
        return self.flat.numel()

    def pack(self) -> None:
        """param.grad -> flat buffer; parameters without a gradient contribute zeros."""
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def allreduce_mean(self, group=None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(world)
        return None

    def unpack(self) -> None:
        """flat buffer -> param.grad (allocating a gradient for parameters that had none)."""
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.to(p.dtype).clone()
            else:
                p.grad.copy_(v)

    def group_norms(self, prefixes: Iterable[str]) -> Dict[str, torch.Tensor]:
        """L2 norm of the (reduced) gradient of every parameter group selected by name prefix."""
        out = {}
        for pre in prefixes:
            sq = [v.pow(2).sum() for n, v in zip(self.names, self.views) if n.startswith(pre)]
            out[pre] = torch.stack(sq).sum().sqrt() if sq else torch.zeros((), device=self.flat.device)
        return out

    def clip_groups_(self, prefixes: Iterable[str], max_norm: float = 1.0) -> Dict[str, torch.Tensor]:
        """Per-group ``clip_grad_norm_(params, max_norm)`` (train.py:541-544) on the flat buffer, no host sync."""
        norms = self.group_norms(prefixes)
        for pre, nrm in norms.items():
            coef = torch.clamp(max_norm / (nrm + 1e-6), max=1.0)
            for n, v in zip(self.names, self.views):
                if n.startswith(pre):
                    v.mul_(coef)
        return norms
