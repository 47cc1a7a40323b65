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
        # gradients, then one "touched" flag per parameter (1.0 where this rank produced a gradient): the flags ride
        # along in the same allreduce, so that a parameter NO rank touched keeps ``.grad = None`` like in the reference's
        # single-process run (the optimizer then skips it: no weight decay / momentum-only update)
        n = self.offsets[-1]
        self._store = torch.zeros(n + len(self.params), device=dev, dtype=torch.float32)
        self.flat = self._store[:n]
        self.touched = self._store[n:]
        self.views = [self.flat[o:o + s].view(p.shape) for o, s, p in zip(self.offsets, self.sizes, self.params)]

    def numel(self) -> int:
        return self.flat.numel()

    def pack(self) -> None:
        """param.grad -> flat buffer; parameters without a gradient contribute zeros (and a 0 flag)."""
        flags = [0.0 if p.grad is None else 1.0 for p in self.params]
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        self._flags_host = flags
        self.touched.copy_(torch.tensor(flags, dtype=torch.float32), non_blocking=True)

    def allreduce_mean(self, group=None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        work = dist.all_reduce(self._store, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        self._flags_host = None                    # the flags now live in the reduced buffer
        if async_op:
            return work
        self.flat.div_(world)
        return None

    def unpack(self) -> None:
        """flat buffer -> param.grad.  A parameter that no rank produced a gradient for keeps ``.grad = None``; one that
        only other ranks touched gets the (mean) gradient allocated here."""
        flags = self._flags_host if getattr(self, "_flags_host", None) is not None else self.touched.tolist()
        for v, p, f in zip(self.views, self.params, flags):
            if f <= 0.0:
                p.grad = None
            elif p.grad is None:
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


class OverlappedGradReducer:
    """Gradient allreduce overlapped with the backward pass ("overlap the allreduce of processor / LoRA grads
    with the remaining UNet backward; adapter grads finish last").

    The flat buffer is cut into contiguous buckets, launched in the order the backward pass completes them --
    unet.up_blocks first, then mid / down blocks, the two adapters last (train.py:538: their gradients need the K/V-image
    gradients of all 16 layers).  A post-accumulate hook copies each gradient into its slice as soon as autograd has
    produced it; when the last EXPECTED gradient of the next bucket in line has arrived, its allreduce is launched
    asynchronously (NCCL runs it on its own stream, under the backward kernels still queued on the compute stream).
    Every rank issues the collectives in the same (bucket) order, whatever its own arrival times.

    Parameters the stochastic fusion rule leaves without a gradient this step (attention_processor.py:413-418; each rank
    draws its own) never fire their hook: the caller passes ``expected`` to :meth:`begin` -- known after the forward
    pass -- so that they do not hold their bucket back; their slices are zero and their "touched" flag is 0.
    ``finish()`` -- after ``backward()`` -- launches what is still pending plus the flags, waits and divides by the world
    size.  With one process nothing is hooked and ``finish()`` is ``pack()``."""

    ORDER = ("unet.up_blocks", "unet.mid_block", "unet.down_blocks", "unet.", "text_adapter.", "image_adapter.")

    def __init__(self, buf: FlatGradBuffer, group=None):
        self.buf, self.group = buf, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.buckets: List[List[int]] = []      # parameter indices, each bucket contiguous in the flat buffer
        self.launch_order: List[int] = []
        self.handles = []
        self.active = False
        if self.world == 1:
            return

        def key(name):
            for k, pre in enumerate(self.ORDER):
                if name.startswith(pre):
                    return k
            return len(self.ORDER)
        run_key, run, keys = None, [], []
        for i, n in enumerate(buf.names):
            k = key(n)
            if run and k != run_key:
                self.buckets.append(run)
                keys.append(run_key)
                run = []
            run_key = k
            run.append(i)
        if run:
            self.buckets.append(run)
            keys.append(run_key)
        self.launch_order = sorted(range(len(self.buckets)), key=lambda b: (keys[b], b))
        self.bucket_of = {i: b for b, idxs in enumerate(self.buckets) for i in idxs}
        for i, p in enumerate(buf.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(p):
            # an autograd Function that returns None for this parameter (a branch the fusion rule dropped) still
            # runs the accumulate node, with nothing to accumulate
            if not self.active or self.seen[i] or p.grad is None:
                return
            self.buf.views[i].copy_(p.grad)
            self.seen[i] = True
            if self.expected[i]:
                self.missing[self.bucket_of[i]] -= 1
                self._launch_ready()
        return hook

    def _launch_ready(self, force: bool = False):
        while self.next < len(self.launch_order):
            b = self.launch_order[self.next]
            if self.missing[b] > 0 and not force:
                return
            idxs = self.buckets[b]
            if force:                               # late gradients that were not expected, never-arrived ones -> zeros
                for i in idxs:
                    p = self.buf.params[i]
                    if not self.seen[i] and p.grad is not None:
                        self.buf.views[i].copy_(p.grad)
                        self.seen[i] = True
            lo, hi = self.buf.offsets[idxs[0]], self.buf.offsets[idxs[-1] + 1]
            self.handles.append(dist.all_reduce(self.buf.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.next += 1

    def begin(self, expected: Sequence[bool] = None):
        """Call after the forward pass, before ``backward()``.  ``expected[i]``: parameter i will receive a gradient."""
        if self.world == 1:
            return
        n = len(self.buf.params)
        self.expected = list(expected) if expected is not None else [True] * n
        self.seen = [False] * n
        self.missing = [sum(1 for i in idxs if self.expected[i]) for idxs in self.buckets]
        self.next = 0
        self.handles = []
        self.early = 0
        self.buf.flat.zero_()
        self.active = True

    def finish(self) -> int:
        """Call after ``backward()``: returns how many buckets had been launched before backward() returned."""
        buf = self.buf
        if self.world == 1:
            buf.pack()
            return 0
        self.active = False
        early = self.next
        self._launch_ready(force=True)
        buf.touched.copy_(torch.tensor([1.0 if s else 0.0 for s in self.seen], dtype=torch.float32), non_blocking=True)
        self.handles.append(dist.all_reduce(buf.touched, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for h in self.handles:
            h.wait()
        buf._flags_host = None
        buf.flat.div_(self.world)
        return early
