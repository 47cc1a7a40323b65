"""Concept-token injection (counterpart of the reference's ``models/clip.py:_inject_concept_embeddings``, :17-24).

The reference monkey-patches ``CLIPTextTransformer.forward`` (transformers 4.40 internals) so that the text adapter's
concept embeddings replace the placeholder token of the prompt *before* the position embeddings and the transformer
layers (:50-63).  The arithmetic of that patch is this one gather/scatter; it runs in the CUDA library (forward and
backward -- in training the gradient reaches the text adapter through it, train.py:495-499).  Wiring it into a concrete
``CLIPTextModel`` needs the real encoder weights and is left to the caller: compute ``token_embedding(ids)``, call
:func:`inject_concept_embeddings`, continue with the encoder's embedding + layer stack.
"""
import ctypes

import torch

from . import _lib
from .ops import _dt, _ptr, _stream


class _InjectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs_embeds, concept, idx):
        B, L, D = inputs_embeds.shape
        T = concept.shape[1]
        out = torch.empty_like(inputs_embeds)
        _lib.check(_lib.lib().pv_inject_concept_fwd(_dt(inputs_embeds), _ptr(inputs_embeds), _ptr(concept), _ptr(idx), _ptr(out),
                                                     B, L, T, D, _stream()), "pv_inject_concept_fwd")
        ctx.save_for_backward(idx)
        ctx.geom = (B, L, T, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        B, L, T, D = ctx.geom
        dout = dout.contiguous()
        din = torch.empty(B, L, D, device=dout.device, dtype=dout.dtype)
        dconcept = torch.empty(B, T, D, device=dout.device, dtype=dout.dtype)
        _lib.check(_lib.lib().pv_inject_concept_bwd(_dt(dout), _ptr(dout), _ptr(idx), _ptr(din), _ptr(dconcept), B, L, T, D,
                                                     _stream()), "pv_inject_concept_bwd")
        return din, dconcept, None


def inject_concept_embeddings(inputs_embeds: torch.Tensor, concept_text_embeddings: torch.Tensor,
                              concept_placeholder_idx) -> torch.Tensor:
    """``inputs_embeds`` [B,L,D], ``concept_text_embeddings`` [B,T,D] (same dtype, bf16 or fp32, CUDA);
    ``concept_placeholder_idx``: B integers (list / tensor), position of the placeholder token of each prompt
    (datasets/utils.py:215-220: word index + 1).  Returns the new [B,L,D] embeddings."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError("photoverse_b200 runs on CUDA only (no CPU fallback)")
    B, L, D = inputs_embeds.shape
    T = concept_text_embeddings.shape[1]
    idx = torch.as_tensor(concept_placeholder_idx, dtype=torch.int32)
    if idx.numel() != B or int(idx.min()) < 0 or int(idx.max()) + T > L:
        raise ValueError(f"placeholder indices must be {B} values with 0 <= idx and idx + {T} <= {L}")
    idx = idx.to(inputs_embeds.device)
    return _InjectFn.apply(inputs_embeds.contiguous(), concept_text_embeddings.to(inputs_embeds.dtype).contiguous(), idx)
