"""SD-1.5 CLIP text tower as a plain-PyTorch host model, with the reference's concept-token injection wired in
(counterpart of ``models/clip.py:29-113`` ``clip_text_transformer_forward``).

transformers' ``CLIPTextModel`` cannot be patched the way the reference does it (the patch needs 4.40 internals; 5.5 is
installed), and the tower is *outside* the hot path anyway: it only produces ``encoder_hidden_states`` [B, 77, 768].  So,
like ``host/unet_sd15.py`` for the UNet, this is host plumbing around the path: a random-initialisable CLIP ViT-L/14
text encoder (hidden 768, 12 layers, 12 heads, MLP 3072, quick-GELU, causal mask, 49 408 tokens, 77 positions) whose
parameter names are those of ``transformers.CLIPTextModel`` (``text_model.embeddings.token_embedding.weight`` ...), so a
real SD-1.5 text-encoder state dict loads with ``load_state_dict``.

The one piece of the reference patch that is *on* the path -- ``_inject_concept_embeddings`` (:17-24): the text adapter's
concept embeddings replace the placeholder token between the token embedding and the position embedding (:50-55) -- is
the CUDA gather/scatter ``pv_inject_concept_fwd/_bwd`` (photoverse_b200.clip).  In training its backward is what carries
the gradient of the denoising loss into the text adapter (train.py:495-499).
"""
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from ..clip import inject_concept_embeddings


class _SelfAttn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.q_proj, self.k_proj = nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.v_proj, self.out_proj = nn.Linear(dim, dim), nn.Linear(dim, dim)

    def forward(self, x):
        B, L, D = x.shape
        split = lambda t: t.view(B, L, self.heads, D // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(split(self.q_proj(x)), split(self.k_proj(x)), split(self.v_proj(x)), is_causal=True)
        return self.out_proj(o.transpose(1, 2).reshape(B, L, D))


class _MLP(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        h = self.fc1(x)
        return self.fc2(h * torch.sigmoid(1.702 * h))            # quick_gelu (CLIP)


class _Layer(nn.Module):
    def __init__(self, dim, heads, hidden):
        super().__init__()
        self.self_attn = _SelfAttn(dim, heads)
        self.layer_norm1, self.layer_norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.mlp = _MLP(dim, hidden)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, vocab, positions, dim):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, dim)
        self.position_embedding = nn.Embedding(positions, dim)


class _Encoder(nn.Module):
    def __init__(self, layers, dim, heads, hidden):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(dim, heads, hidden) for _ in range(layers)])


class _TextTransformer(nn.Module):
    def __init__(self, vocab, positions, dim, layers, heads, hidden):
        super().__init__()
        self.embeddings = _Embeddings(vocab, positions, dim)
        self.encoder = _Encoder(layers, dim, heads, hidden)
        self.final_layer_norm = nn.LayerNorm(dim)


class ConceptTextEncoder(nn.Module):
    """``text_encoder(inputs)[0]`` of the reference (infer.py:93-96, train.py:497-499): ``inputs`` is the reference's
    dict ``{"text_input_ids": [B, 77] int64, "concept_text_embeddings": [B, T, 768] | absent,
    "concept_placeholder_idx": B ints | absent}``; returns ``(last_hidden_state [B, 77, 768],)``."""

    def __init__(self, vocab_size=49408, max_positions=77, hidden=768, layers=12, heads=12, mlp=3072):
        super().__init__()
        self.text_model = _TextTransformer(vocab_size, max_positions, hidden, layers, heads, mlp)

    def forward(self, inputs):
        if not isinstance(inputs, dict) or "text_input_ids" not in inputs:
            raise ValueError("You have to specify either input_ids")            # the reference's message (clip.py:48)
        ids = inputs["text_input_ids"]
        concept: Optional[torch.Tensor] = inputs.get("concept_text_embeddings")
        idx = inputs.get("concept_placeholder_idx")
        tm = self.text_model
        ids = ids.view(-1, ids.shape[-1])
        emb = tm.embeddings.token_embedding(ids)                                 # clip.py:56
        if concept is not None:
            emb = inject_concept_embeddings(emb, concept, idx)                   # clip.py:57-58 -> CUDA kernel
        pos = tm.embeddings.position_embedding.weight[: ids.shape[1]]
        h = emb + pos                                                            # clip.py:62 (CLIPTextEmbeddings)
        for layer in tm.encoder.layers:
            h = layer(h)
        return (tm.final_layer_norm(h),)
