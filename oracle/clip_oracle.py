"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's concept-token injection.

Follows /root/reference/models/clip.py:17-24 (``_inject_concept_embeddings``):
    :18     new = inputs_embeds.clone()
    :19     T = concept_text_embeddings.shape[1]
    :21     leftover = L - T - idx
    :22     new[b, idx+T:] = inputs_embeds[b, idx+1 : idx+1+leftover]     (tail shifted right by T-1, truncated at L;
                                                                           the placeholder token inputs_embeds[b, idx] is dropped)
    :23     new[b, idx:idx+T] = concept_text_embeddings[b]
Pinned against the verbatim function (oracle/ref_loader.load_reference_inject_fn) in tests/test_oracle_golden.py and
through tests/golden/inject_concept.npz.  Differentiable torch, so autograd through it checks the backward kernel.
"""
import torch


def inject_concept_embeddings(inputs_embeds: torch.Tensor, concept: torch.Tensor, idx) -> torch.Tensor:
    B, L, _ = inputs_embeds.shape
    T = concept.shape[1]
    rows = []
    for b in range(B):
        i = int(idx[b])
        leftover = L - T - i
        rows.append(torch.cat([inputs_embeds[b, :i], concept[b], inputs_embeds[b, i + 1:i + 1 + leftover]], dim=0))
    return torch.stack(rows)


def inject_case(seed: int = 5, B: int = 3, L: int = 77, T: int = 5, D: int = 768, dtype=torch.float32):
    from . import detgen
    x = detgen.unit_variance((B, L, D), seed * 10 + 1, dtype)
    c = detgen.unit_variance((B, T, D), seed * 10 + 2, dtype)
    idx = [5, 1, L - T][:B] + [7] * max(0, B - 3)        # includes the two edges: early placeholder, concept ends at L
    return x, c, idx
