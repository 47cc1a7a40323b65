"""The training objective of the path as one CUDA reduction (counterpart of reference ``train.py:509-535``):

    loss = MSE(noise_pred, noise) + 0.01 * mean|concept_text_embeddings| + 0.001 * mean(||V_ip||)

``v_ip_norms`` is what ``get_visual_cross_attention_values_norm(unet)`` returns (models/unet.py:38-47: the stacked
``to_v_ip_norm`` side outputs of the 16 attn2 processors).  Forward = ``pv_train_loss_fwd`` (two launches, fixed summation
order), backward = ``pv_train_loss_bwd`` (one launch).  No PyTorch fallback: CPU tensors raise.
"""
import torch

from . import _lib
from .ops import _dt, _ptr, _stream, _workspace


class _TrainLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, concept, vnorm, w_text, w_vis):
        out4 = torch.empty(4, device=pred.device, dtype=torch.float32)
        lib = _lib.lib()
        ws = _workspace(int(lib.pv_train_loss_ws_bytes()), pred.device)
        _lib.check(lib.pv_train_loss_fwd(_dt(pred), _ptr(pred), _ptr(target), pred.numel(), _ptr(concept), concept.numel(),
                                         _ptr(vnorm), vnorm.numel(), float(w_text), float(w_vis), _ptr(out4), _ptr(ws), _stream()),
                   "pv_train_loss_fwd")
        ctx.save_for_backward(pred, target, concept)
        ctx.k, ctx.w, ctx.vshape = vnorm.numel(), (float(w_text), float(w_vis)), vnorm.shape
        ctx.mark_non_differentiable(out4)
        return out4[0], out4

    @staticmethod
    def backward(ctx, gloss, _gparts):
        pred, target, concept = ctx.saved_tensors
        g = gloss.detach().to(torch.float32).reshape(1).contiguous()
        d_pred, d_concept = torch.empty_like(pred), torch.empty_like(concept)
        d_vnorm = torch.empty(ctx.vshape, device=pred.device, dtype=pred.dtype)
        _lib.check(_lib.lib().pv_train_loss_bwd(_dt(pred), _ptr(pred), _ptr(target), pred.numel(), _ptr(concept), concept.numel(),
                                                ctx.k, ctx.w[0], ctx.w[1], _ptr(g), _ptr(d_pred), _ptr(d_concept), _ptr(d_vnorm),
                                                _stream()), "pv_train_loss_bwd")
        return d_pred, None, d_concept, d_vnorm, None, None


def train_loss(noise_pred: torch.Tensor, noise: torch.Tensor, concept_text_embeddings: torch.Tensor, v_ip_norms: torch.Tensor,
               w_text: float = 0.01, w_vis: float = 0.001):
    """Returns (loss, (l_mse, l_text, l_vis)) -- scalars on the device, fp32.  All inputs share one dtype (bf16 / fp32)."""
    if not noise_pred.is_cuda:
        raise RuntimeError("photoverse_b200 runs on CUDA only (no CPU fallback)")
    dt = noise_pred.dtype
    args = [t.to(dt).contiguous() for t in (noise_pred, noise, concept_text_embeddings, v_ip_norms)]
    loss, parts = _TrainLossFn.apply(*args, w_text, w_vis)
    return loss, (parts[1], parts[2], parts[3])
