"""Losses with the reference's names and semantics (reference Utils/Utils.py:45-98) on the fused
CUDA kernels (dmm_bpr_fwd_bwd, dmm_infonce_fwd/bwd).  Star-imported by Model.py and Main.py like in
the reference; only the three functions the hot path uses are exported."""
from __future__ import annotations

import torch
from torch import Tensor

__all__ = ["l2_reg_loss", "InfoNCE", "bpr_loss"]


def l2_reg_loss(reg: float, embeddings: list, device=None) -> Tensor:
    """reg * sum_e ||e||^2 (Utils/Utils.py:45-54)."""
    emb_loss = None
    for emb in embeddings:
        s = torch.sum(emb ** 2)
        emb_loss = s if emb_loss is None else emb_loss + s
    if emb_loss is None:
        emb_loss = torch.tensor(0., device=device)
    return emb_loss * reg


def InfoNCE(batch_view1: Tensor, batch_view2: Tensor, idx: Tensor, temperature: float, b_cos: bool = True):
    """Average in-batch InfoNCE over gathered rows ``idx`` (Utils/Utils.py:57-75).  The B x B logits
    never reach HBM; b_cos=False (never used by the reference) is not supported by the kernel."""
    from ..autograd import InfoNCEFn
    if batch_view1.shape[1:] != batch_view2.shape[1:]:
        raise ValueError(
            f"InfoNCE expected the same shape for two views. But got view1.shape={batch_view1[idx].shape} "
            f"and view2.shape={batch_view2[idx].shape}.")
    if not b_cos:
        raise NotImplementedError("InfoNCE(b_cos=False) is not on the reference's hot path")
    return InfoNCEFn.apply(batch_view1, batch_view2, idx.to(batch_view1.device).long(), float(temperature))


def bpr_loss(user_emb: Tensor, pos_item_emb: Tensor, neg_item_embs: Tensor):
    """mean(-log(10e-6 + sigmoid(u.p - u.n))) (Utils/Utils.py:78-98)."""
    from ..autograd import BPRFn
    if not (user_emb.shape == pos_item_emb.shape == neg_item_embs.shape):
        raise ValueError("bpr_loss expects three [batch, dim] tensors of the same shape")
    return BPRFn.apply(user_emb, pos_item_emb, neg_item_embs)
