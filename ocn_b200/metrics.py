"""Ranking metrics of the drivers on the device: the ogb ``Evaluator.eval`` calls of
NeighborOverlap_large.py:162-179 (Hits@K) and NeighborOverlapCitation2.py:256-259 (MRR) without moving
the scores to the host."""
from __future__ import annotations

from typing import Dict, Iterable

import torch
from torch import Tensor

from . import _lib
from .cn import _stream
from .graph import _require_cuda


def mrr_list(y_pred_pos: Tensor, y_pred_neg: Tensor) -> Tensor:
    """``evaluator.eval({'y_pred_pos': [B], 'y_pred_neg': [B, K]})['mrr_list']`` (ogbl-citation2)."""
    _require_cuda(y_pred_pos)
    pos = y_pred_pos.reshape(-1).float().contiguous()
    neg = y_pred_neg.float().contiguous().view(pos.numel(), -1)
    out = torch.empty_like(pos)
    with torch.cuda.device(pos.device):
        _lib.check(_lib.lib().ocn_mrr(_lib.ptr(pos), _lib.ptr(neg), pos.numel(), neg.shape[1], _lib.ptr(out),
                                      _stream(pos.device)), "ocn_mrr")
    return out


def hits_at_k(y_pred_pos: Tensor, y_pred_neg: Tensor, k: int) -> Tensor:
    """``evaluator.eval(...)[f'hits@{K}']`` as a 0-dim device tensor (no host sync)."""
    _require_cuda(y_pred_pos)
    pos = y_pred_pos.reshape(-1).float().contiguous()
    neg = y_pred_neg.reshape(-1).float().contiguous()
    L = _lib.lib()
    nb = L.ocn_hits_bytes(neg.numel())
    scratch = torch.empty(nb, dtype=torch.uint8, device=pos.device)
    out = torch.empty(1, dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        _lib.check(L.ocn_hits_at_k(_lib.ptr(pos), pos.numel(), _lib.ptr(neg), neg.numel(), int(k), _lib.ptr(scratch),
                                   nb, _lib.ptr(out), _stream(pos.device)), "ocn_hits_at_k")
    return out[0]


def evaluate_hits(pos_train: Tensor, pos_valid: Tensor, neg_valid: Tensor, pos_test: Tensor, neg_test: Tensor,
                  ks: Iterable[int] = (1, 3, 10, 20, 50, 100)) -> Dict[str, tuple]:
    """The result dictionary of ``test()`` (NeighborOverlap_large.py:160-179): train hits use the validation
    negatives.  One host read at the end."""
    res = {}
    for K in ks:
        res[f"Hits@{K}"] = torch.stack((hits_at_k(pos_train, neg_valid, K), hits_at_k(pos_valid, neg_valid, K),
                                        hits_at_k(pos_test, neg_test, K)))
    host = torch.stack(list(res.values())).cpu()
    return {k: tuple(float(v) for v in host[i]) for i, k in enumerate(res)}
