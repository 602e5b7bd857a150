"""Explicit-matrix variant of the predictor bodies: the reference's own data flow for the _large
drivers, where ``cn1 = adjoverlap(adj, adj, e)`` and ``cn2 = adjoverlap(adj, adj2, e)`` arrive as sparse
[B x N] matrices and ``adj2`` may be the folded adj2byblock matrix (NeighborOverlap_large.py:68-82).

The intersections, A^2 and the SpMMs run in libocn_b200 (ocn_rows_intersect_*, ocn_spgemm_a2_*, ocn_spmm_csr);
the batch normalisation / orthogonalisation between them is restated with torch ops ON THE GPU in the
reference's own order (model.py:2261-2423, 3114-3126) -- it is glue over O(nnz(cn)) entries, not a hot kernel.
Nothing here runs on the CPU.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .cn import SparseRows
from .sparse_ops import spmm_add


def _col_sum(sp: SparseRows) -> Tensor:
    return torch.zeros(sp.shape[1], dtype=torch.float32, device=sp.col.device).index_add_(0, sp.col, sp.value)


def _normalise_cn1(cn1: SparseRows, fill: float) -> SparseRows:
    col_sum = _col_sum(cn1)
    col_sum[col_sum == 0] = 1
    inv = 1 / col_sum
    inv[col_sum == 1] = fill
    return SparseRows(cn1.rowptr, cn1.col, cn1.value * inv[cn1.col], cn1.shape)


def _hadamard_sum(a: SparseRows, b: SparseRows) -> Tensor:
    """sum of a*b over the common pattern (innerprod1, model.py:2241-2250)."""
    n = a.shape[1]
    ka, kb = a.row() * n + a.col, b.row() * n + b.col
    if ka.numel() == 0 or kb.numel() == 0:
        return torch.zeros((), device=a.col.device)
    idx = torch.searchsorted(kb, ka).clamp_(max=kb.numel() - 1)
    hit = kb[idx] == ka
    return (a.value[hit] * b.value[idx[hit]]).sum()


def _orthogonalise(cn2: SparseRows, base: SparseRows, coeff: Tensor) -> SparseRows:
    n, B = cn2.shape[1], cn2.shape[0]
    k2, k1 = cn2.row() * n + cn2.col, base.row() * n + base.col
    uk, inv = torch.unique(torch.cat((k2, k1)), return_inverse=True)
    v2 = torch.zeros(uk.numel(), dtype=torch.float32, device=uk.device)
    v1 = torch.zeros_like(v2)
    v2[inv[:k2.numel()]] = cn2.value
    v1[inv[k2.numel():]] = base.value
    new = v2 - coeff * v1
    row, col = torch.div(uk, n, rounding_mode="floor"), uk % n
    col_sum = torch.zeros(n, dtype=torch.float32, device=uk.device).index_add_(0, col, new)
    col_sum[col_sum == 0] = 1
    rowptr = torch.zeros(B + 1, dtype=torch.int64, device=uk.device)
    torch.cumsum(torch.bincount(row, minlength=B), 0, out=rowptr[1:])
    return SparseRows(rowptr, col, new * (1 / col_sum)[col], cn2.shape)


def cn5_explicit(pred, x: Tensor, cn1: SparseRows, cn2: SparseRows, tar_ei: Tensor):
    """model.py:2261-2429 on explicit matrices; updates ``pred.innerprod`` / ``pred.n`` in training mode."""
    ncn1 = _normalise_cn1(cn1, 0.0)
    if pred.training:
        pred._running_mean_update(_hadamard_sum(cn2, ncn1).detach())
    ip = pred.innerprod.detach().float()
    if cn1.col.numel() + cn2.col.numel() > 0:
        scale = ncn1.value.abs().max() if ncn1.value.numel() else torch.zeros((), device=x.device)
    else:
        scale = torch.ones((), device=x.device)
    coeff = torch.where(scale > 0, ip / torch.where(scale > 0, scale, torch.ones_like(scale)), ip)
    ncn2 = _orthogonalise(cn2, ncn1, coeff)
    return spmm_add(ncn1, x), spmm_add(ncn2, x), x[tar_ei[0]] * x[tar_ei[1]]


def cn7_explicit(pred, x: Tensor, cn1: SparseRows, cn2: SparseRows, tar_ei: Tensor, fill: float):
    """model.py:3114-3216 on explicit matrices (identity polynomial, raw cn2: SURVEY Q4)."""
    return spmm_add(_normalise_cn1(cn1, float(fill)), x), spmm_add(cn2, x), x[tar_ei[0]] * x[tar_ei[1]]
