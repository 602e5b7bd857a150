"""``torch_sparse.matmul``: ``spmm_add / spmm_mean / spmm_max`` (model.py:6, 45-53, 2426-2427) on ``ocn_spmm_csr``
(warp per row, float4 gathers) with the scatter backward ``ocn_spmm_csr_bwd`` / ``ocn_spmm_csr_max_bwd``."""
from __future__ import annotations

from torch import Tensor

from ...sparse_ops import _SpmmFn, _REDUCE
from .tensor import SparseTensor, _need_cuda


def _spmm(src: SparseTensor, other: Tensor, reduce: str) -> Tensor:
    _need_cuda(src._col, f"spmm_{reduce}")
    _need_cuda(other, f"spmm_{reduce}")
    out = _SpmmFn.apply(other, src._rowptr, src._col32(), src._fvalue(), src._sizes[0], _REDUCE[reduce])
    return out if other.dtype == out.dtype else out.to(other.dtype)


def spmm_add(src: SparseTensor, other: Tensor) -> Tensor:
    return _spmm(src, other, "sum")


spmm_sum = spmm_add


def spmm_mean(src: SparseTensor, other: Tensor) -> Tensor:
    return _spmm(src, other, "mean")


def spmm_max(src: SparseTensor, other: Tensor):
    return _spmm(src, other, "max"), None   # the reference reads [0] only (model.py:47)


def spmm(src: SparseTensor, other: Tensor, reduce: str = "sum") -> Tensor:
    return _spmm(src, other, reduce)


def matmul(src: SparseTensor, other: Tensor, reduce: str = "sum") -> Tensor:
    if not isinstance(other, Tensor):
        raise NotImplementedError("sparse @ sparse: use ocn_b200.spgemm_a2 (the reference multiplies torch COO tensors, "
                                  "NeighborOverlap_large.py:74)")
    return _spmm(src, other, reduce)
