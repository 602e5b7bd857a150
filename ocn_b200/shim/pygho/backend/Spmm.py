"""``pygho.backend.Spmm.spmm`` (model.py:100-102, 130-132: mean / max aggregation of PureConv2 / PureConv3) on
``ocn_spmm_csr``."""
from __future__ import annotations

from torch import Tensor

from .. import SparseTensor, _need_cuda
from ....sparse_ops import _SpmmFn

_AGGR = {"sum": 0, "add": 0, "mean": 1, "max": 2, "amax": 2}


def spmm(A: SparseTensor, dim1: int, X: Tensor, aggr: str = "sum") -> Tensor:
    if dim1 != 1:
        raise NotImplementedError("the reference only calls spmm(A, 1, X)")
    _need_cuda(X, "spmm")
    rowptr, col = A._csr()
    out = _SpmmFn.apply(X, rowptr, col, A._fvalues(), A.shape[0], _AGGR[aggr])
    return out if out.dtype == X.dtype else out.to(X.dtype)
