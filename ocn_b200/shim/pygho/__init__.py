"""Stand-in for ``pygho.SparseTensor`` (the subset qingpingmo/OCN touches) on libocn_b200's CUDA kernels.

Census: constructor (NeighborOverlapCitation2.py:151, model.py:2283-2286), ``index_select([0], idx)`` (:79-81),
``to_torch_sparse_coo`` (:88-89, model.py:104-140), ``tuplewiseapply`` / ``sum(dims=1)`` / ``indices`` / ``values``
(model.py:98-141), ``shape`` / ``nnz``.

Row selections and products are LAZY: ``adj.index_select([0], idx)`` and ``spspmm(Ej, 1, adj, 0)`` return objects
that remember how they were made.  ``spsphadamard(Ei, Ej)`` / ``spsphadamard(Ei, spspmm(Ej, 1, adj, 0))`` -- the
text of ``get_cn1_cn2`` (NeighborOverlapCitation2.py:78-85) -- then runs the fused common-neighbour kernels on the
graph itself (``ocn_rows_intersect_*``, ``ocn_cn_plan/build/extract_*``) and ``Ej . A`` is never materialised.  Any
other use of a lazy object (``.indices``, ``.values``, a product with something else) materialises it on the GPU.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor

from ... import _lib
from ...graph import Graph


def _need_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.OcnError(f"pygho shim: {what} runs on CUDA tensors only (ocn_b200 has no CPU fallback); got {t.device}")


class CooView:
    """What ``to_torch_sparse_coo()`` hands back: reads like a coalesced torch sparse COO tensor (``indices()``,
    ``values()``, ``shape``, ``coalesce()``, ``to_dense()``) and multiplies a dense matrix through ``ocn_spmm_csr``
    (``adj_t.to_torch_sparse_coo() @ x``, model.py:104-140) instead of cuSPARSE.  Anything else is forwarded to a real
    ``torch.sparse_coo_tensor`` built on first use."""

    def __init__(self, owner: "SparseTensor"):
        self._owner = owner
        self._real = None

    @property
    def shape(self):
        return torch.Size(self._owner.shape)

    @property
    def device(self):
        return self._owner.indices.device

    @property
    def is_sparse(self) -> bool:
        return True

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    def indices(self) -> Tensor:
        return self._owner.indices

    def values(self) -> Tensor:
        return self._owner.values

    def _nnz(self) -> int:
        return self._owner.nnz

    def coalesce(self) -> "CooView":
        return self

    def is_coalesced(self) -> bool:
        return True

    def torch(self) -> Tensor:
        if self._real is None:
            o = self._owner
            v = o.values if o.values is not None else torch.ones(o.nnz, device=o.indices.device)
            self._real = torch.sparse_coo_tensor(o.indices, v, o.shape, is_coalesced=True)
        return self._real

    def to_dense(self) -> Tensor:
        return self.torch().to_dense()

    def __matmul__(self, x: Tensor) -> Tensor:
        from ...sparse_ops import _SpmmFn
        o = self._owner
        _need_cuda(x, "sparse @ dense")
        rowptr, col = o._csr()
        v = o.values
        if v is not None and v.dim() != 1:
            return self.torch() @ x
        out = _SpmmFn.apply(x, rowptr, col, None if v is None else v.float().contiguous(), o.shape[0], 0)
        return out if out.dtype == x.dtype else out.to(x.dtype)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.torch(), name)


class SparseTensor:
    """2-D pygho sparse tensor: ``indices int64 [2, nnz]`` sorted by (row, col), ``values``, ``shape``."""

    def __init__(self, indices: Tensor, values: Optional[Tensor] = None, shape=None, is_coalesced: bool = False,
                 reduce_op: str = "sum"):
        indices = indices.to(torch.int64)
        if shape is None:
            shape = tuple((indices.max(dim=1).values + 1).tolist()) if indices.numel() else (0, 0)
        self._shape = tuple(int(s) for s in shape)
        if not is_coalesced and indices.shape[1] > 1:
            v = values if values is not None else torch.ones(indices.shape[1], device=indices.device)
            t = torch.sparse_coo_tensor(indices, v, self._shape[:2] + tuple(v.shape[1:])).coalesce()
            indices, values = t.indices(), (t.values() if values is not None else None)
        self._indices, self._values = indices, values
        self._rowptr = self._col32 = None
        self._unit = None          # values are all ones (decided once, on first need)
        self._graph = None

    # -- storage (lazy subclasses override _materialise) --------------------------------------
    def _materialise(self) -> None:
        pass

    @property
    def indices(self) -> Tensor:
        self._materialise()
        return self._indices

    @property
    def values(self) -> Optional[Tensor]:
        self._materialise()
        return self._values

    @property
    def shape(self):
        return self._shape

    @property
    def sparseshape(self):
        return self._shape

    @property
    def nnz(self) -> int:
        return int(self.indices.shape[1])

    @property
    def sparse_dim(self) -> int:
        return 2

    @property
    def device(self):
        return self.indices.device

    def is_coalesced(self) -> bool:
        return True

    def _csr(self):
        if self._rowptr is None:
            ind = self.indices
            rp = torch.zeros(self._shape[0] + 1, dtype=torch.int64, device=ind.device)
            if ind.shape[1]:
                torch.cumsum(torch.bincount(ind[0], minlength=self._shape[0])[:self._shape[0]], 0, out=rp[1:])
            self._rowptr, self._col32 = rp, ind[1].to(torch.int32).contiguous()
        return self._rowptr, self._col32

    def _fvalues(self) -> Optional[Tensor]:
        v = self.values
        return None if v is None else v.reshape(v.shape[0], -1)[:, 0].to(torch.float32).contiguous()

    def unit_valued(self) -> bool:
        if self._unit is None:
            v = self.values
            self._unit = v is None or (v.dim() == 1 and bool((v == 1).all()))
        return self._unit

    def graph(self) -> Graph:
        """Square unit-valued matrix as the ``ocn_b200.Graph`` the fused kernels take (buffers shared)."""
        if self._graph is None:
            rp, c = self._csr()
            self._graph = Graph(rp, c, self._shape[0], None, self._shape[1])
        return self._graph

    # -- the methods the reference calls ------------------------------------------------------
    def index_select(self, dims, index: Tensor) -> "SparseTensor":
        dims = [dims] if isinstance(dims, int) else list(dims)
        if dims != [0]:
            raise NotImplementedError("the reference only selects rows: index_select([0], idx)")
        return RowSelect(self, index.reshape(-1))

    def to_torch_sparse_coo(self) -> CooView:
        return CooView(self)

    def tuplewiseapply(self, fn: Callable[[Tensor], Tensor]) -> "SparseTensor":
        out = SparseTensor(self.indices, fn(self.values), self._shape, is_coalesced=True)
        out._rowptr, out._col32 = self._rowptr, self._col32
        return out

    def sum(self, dims=1) -> Tensor:
        d = dims if isinstance(dims, int) else list(dims)[0]
        _need_cuda(self.indices, "SparseTensor.sum")
        rp, col = self._csr()
        v = self._fvalues()
        dev = self.indices.device
        dtype = self.values.dtype if self.values is not None else torch.get_default_dtype()
        if d == 1:
            if v is None or self.unit_valued():
                return (rp[1:] - rp[:-1]).to(dtype)
            from ...sparse_ops import _spmm_raw
            ones = torch.ones(self._shape[1], 1, dtype=torch.float32, device=dev)
            return _spmm_raw(rp, col, v, self._shape[0], ones, 0).view(-1).to(dtype)
        out = torch.zeros(self._shape[1], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ocn_csr_colsum(_lib.ptr(col), _lib.ptr(v), int(col.numel()), self._shape[1],
                                                 _lib.ptr(out), torch.cuda.current_stream(dev).cuda_stream), "ocn_csr_colsum")
        return out.to(dtype)

    def to(self, device, non_blocking: bool = False) -> "SparseTensor":
        return SparseTensor(self.indices.to(device, non_blocking=non_blocking),
                            None if self.values is None else self.values.to(device, non_blocking=non_blocking),
                            self._shape, is_coalesced=True)

    def __repr__(self) -> str:
        return f"pygho.SparseTensor(shape={self._shape}, lazy={type(self).__name__ != 'SparseTensor'})  # ocn_b200 shim"


class RowSelect(SparseTensor):
    """``parent.index_select([0], idx)`` not yet materialised (row b = row idx[b] of the parent)."""

    def __init__(self, parent: SparseTensor, idx: Tensor):
        self.parent, self.idx = parent, idx.to(torch.int64)
        self._shape = (int(idx.numel()), parent.shape[1])
        self._indices = self._values = None
        self._rowptr = self._col32 = None
        self._unit = parent._unit
        self._graph = None
        self._done = False

    def _materialise(self) -> None:
        if self._done:
            return
        from ..torch_sparse.tensor import SparseTensor as TS, gather_rows
        p = self.parent
        rp, col = p._csr()
        src = TS._from_csr(rp, col, p._fvalues() if p.values is not None else None, p.shape)
        out = gather_rows(src, self.idx)
        self._rowptr, self._col32 = out._rowptr, out._col32()
        self._indices = torch.stack((out._row64(), out._col64()))
        v = out._value
        if v is not None and p.values is not None:
            v = v.to(p.values.dtype)
        self._values = v
        self._done = True


class Product(SparseTensor):
    """``spspmm(left, 1, right, 0)`` not yet materialised."""

    def __init__(self, left: SparseTensor, right: SparseTensor):
        self.left, self.right = left, right
        self._shape = (left.shape[0], right.shape[1])
        self._indices = self._values = None
        self._rowptr = self._col32 = None
        self._unit = None
        self._graph = None
        self._done = False

    def chain(self):
        """(row selection, graph, number of products) when this is ((adj[idx] . adj) . adj ...) over ONE matrix."""
        k, node = 0, self
        while isinstance(node, Product):
            if k and node.right is not self.right:
                return None
            k, node = k + 1, node.left
        if isinstance(node, RowSelect) and node.parent is self.right:
            return node, self.right, k
        return None

    def _materialise(self) -> None:
        if self._done:
            return
        from .backend.Spspmm import _spspmm_explicit
        out = _spspmm_explicit(self.left, self.right)
        self._indices, self._values = out._indices, out._values
        self._done = True
