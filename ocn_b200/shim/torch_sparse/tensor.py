"""``torch_sparse.SparseTensor`` on the CUDA kernels of libocn_b200 -- the subset the reference touches.

Census (reference call sites): constructor (utils.py:146-151, model.py:2277, NeighborOverlapCitation2.py:101-102),
``from_edge_index`` + ``to_symmetric`` (NeighborOverlap_large.py:59-63), ``from_torch_sparse_coo_tensor`` (:70-74),
``from_dense`` / ``+`` (utils.py:311-321), ``adj[idx]`` (utils.py:256-257), ``storage.row/col/value/rowcount``
(utils.py:156-157, 116, model.py:222-226), ``sum`` / ``mul`` (model.py:2261-2272), ``coo`` / ``csr`` / ``sizes`` /
``size`` / ``device`` / ``to_torch_sparse_coo_tensor`` / ``to_dense`` / ``fill_value_`` / ``set_value_`` /
``coalesce`` and ``masked_select_nnz`` (model.py:223).

Layout: rows ascending, columns ascending inside a row; ``rowptr int64``, ``col`` kept twice on demand (int32 for the
kernels, int64 for callers that do arithmetic on it, utils.py:156).  Everything that computes runs on CUDA tensors
through ``include/ocn_b200.h`` (row gather, column sums, SpMM, graph build, entry selection) with torch CUDA ops for
the bookkeeping around them; a matrix on the CPU is a container only (the drivers build ``data.adj_t`` on the host and
move it, ogbdataset.py:44-45) -- computing on it raises ``OcnError``: there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ... import _lib
from ...graph import Graph


def _need_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.OcnError(f"torch_sparse shim: {what} runs on CUDA tensors only (ocn_b200 has no CPU fallback); "
                            f"the matrix is on {t.device} -- move it with .to(device) first")


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _Storage:
    """``SparseTensor.storage``: the accessors utils.py / model.py read."""

    def __init__(self, owner: "SparseTensor"):
        self._o = owner

    def row(self) -> Tensor:
        return self._o._row64()

    def col(self) -> Tensor:
        return self._o._col64()

    def value(self) -> Optional[Tensor]:
        return self._o._value

    def has_value(self) -> bool:
        return self._o._value is not None

    def rowptr(self) -> Tensor:
        return self._o._rowptr

    def rowcount(self) -> Tensor:
        rp = self._o._rowptr
        return rp[1:] - rp[:-1]

    def sparse_sizes(self):
        return self._o._sizes

    def set_value_(self, value, layout=None):
        self._o._value = value
        return self


class SparseTensor:
    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None, col: Optional[Tensor] = None,
                 value: Optional[Tensor] = None, sparse_sizes=None, is_sorted: bool = False, trust_data: bool = False):
        if col is None or (row is None and rowptr is None):
            raise ValueError("SparseTensor needs col and one of row / rowptr")
        dev = col.device
        if sparse_sizes is None or sparse_sizes[0] is None or sparse_sizes[1] is None:
            m = (rowptr.numel() - 1) if rowptr is not None else (int(row.max()) + 1 if row.numel() else 0)
            n = int(col.max()) + 1 if col.numel() else 0
            given = sparse_sizes or (None, None)
            sparse_sizes = (m if given[0] is None else given[0], n if given[1] is None else given[1])
        sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        unique = False
        if rowptr is None:
            row = row.to(torch.int64)
            col = col.to(torch.int64) if col.dtype != torch.int32 else col
            if row.numel() <= 1:
                unique = True
            elif not is_sorted:
                key = row * max(sizes[1], 1) + col
                step = int((key[1:] - key[:-1]).min())           # one read-back: torch_sparse checks before it sorts, too
                if step < 0:
                    perm = torch.argsort(key, stable=True)
                    row, col = row[perm], col[perm]
                    value = None if value is None else value[perm]
                else:
                    unique = step > 0
            rowptr = torch.zeros(sizes[0] + 1, dtype=torch.int64, device=dev)
            if row.numel():
                torch.cumsum(torch.bincount(row, minlength=sizes[0])[:sizes[0]], 0, out=rowptr[1:])
            self._row = row
        else:
            rowptr = rowptr.to(torch.int64)
            self._row = None if row is None else row.to(torch.int64)
        self._rowptr = rowptr.contiguous()
        self._col = col.contiguous()          # int64 or int32, whichever arrived; the other is made on demand
        self._col_other = None
        self._value = value
        self._sizes = sizes
        self._unique = unique                  # rows hold no duplicate column (known for matrices our kernels wrote)
        self._graph = None
        self.storage = _Storage(self)

    # ---- internal views --------------------------------------------------------------------
    @classmethod
    def _from_csr(cls, rowptr: Tensor, col: Tensor, value: Optional[Tensor], sizes, unique: bool = True) -> "SparseTensor":
        out = cls(rowptr=rowptr, col=col, value=value, sparse_sizes=sizes, is_sorted=True)
        out._unique = unique
        return out

    def _row64(self) -> Tensor:
        if self._row is None:
            rp = self._rowptr
            self._row = torch.repeat_interleave(torch.arange(self._sizes[0], device=rp.device), rp[1:] - rp[:-1],
                                                output_size=self.nnz())
        return self._row

    def _col64(self) -> Tensor:
        if self._col.dtype == torch.int64:
            return self._col
        if self._col_other is None:
            self._col_other = self._col.to(torch.int64)
        return self._col_other

    def _col32(self) -> Tensor:
        if self._col.dtype == torch.int32:
            return self._col
        if self._col_other is None:
            self._col_other = self._col.to(torch.int32)
        return self._col_other

    def _fvalue(self) -> Optional[Tensor]:
        v = self._value
        return None if v is None else v.to(torch.float32).contiguous()

    def graph(self) -> Graph:
        """The same matrix as the ``ocn_b200.Graph`` handle the fused ops take (shared buffers)."""
        if self._graph is None or self._graph.value is not self._value:
            self._graph = Graph(self._rowptr, self._col32(), self._sizes[0], self._fvalue(), self._sizes[1])
        return self._graph

    # ---- constructors ----------------------------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index: Tensor, edge_attr: Optional[Tensor] = None, sparse_sizes=None,
                        is_sorted: bool = False, trust_data: bool = False) -> "SparseTensor":
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes, is_sorted=is_sorted)

    @classmethod
    def from_torch_sparse_coo_tensor(cls, mat: Tensor, has_value: bool = True) -> "SparseTensor":
        mat = mat.coalesce()
        r, c = mat.indices()
        out = cls(row=r, col=c, value=mat.values() if has_value else None, sparse_sizes=tuple(mat.shape), is_sorted=True)
        out._unique = True
        return out

    @classmethod
    def from_dense(cls, mat: Tensor, has_value: bool = True) -> "SparseTensor":
        r, c = torch.nonzero(mat, as_tuple=True)
        out = cls(row=r, col=c, value=mat[r, c] if has_value else None, sparse_sizes=tuple(mat.shape), is_sorted=True)
        out._unique = True
        return out

    # ---- shape / access --------------------------------------------------------------------
    def sizes(self):
        return list(self._sizes)

    def sparse_sizes(self):
        return self._sizes

    def size(self, dim: int) -> int:
        return self._sizes[dim]

    def sparse_size(self, dim: int) -> int:
        return self._sizes[dim]

    def nnz(self) -> int:
        return int(self._col.numel())

    def numel(self) -> int:
        return self.nnz()

    def device(self):
        return self._col.device

    def is_cuda(self) -> bool:
        return self._col.is_cuda

    def has_value(self) -> bool:
        return self._value is not None

    def coo(self):
        return self._row64(), self._col64(), self._value

    def csr(self):
        return self._rowptr, self._col64(), self._value

    def to_device(self, device, non_blocking: bool = False) -> "SparseTensor":
        device = torch.device(device)
        if device == self._col.device or (device.type == "cuda" and device.index is None and self._col.is_cuda):
            return self
        out = SparseTensor._from_csr(self._rowptr.to(device, non_blocking=non_blocking),
                                     self._col.to(device, non_blocking=non_blocking),
                                     None if self._value is None else self._value.to(device, non_blocking=non_blocking),
                                     self._sizes, self._unique)
        return out

    def to(self, *args, **kwargs) -> "SparseTensor":
        device = kwargs.get("device")
        for a in args:
            if isinstance(a, (torch.device, str)) or (isinstance(a, int) and not isinstance(a, bool)):
                device = a
        out = self if device is None else self.to_device(device, kwargs.get("non_blocking", False))
        dtype = kwargs.get("dtype", next((a for a in args if isinstance(a, torch.dtype)), None))
        if dtype is not None and out._value is not None:
            out = SparseTensor._from_csr(out._rowptr, out._col, out._value.to(dtype), out._sizes, out._unique)
        return out

    def cuda(self, device=None) -> "SparseTensor":
        return self.to_device(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def cpu(self) -> "SparseTensor":
        return self.to_device("cpu")

    def fill_value_(self, fill_value: float, dtype=None) -> "SparseTensor":
        self._value = torch.full((self.nnz(),), fill_value, dtype=dtype or torch.get_default_dtype(), device=self._col.device)
        return self

    def fill_value(self, fill_value: float, dtype=None) -> "SparseTensor":
        return SparseTensor._from_csr(self._rowptr, self._col, None, self._sizes, self._unique).fill_value_(fill_value, dtype)

    def set_value_(self, value: Optional[Tensor], layout=None) -> "SparseTensor":
        self._value = value
        return self

    def set_value(self, value: Optional[Tensor], layout=None) -> "SparseTensor":
        return SparseTensor._from_csr(self._rowptr, self._col, value, self._sizes, self._unique)

    # ---- algebra ---------------------------------------------------------------------------
    def coalesce(self, reduce: str = "sum") -> "SparseTensor":
        if self._unique or self.nnz() == 0:
            return self
        key = self._row64() * max(self._sizes[1], 1) + self._col64()
        if bool((key[1:] > key[:-1]).all()):
            self._unique = True
            return self
        uk, inv = torch.unique(key, return_inverse=True)
        v = None
        if self._value is not None:
            v = torch.zeros(uk.numel(), dtype=self._value.dtype, device=key.device).index_add_(0, inv, self._value)
        w = max(self._sizes[1], 1)
        out = SparseTensor(row=torch.div(uk, w, rounding_mode="floor"), col=uk % w, value=v, sparse_sizes=self._sizes,
                           is_sorted=True)
        out._unique = True
        return out

    def to_symmetric(self, reduce: str = "sum") -> "SparseTensor":
        """Union of (r, c) and (c, r), duplicates merged (NeighborOverlap_large.py:63).  A matrix without values on the
        GPU goes through ``ocn_graph_build_*`` (one radix sort + run-length encode)."""
        n = max(self._sizes)
        if self._value is None and self._col.is_cuda:
            g = Graph.from_edge_index(torch.stack((self._row64(), self._col64())), n, symmetric=True)
            return SparseTensor._from_csr(g.rowptr, g.col, None, (n, n))
        row = torch.cat((self._row64(), self._col64()))
        col = torch.cat((self._col64(), self._row64()))
        v = None if self._value is None else torch.cat((self._value, self._value))
        return SparseTensor(row=row, col=col, value=v, sparse_sizes=(n, n)).coalesce(reduce)

    def sum(self, dim: Optional[int] = None) -> Tensor:
        dev = self._col.device
        if dim is None:
            return self._value.sum() if self._value is not None else torch.tensor(float(self.nnz()), device=dev)
        _need_cuda(self._col, "SparseTensor.sum")
        dtype = self._value.dtype if self._value is not None else torch.get_default_dtype()
        if dim in (0, -2):
            out = torch.zeros(self._sizes[1], dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().ocn_csr_colsum(_lib.ptr(self._col32()), _lib.ptr(self._fvalue()), self.nnz(),
                                                     self._sizes[1], _lib.ptr(out), _stream(dev)), "ocn_csr_colsum")
            return out.to(dtype)
        if self._value is None:
            return self.storage.rowcount().to(dtype)
        from ...sparse_ops import _spmm_raw
        ones = torch.ones(self._sizes[1], 1, dtype=torch.float32, device=dev)
        return _spmm_raw(self._rowptr, self._col32(), self._fvalue(), self._sizes[0], ones, 0).view(-1).to(dtype)

    def mul(self, other: Tensor) -> "SparseTensor":
        v = self._value if self._value is not None else torch.ones(self.nnz(), dtype=other.dtype, device=self._col.device)
        if other.dim() == 2 and other.size(0) == 1:
            nv = v * other[0, self._col64()]
        elif other.dim() == 2 and other.size(1) == 1:
            nv = v * other[self._row64(), 0]
        else:
            raise ValueError("mul: expected a [1, N] or [M, 1] dense operand")
        return SparseTensor._from_csr(self._rowptr, self._col, nv, self._sizes, self._unique)

    def __mul__(self, other):
        return self.mul(other)

    def __add__(self, other: "SparseTensor") -> "SparseTensor":
        """sparse + sparse: COO concatenation, sizes = element-wise max, duplicates summed (utils.py:321)."""
        sizes = (max(self._sizes[0], other._sizes[0]), max(self._sizes[1], other._sizes[1]))
        dt = torch.get_default_dtype()
        dev = self._col.device
        a = self._value if self._value is not None else torch.ones(self.nnz(), device=dev)
        b = other._value if other._value is not None else torch.ones(other.nnz(), device=dev)
        return SparseTensor(row=torch.cat((self._row64(), other._row64())), col=torch.cat((self._col64(), other._col64())),
                            value=torch.cat((a.to(dt), b.to(dt))), sparse_sizes=sizes).coalesce("sum")

    def index_select(self, dim: int, idx: Tensor) -> "SparseTensor":
        if dim != 0:
            raise NotImplementedError("the reference only gathers rows (utils.py:256-257)")
        return gather_rows(self, idx)

    def __getitem__(self, idx) -> "SparseTensor":
        if isinstance(idx, Tensor) and idx.dtype == torch.bool:
            idx = torch.nonzero(idx).flatten()
        if not isinstance(idx, Tensor):
            raise NotImplementedError("the reference only indexes a SparseTensor with a LongTensor / bool mask of rows")
        return gather_rows(self, idx)

    # ---- conversions -----------------------------------------------------------------------
    def to_torch_sparse_coo_tensor(self, dtype=None) -> Tensor:
        dev = self._col.device
        v = self._value if self._value is not None else torch.ones(self.nnz(), dtype=dtype or torch.get_default_dtype(),
                                                                   device=dev)
        idx = torch.stack((self._row64(), self._col64()))
        # entries are sorted; when they are known to be unique too the tensor is marked coalesced, so the reference's
        # .coalesce() calls (model.py:2293-2294) cost nothing instead of a sort
        return torch.sparse_coo_tensor(idx, v, self._sizes, is_coalesced=True if self._unique else None)

    def to_dense(self, dtype=None) -> Tensor:
        return self.to_torch_sparse_coo_tensor(dtype).to_dense()

    def t(self) -> "SparseTensor":
        return SparseTensor(row=self._col64(), col=self._row64(), value=self._value,
                            sparse_sizes=(self._sizes[1], self._sizes[0]))

    def __repr__(self) -> str:
        return (f"SparseTensor(nnz={self.nnz()}, sizes={self._sizes}, has_value={self._value is not None}, "
                f"device={self._col.device})  # ocn_b200 shim")


def gather_rows(src: SparseTensor, idx: Tensor) -> SparseTensor:
    """``src[idx]``: ``ocn_rows_gather_count/fill`` (row b of the result = row idx[b] of ``src``)."""
    _need_cuda(src._col, "row gather (adj[idx] / index_select)")
    dev = src._col.device
    idx = idx.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    B = int(idx.numel())
    L = _lib.lib()
    counts = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    rowptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    val = src._fvalue()
    with torch.cuda.device(dev):
        st = _stream(dev)
        _lib.check(L.ocn_rows_gather_count(_lib.ptr(src._rowptr), src._sizes[0], _lib.ptr(idx), B, _lib.ptr(counts), st),
                   "ocn_rows_gather_count")
        torch.cumsum(counts[:B], 0, out=rowptr[1:])
        nnz, bad = torch.stack((rowptr[-1], counts[B])).tolist()
        if bad:
            raise IndexError(f"index out of range: {bad} of the selected rows are outside [0, {src._sizes[0]})")
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        out_val = torch.empty(nnz, dtype=torch.float32, device=dev) if val is not None else None
        if nnz:
            _lib.check(L.ocn_rows_gather_fill(_lib.ptr(src._rowptr), _lib.ptr(src._col32()), _lib.ptr(val), src._sizes[0],
                                              _lib.ptr(idx), B, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(out_val), st),
                       "ocn_rows_gather_fill")
    if out_val is not None and src._value.dtype != torch.float32:
        out_val = out_val.to(src._value.dtype)
    return SparseTensor._from_csr(rowptr, col, out_val, (B, src._sizes[1]), src._unique)


def masked_select_nnz(src: SparseTensor, mask: Tensor, layout: Optional[str] = None) -> SparseTensor:
    """Keep the stored entries with ``mask`` set, in order (DropAdj, model.py:222-223): ``ocn_graph_select_*``."""
    _need_cuda(src._col, "masked_select_nnz")
    dev = src._col.device
    if mask.numel() != src.nnz():
        raise ValueError("mask must have one entry per stored element")
    L = _lib.lib()
    n = src._sizes[0]
    keep = mask.to(device=dev, dtype=torch.uint8).contiguous()
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    val = src._fvalue()
    with torch.cuda.device(dev):
        st = _stream(dev)
        if n:
            _lib.check(L.ocn_graph_select_count(_lib.ptr(src._rowptr), _lib.ptr(keep), n, _lib.ptr(counts), st),
                       "ocn_graph_select_count")
            torch.cumsum(counts, 0, out=rowptr[1:])
        nnz = int(rowptr[-1].item())
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        out_val = torch.empty(nnz, dtype=torch.float32, device=dev) if val is not None else None
        if nnz:
            _lib.check(L.ocn_graph_select_fill(_lib.ptr(src._rowptr), _lib.ptr(src._col32()), _lib.ptr(val), _lib.ptr(keep),
                                               n, 1.0, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(out_val), st),
                       "ocn_graph_select_fill")
    return SparseTensor._from_csr(rowptr, col, out_val, src._sizes, src._unique)
