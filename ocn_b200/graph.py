"""Device-resident CSR graph handle (the torch_sparse.SparseTensor / pygho.SparseTensor of the
reference narrowed to what the hot path reads: rowptr int64, col int32, optional fp32 values)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


class Graph:
    """Square N x N adjacency in CSR, columns ascending and unique inside each row.

    Built the way the reference builds ``data.adj_t``:
    ``SparseTensor.from_edge_index(ei, sparse_sizes=(N, N)).to_symmetric()`` (ogbdataset.py:44-45,
    NeighborOverlap_large.py:59-63).
    """

    def __init__(self, rowptr: Tensor, col: Tensor, n: Optional[int] = None, value: Optional[Tensor] = None,
                 n_cols: Optional[int] = None, mult: Optional[Tensor] = None, symmetric: Optional[bool] = None):
        if rowptr.dtype != torch.int64:
            rowptr = rowptr.to(torch.int64)
        if col.dtype != torch.int32:
            col = col.to(torch.int32)
        self.rowptr = rowptr.contiguous()
        self.col = col.contiguous()
        self.n = int(n if n is not None else rowptr.numel() - 1)
        self.n_cols = int(n_cols if n_cols is not None else self.n)
        self.value = None if value is None else value.to(torch.float32).contiguous()
        if self.rowptr.numel() != self.n + 1:
            raise ValueError(f"rowptr has {self.rowptr.numel()} entries, expected n+1 = {self.n + 1}")
        self.mult = mult  # int32[nnz]: list edges per entry (from_edge_index(..., with_multiplicity=True)); for masked()
        # A == A^T (pattern and values)?  None = not known yet: is_symmetric() checks the pattern once on the device.
        # The GCN-normalised aggregation uses it to pick its backward (A-hat itself, or the scatter by A-hat^T).
        self.symmetric = symmetric
        self._ws = {}

    # -- construction ------------------------------------------------------------------------
    @staticmethod
    def from_edge_index(edge_index: Tensor, n: int, symmetric: bool = True, keep: Optional[Tensor] = None,
                        with_multiplicity: bool = False) -> "Graph":
        """Sort by (row, col), drop duplicates (and, like to_symmetric(), union both directions).
        ``keep`` is the reference's ``adjmask`` (bool [E]).  CUDA edge lists go through ``ocn_graph_build_*``
        (one radix sort + run-length encode in libocn_b200); a CPU edge list is only host-side data
        preparation (synthetic generators, tests) and yields a host graph no op accepts until it is moved."""
        if edge_index.is_cuda:
            return _build_on_device(edge_index, n, symmetric, keep, with_multiplicity)
        src, dst = edge_index[0].to(torch.int64), edge_index[1].to(torch.int64)
        if keep is not None:
            src, dst = src[keep], dst[keep]
        if symmetric:
            src, dst = torch.cat((src, dst)), torch.cat((dst, src))
        key = torch.unique(src * n + dst)
        row = torch.div(key, n, rounding_mode="floor")
        col = (key - row * n).to(torch.int32)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
        torch.cumsum(torch.bincount(row, minlength=n), 0, out=rowptr[1:])
        return Graph(rowptr, col, n, symmetric=True if symmetric else None)

    def is_symmetric(self) -> bool:
        """Pattern and values equal those of the transpose.  Known from the construction where possible; otherwise
        the pattern is checked once on the device (``ocn_graph_validate`` bit 3) and a matrix with explicit values
        counts as not symmetric."""
        if self.symmetric is None:
            self.symmetric = self.value is None and self.n == self.n_cols and (self.validate() & 8) == 0
        return self.symmetric

    def drop_entries(self, keep: Tensor, scale: float = 1.0) -> "Graph":
        """``torch_sparse.masked_select_nnz(adj, keep)`` followed by ``value * scale`` (DropAdj, model.py:219-229):
        a new CSR with the kept entries in order and fp32 values (val or 1) * scale."""
        _require_cuda(self.col)
        if keep.numel() != self.nnz:
            raise ValueError("keep must have one entry per stored element")
        L = _lib.lib()
        dev = self.device
        keep = keep.to(device=dev, dtype=torch.uint8).contiguous()
        counts = torch.empty(self.n, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.ocn_graph_select_count(_lib.ptr(self.rowptr), _lib.ptr(keep), self.n, _lib.ptr(counts), st),
                       "ocn_graph_select_count")
            rowptr = torch.zeros(self.n + 1, dtype=torch.int64, device=dev)
            torch.cumsum(counts, 0, out=rowptr[1:])
            nnz = int(rowptr[-1].item())
            col = torch.empty(nnz, dtype=torch.int32, device=dev)
            val = torch.empty(nnz, dtype=torch.float32, device=dev)
            if nnz:
                _lib.check(L.ocn_graph_select_fill(_lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.value),
                                                   _lib.ptr(keep), self.n, float(scale), _lib.ptr(rowptr), _lib.ptr(col),
                                                   _lib.ptr(val), st), "ocn_graph_select_fill")
        return Graph(rowptr, col, self.n, val, self.n_cols, symmetric=False)

    def masked(self, links: Tensor, symmetric: bool = True) -> "Graph":
        """The adjacency of one --maskinput training batch (NeighborOverlap_large.py:56-63): this graph rebuilt
        without the list edges ``links`` [2, M] -- by decrementing entry multiplicities and compacting the rows
        (``ocn_graph_mask_*``), not by re-sorting the edge list."""
        _require_cuda(self.col)
        _require_cuda(links)
        if self.mult is None:
            raise ValueError("masked() needs the multiplicity of every entry (how many list edges map onto it: duplicates, "
                             "both directions of one edge): build the graph with from_edge_index(..., with_multiplicity=True)."
                             "  Without it, masking one list edge would delete an entry another list edge still supplies "
                             "(the reference re-sorts the remaining list and keeps it, NeighborOverlap_large.py:56-63)")
        L = _lib.lib()
        dev = self.device
        src, dst = links[0].to(torch.int64).contiguous(), links[1].to(torch.int64).contiguous()
        M = int(src.numel())
        dec_key = ("mask_dec", torch.cuda.current_stream(dev).cuda_stream)   # one work array per stream, zero between calls
        dec = self._ws.get(dec_key)
        if dec is None:
            dec = torch.zeros(max(1, self.nnz), dtype=torch.int32, device=dev)
            self._ws[dec_key] = dec
        nb = L.ocn_graph_mask_bytes(self.n, M)
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        rowptr = torch.empty(self.n + 1, dtype=torch.int64, device=dev)
        info = torch.zeros(2, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            self._ws.pop(dec_key)  # put back only when both calls went through: a failed call leaves it dirty
            _lib.check(L.ocn_graph_mask_count(_lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.mult), self.n,
                                              self.nnz, _lib.ptr(src), _lib.ptr(dst), M, int(symmetric), _lib.ptr(dec),
                                              _lib.ptr(scratch), nb, _lib.ptr(rowptr), _lib.ptr(info), st),
                       "ocn_graph_mask_count")
            nnz, missing = info.tolist()
            col = torch.empty(nnz, dtype=torch.int32, device=dev)
            mult = torch.empty(nnz, dtype=torch.int32, device=dev) if self.mult is not None else None
            _lib.check(L.ocn_graph_mask_fill(_lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.mult), self.n,
                                             self.nnz, _lib.ptr(src), _lib.ptr(dst), M, int(symmetric), _lib.ptr(dec),
                                             _lib.ptr(scratch), _lib.ptr(col) if nnz else None, _lib.ptr(mult), st),
                       "ocn_graph_mask_fill")
        self._ws[dec_key] = dec
        if missing:
            raise ValueError(f"{missing} masked links are not edges of this graph")
        return Graph(rowptr, col, self.n, None, self.n_cols, mult, symmetric=self.symmetric if symmetric else None)

    def to(self, device) -> "Graph":
        return Graph(self.rowptr.to(device), self.col.to(device), self.n,
                     None if self.value is None else self.value.to(device), self.n_cols,
                     None if self.mult is None else self.mult.to(device), self.symmetric)

    # -- accessors mirroring what the reference reads ------------------------------------------
    @property
    def device(self):
        return self.col.device

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def sizes(self) -> Tuple[int, int]:
        return (self.n, self.n_cols)

    def degree(self) -> Tensor:
        return self.rowptr[1:] - self.rowptr[:-1]

    def row(self) -> Tensor:
        return torch.repeat_interleave(torch.arange(self.n, device=self.device), self.degree())

    def coo(self):
        return self.row(), self.col.to(torch.int64), self.value

    def csr(self):
        return self.rowptr, self.col.to(torch.int64), self.value

    # -- checks -----------------------------------------------------------------------------
    def validate(self) -> int:
        """Bit flags (0 = fine): 1 column out of range, 2 row not strictly ascending,
        4 rowptr broken, 8 not symmetric."""
        _require_cuda(self.col)
        flags = torch.zeros(1, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ocn_graph_validate(_lib.ptr(self.rowptr), _lib.ptr(self.col), self.n, self.nnz,
                                                     _lib.ptr(flags), st), "ocn_graph_validate")
        return int(flags.item())


def _build_on_device(edge_index: Tensor, n: int, symmetric: bool, keep: Optional[Tensor], with_mult: bool) -> Graph:
    L = _lib.lib()
    dev = edge_index.device
    src, dst = edge_index[0].to(torch.int64).contiguous(), edge_index[1].to(torch.int64).contiguous()
    E = int(src.numel())
    if keep is not None:
        keep = keep.to(device=dev, dtype=torch.uint8).contiguous()
        if keep.numel() != E:
            raise ValueError("keep mask must have one entry per edge")
    nb = L.ocn_graph_build_bytes(E, int(symmetric))
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    info = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.ocn_graph_build_count(_lib.ptr(src), _lib.ptr(dst), _lib.ptr(keep), E, n, int(symmetric),
                                           _lib.ptr(scratch), nb, _lib.ptr(rowptr), _lib.ptr(info), st),
                   "ocn_graph_build_count")
        nnz, bad = info.tolist()
        if bad:
            raise ValueError(f"{bad} edges have an endpoint outside [0, {n})")
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        mult = torch.empty(nnz, dtype=torch.int32, device=dev) if with_mult else None
        _lib.check(L.ocn_graph_build_fill(_lib.ptr(scratch), E, int(symmetric), n, nnz, _lib.ptr(col), _lib.ptr(mult), st),
                   "ocn_graph_build_fill")
    return Graph(rowptr, col, n, None, None, mult, symmetric=True if symmetric else None)


def _require_cuda(t: Tensor):
    if not t.is_cuda:
        raise _lib.OcnError("ocn_b200 ops run on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
