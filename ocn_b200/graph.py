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
                 n_cols: Optional[int] = None):
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
        self._ws = {}

    # -- construction ------------------------------------------------------------------------
    @staticmethod
    def from_edge_index(edge_index: Tensor, n: int, symmetric: bool = True) -> "Graph":
        """Sort by (row, col), drop duplicates (and, like to_symmetric(), union both directions)."""
        src, dst = edge_index[0].to(torch.int64), edge_index[1].to(torch.int64)
        if symmetric:
            src, dst = torch.cat((src, dst)), torch.cat((dst, src))
        key = torch.unique(src * n + dst)
        row = torch.div(key, n, rounding_mode="floor")
        col = (key - row * n).to(torch.int32)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
        torch.cumsum(torch.bincount(row, minlength=n), 0, out=rowptr[1:])
        return Graph(rowptr, col, n)

    def to(self, device) -> "Graph":
        return Graph(self.rowptr.to(device), self.col.to(device), self.n,
                     None if self.value is None else self.value.to(device), self.n_cols)

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


def _require_cuda(t: Tensor):
    if not t.is_cuda:
        raise _lib.OcnError("ocn_b200 ops run on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
