"""SparseTensor stand-in: COO sorted by (row, col), optional value (None == implicit ones)."""
from __future__ import annotations

import torch


class _Storage:
    def __init__(self, owner):
        self._o = owner

    def row(self):
        return self._o._row

    def col(self):
        return self._o._col

    def value(self):
        return self._o._value

    def has_value(self):
        return self._o._value is not None

    def rowcount(self):
        return torch.bincount(self._o._row, minlength=self._o._sizes[0])

    def rowptr(self):
        rp = torch.zeros(self._o._sizes[0] + 1, dtype=torch.long)
        torch.cumsum(self.rowcount(), 0, out=rp[1:])
        return rp

    def set_value_(self, value, layout=None):
        self._o._value = value
        return self


class SparseTensor:
    def __init__(self, row=None, rowptr=None, col=None, value=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        if row is None:
            n = rowptr.numel() - 1
            row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
        row, col = row.long(), col.long()
        if sparse_sizes is None:
            sparse_sizes = (int(row.max()) + 1 if row.numel() else 0, int(col.max()) + 1 if col.numel() else 0)
        sizes = tuple(int(s) for s in sparse_sizes)
        if not is_sorted and row.numel():
            key = row * max(sizes[1], 1) + col
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            value = None if value is None else value[perm]
        self._row, self._col, self._value, self._sizes = row, col, value, sizes
        self.storage = _Storage(self)

    # ---- constructors
    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes, is_sorted=is_sorted)

    @classmethod
    def from_torch_sparse_coo_tensor(cls, mat, has_value=True):
        mat = mat.coalesce()
        r, c = mat.indices()
        return cls(row=r, col=c, value=mat.values() if has_value else None, sparse_sizes=mat.shape, is_sorted=True)

    @classmethod
    def from_dense(cls, mat, has_value=True):
        r, c = torch.nonzero(mat, as_tuple=True)
        return cls(row=r, col=c, value=mat[r, c] if has_value else None, sparse_sizes=mat.shape, is_sorted=True)

    # ---- shape / access
    def sizes(self):
        return list(self._sizes)

    def sparse_sizes(self):
        return self._sizes

    def size(self, dim):
        return self._sizes[dim]

    def nnz(self):
        return int(self._row.numel())

    def device(self):
        return self._row.device

    def to_device(self, device, non_blocking=False):
        return self

    def to(self, *a, **k):
        return self

    def coo(self):
        return self._row, self._col, self._value

    def csr(self):
        return self.storage.rowptr(), self._col, self._value

    def has_value(self):
        return self._value is not None

    def fill_value_(self, v, dtype=None):
        self._value = torch.full((self.nnz(),), v, dtype=dtype or torch.get_default_dtype())
        return self

    def fill_value(self, v, dtype=None):
        return SparseTensor(row=self._row, col=self._col, sparse_sizes=self._sizes, is_sorted=True).fill_value_(v, dtype)

    def set_value_(self, value, layout=None):
        self._value = value
        return self

    # ---- algebra
    def coalesce(self, reduce="sum"):
        key = self._row * max(self._sizes[1], 1) + self._col
        uk, inv = torch.unique(key, return_inverse=True)
        v = None
        if self._value is not None:
            v = torch.zeros(uk.numel(), dtype=self._value.dtype).index_add_(0, inv, self._value)
        w = max(self._sizes[1], 1)
        return SparseTensor(row=torch.div(uk, w, rounding_mode="floor"), col=uk % w, value=v, sparse_sizes=self._sizes,
                            is_sorted=True)

    def to_symmetric(self, reduce="sum"):
        n = max(self._sizes)
        row = torch.cat((self._row, self._col))
        col = torch.cat((self._col, self._row))
        v = None if self._value is None else torch.cat((self._value, self._value))
        return SparseTensor(row=row, col=col, value=v, sparse_sizes=(n, n)).coalesce(reduce)

    def sum(self, dim=None):
        v = self._value if self._value is not None else torch.ones(self.nnz(), dtype=torch.get_default_dtype())
        if dim is None:
            return v.sum()
        if dim in (0,):
            return torch.zeros(self._sizes[1], dtype=v.dtype).index_add_(0, self._col, v)
        return torch.zeros(self._sizes[0], dtype=v.dtype).index_add_(0, self._row, v)

    def mul(self, other):
        v = self._value if self._value is not None else torch.ones(self.nnz(), dtype=other.dtype)
        if other.dim() == 2 and other.size(0) == 1:
            nv = v * other[0, self._col]
        elif other.dim() == 2 and other.size(1) == 1:
            nv = v * other[self._row, 0]
        else:
            raise ValueError("mul: expected a [1,N] or [M,1] dense operand")
        return SparseTensor(row=self._row, col=self._col, value=nv, sparse_sizes=self._sizes, is_sorted=True)

    def __add__(self, other):
        sizes = (max(self._sizes[0], other._sizes[0]), max(self._sizes[1], other._sizes[1]))
        a = self._value if self._value is not None else torch.ones(self.nnz())
        b = other._value if other._value is not None else torch.ones(other.nnz())
        return SparseTensor(row=torch.cat((self._row, other._row)), col=torch.cat((self._col, other._col)),
                            value=torch.cat((a.to(torch.get_default_dtype()), b.to(torch.get_default_dtype()))),
                            sparse_sizes=sizes).coalesce("sum")

    def index_select(self, dim, idx):
        assert dim == 0
        rp = self.storage.rowptr()
        start, cnt = rp[idx], rp[idx + 1] - rp[idx]
        total = int(cnt.sum())
        out_row = torch.repeat_interleave(torch.arange(idx.numel()), cnt)
        pos = torch.arange(total) + torch.repeat_interleave(start - (torch.cumsum(cnt, 0) - cnt), cnt)
        return SparseTensor(row=out_row, col=self._col[pos], value=None if self._value is None else self._value[pos],
                            sparse_sizes=(idx.numel(), self._sizes[1]), is_sorted=True)

    def __getitem__(self, idx):
        if isinstance(idx, torch.Tensor) and idx.dtype == torch.bool:
            idx = torch.nonzero(idx).flatten()
        return self.index_select(0, idx)

    # ---- conversions
    def to_torch_sparse_coo_tensor(self, dtype=None):
        v = self._value if self._value is not None else torch.ones(self.nnz(), dtype=dtype or torch.get_default_dtype())
        return torch.sparse_coo_tensor(torch.stack((self._row, self._col)), v, self._sizes)

    def to_dense(self, dtype=None):
        return self.to_torch_sparse_coo_tensor(dtype).to_dense()

    def __repr__(self):
        return f"SparseTensor(row={self._row}, col={self._col}, val={self._value}, sizes={self._sizes})"


def masked_select_nnz(src, mask, layout="coo"):
    return SparseTensor(row=src._row[mask], col=src._col[mask], value=None if src._value is None else src._value[mask],
                        sparse_sizes=src._sizes, is_sorted=True)
