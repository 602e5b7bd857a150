"""Stand-in for pygho.SparseTensor (subset). TEST INFRASTRUCTURE ONLY -- see ../README.md."""
import torch


class SparseTensor:
    def __init__(self, indices, values=None, shape=None, is_coalesced=False, reduce_op="sum"):
        self.indices = indices.long()
        self.values = values
        self.shape = tuple(int(s) for s in shape)
        if not is_coalesced and self.indices.numel():
            t = torch.sparse_coo_tensor(self.indices, values, self.shape).coalesce()
            self.indices, self.values = t.indices(), t.values()

    @property
    def nnz(self):
        return self.indices.shape[1]

    def to_torch_sparse_coo(self):
        v = self.values if self.values is not None else torch.ones(self.nnz)
        return torch.sparse_coo_tensor(self.indices, v, self.shape).coalesce()

    def tuplewiseapply(self, fn):
        return SparseTensor(self.indices, fn(self.values), self.shape, is_coalesced=True)

    def sum(self, dims=1):
        d = dims if isinstance(dims, int) else dims[0]
        keep = 1 - d
        return torch.zeros(self.shape[keep], dtype=self.values.dtype).index_add_(0, self.indices[keep], self.values)

    def index_select(self, dims, index):
        """rows ``index[0]`` of a 2-D matrix, output row = position in ``index`` (fork-only API)."""
        assert list(dims) == [0]
        idx = index.reshape(-1)
        n = self.shape[0]
        rp = torch.zeros(n + 1, dtype=torch.long)
        torch.cumsum(torch.bincount(self.indices[0], minlength=n), 0, out=rp[1:])
        start, cnt = rp[idx], rp[idx + 1] - rp[idx]
        total = int(cnt.sum())
        out_row = torch.repeat_interleave(torch.arange(idx.numel()), cnt)
        pos = torch.arange(total) + torch.repeat_interleave(start - (torch.cumsum(cnt, 0) - cnt), cnt)
        return SparseTensor(torch.stack((out_row, self.indices[1][pos])), self.values[pos], (idx.numel(), self.shape[1]),
                            is_coalesced=True)
