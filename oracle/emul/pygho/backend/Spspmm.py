"""pygho.backend.Spspmm stand-in: expand + unique + scatter-sum spspmm, index-intersection hadamard."""
import torch

from .. import SparseTensor


def spsphadamard(A, B):
    w = A.shape[1]
    ka = A.indices[0] * w + A.indices[1]
    kb = B.indices[0] * w + B.indices[1]
    if ka.numel() == 0 or kb.numel() == 0:
        return SparseTensor(torch.zeros(2, 0, dtype=torch.long), torch.zeros(0), A.shape, is_coalesced=True)
    idx = torch.searchsorted(kb, ka).clamp_(max=kb.numel() - 1)
    hit = kb[idx] == ka
    return SparseTensor(A.indices[:, hit], A.values[hit] * B.values[idx[hit]], A.shape, is_coalesced=True)


def spspmm(A, dim1, B, dim2, aggr="sum"):
    assert dim1 == 1 and dim2 == 0
    n = B.shape[0]
    rp = torch.zeros(n + 1, dtype=torch.long)
    torch.cumsum(torch.bincount(B.indices[0], minlength=n), 0, out=rp[1:])
    k = A.indices[1]
    start, cnt = rp[k], rp[k + 1] - rp[k]
    total = int(cnt.sum())
    owner = torch.repeat_interleave(torch.arange(A.nnz), cnt)
    pos = torch.arange(total) + torch.repeat_interleave(start - (torch.cumsum(cnt, 0) - cnt), cnt)
    key = A.indices[0][owner] * B.shape[1] + B.indices[1][pos]
    val = A.values[owner] * B.values[pos]
    uk, inv = torch.unique(key, return_inverse=True)
    out = torch.zeros(uk.numel(), dtype=val.dtype).index_add_(0, inv, val)
    return SparseTensor(torch.stack((torch.div(uk, B.shape[1], rounding_mode="floor"), uk % B.shape[1])), out,
                        (A.shape[0], B.shape[1]), is_coalesced=True)
