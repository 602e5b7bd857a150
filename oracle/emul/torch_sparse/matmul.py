"""torch_sparse.matmul stand-in: CSR SpMM with sum / mean / max reduction."""
import torch


def _vals(src, like):
    return src._value if src._value is not None else torch.ones(src.nnz(), dtype=like.dtype)


def spmm_add(src, other):
    out = torch.zeros(src.size(0), other.size(1), dtype=other.dtype)
    return out.index_add_(0, src._row, _vals(src, other).unsqueeze(1) * other[src._col])


spmm_sum = spmm_add


def spmm_mean(src, other):
    deg = torch.bincount(src._row, minlength=src.size(0)).clamp(min=1).to(other.dtype)
    return spmm_add(src, other) / deg.unsqueeze(1)


def spmm_max(src, other):
    out = torch.full((src.size(0), other.size(1)), float("-inf"), dtype=other.dtype)
    out.index_reduce_(0, src._row, _vals(src, other).unsqueeze(1) * other[src._col], "amax", include_self=True)
    out[torch.isinf(out)] = 0
    return out, None


def matmul(src, other, reduce="sum"):
    return {"sum": spmm_add, "add": spmm_add, "mean": spmm_mean}[reduce](src, other)
