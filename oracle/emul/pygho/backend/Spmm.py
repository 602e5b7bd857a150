"""pygho.backend.Spmm stand-in (imported by the reference, only used by mean/max PureConv2/3)."""
import torch


def spmm(A, dim1, X, aggr="sum"):
    v = A.values.reshape(A.values.shape[0], -1)
    out = torch.zeros(A.shape[0], X.shape[1], dtype=X.dtype)
    if aggr == "sum":
        return out.index_add_(0, A.indices[0], v * X[A.indices[1]])
    if aggr == "mean":
        deg = torch.bincount(A.indices[0], minlength=A.shape[0]).clamp(min=1).to(X.dtype)
        return out.index_add_(0, A.indices[0], v * X[A.indices[1]]) / deg.unsqueeze(1)
    raise NotImplementedError(aggr)
