"""SpMM / GCN aggregation / A^2 SpGEMM wrappers (pieces 2 and 3 of north_star) with autograd.

Reference call sites: ``spmm_add/spmm_mean/spmm_max`` (model.py:6,45-53,2426-2427),
``PureConv`` (model.py:42-55), ``PureConv3`` (model.py:128-142), ``GCNConv`` via ``convdict``
(model.py:58-71), ``spadj @ spadj`` and ``sparse_tensor_multiply`` (NeighborOverlap_large.py:68-74,
utils.py:287-329).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .cn import SparseRows, _stream
from .graph import Graph, _require_cuda

_REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2}


def _csr_of(src):
    if isinstance(src, Graph):
        return src.rowptr, src.col, src.value, src.n
    if isinstance(src, SparseRows):
        return src.rowptr, src.col.to(torch.int32), src.value, src.shape[0]
    raise TypeError(f"expected Graph or SparseRows, got {type(src)}")


def _spmm_raw(rowptr, col, val, rows, x, reduce: int) -> Tensor:
    _require_cuda(x)
    x = x.contiguous()
    out = torch.empty(rows, x.shape[1], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().ocn_spmm_csr(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), rows, _lib.ptr(x),
                                           x.shape[1], reduce, _lib.ptr(out), _stream(x.device)), "ocn_spmm_csr")
    return out


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rowptr, col, val, rows, reduce):
        xf = x.float().contiguous()
        ctx.save_for_backward(rowptr, col, val, xf if reduce == 2 else None)
        ctx.rows, ctx.reduce, ctx.xshape = rows, reduce, x.shape
        return _spmm_raw(rowptr, col, val, rows, xf, reduce)

    @staticmethod
    def backward(ctx, g):
        rowptr, col, val, xf = ctx.saved_tensors
        g = g.contiguous().float()
        gx = torch.zeros(ctx.xshape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            if ctx.reduce == 2:
                _lib.check(_lib.lib().ocn_spmm_csr_max_bwd(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), ctx.rows,
                                                           _lib.ptr(xf), _lib.ptr(g), g.shape[1], _lib.ptr(gx),
                                                           _stream(g.device)), "ocn_spmm_csr_max_bwd")
                return gx, None, None, None, None, None
            _lib.check(_lib.lib().ocn_spmm_csr_bwd(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), ctx.rows,
                                                   _lib.ptr(g), g.shape[1], ctx.reduce, _lib.ptr(gx),
                                                   _stream(g.device)), "ocn_spmm_csr_bwd")
        return gx, None, None, None, None, None


def spmm(src, other: Tensor, reduce: str = "sum") -> Tensor:
    rowptr, col, val, rows = _csr_of(src)
    return _SpmmFn.apply(other, rowptr, col, val, rows, _REDUCE[reduce])


def spmm_add(src, other: Tensor) -> Tensor:
    return spmm(src, other, "sum")


def spmm_mean(src, other: Tensor) -> Tensor:
    return spmm(src, other, "mean")


def spmm_max(src, other: Tensor):
    return spmm(src, other, "max"), None  # the reference only reads [0] (model.py:47)


# ---- GCN-normalised aggregation ---------------------------------------------------------------

def gcn_norm(g: Graph, edge_w: Optional[Tensor] = None) -> Tensor:
    """rsqrt(1 + rowsum(A)) (model.py:51, :136); the row sums run over ``edge_w`` (default: the graph's own values,
    e.g. the rescaled survivors of DropAdj, or ones)."""
    edge_w = edge_w if edge_w is not None else g.value
    out = torch.empty(g.n, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().ocn_gcn_norm(_lib.ptr(g.rowptr), _lib.ptr(edge_w), g.n, _lib.ptr(out), _stream(g.device)),
                   "ocn_gcn_norm")
    return out


def _gcn_raw(g: Graph, edge_w, norm, mode: int, x: Tensor) -> Tensor:
    x = x.contiguous()
    out = torch.empty(g.n, x.shape[1], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().ocn_gcn_spmm(_lib.ptr(g.rowptr), _lib.ptr(g.col), _lib.ptr(edge_w), g.n, _lib.ptr(norm),
                                           mode, _lib.ptr(x), x.shape[1], _lib.ptr(out), _stream(x.device)),
                   "ocn_gcn_spmm")
    return out


def _gcn_transpose_raw(g: Graph, edge_w, norm, mode: int, grad: Tensor) -> Tensor:
    """grad_x = A-hat^T grad for a matrix that is not its own transpose (a directed graph, or the adjacency after
    DropAdj, whose two directions of an edge are dropped independently, model.py:222):
        mode 3: A-hat = D (A_w + I) D   ->  D (A_w^T (D g)) + D^2 g
        mode 4: A-hat = D A_w D         ->  D (A_w^T (D g))
    with the transpose product as a scatter (ocn_spmm_csr_bwd: vector reductions into the rows of grad_x)."""
    grad = grad.contiguous().float()
    t = grad * norm.unsqueeze(1)
    u = torch.zeros_like(t)
    with torch.cuda.device(grad.device):
        _lib.check(_lib.lib().ocn_spmm_csr_bwd(_lib.ptr(g.rowptr), _lib.ptr(g.col), _lib.ptr(edge_w), g.n, _lib.ptr(t),
                                               t.shape[1], 0, _lib.ptr(u), _stream(grad.device)), "ocn_spmm_csr_bwd")
    u = u * norm.unsqueeze(1)
    if mode == 3:
        u = u + grad * (norm * norm).unsqueeze(1)
    return u


class _GcnFn(torch.autograd.Function):
    """out = A-hat x.  A symmetric unit-weight adjacency (what the reference's to_symmetric() graphs are) is its own
    transpose: grad_x = A-hat grad_out by the same gather kernel, run-to-run deterministic.  Anything else -- explicit
    edge values, a directed graph -- takes the transpose scatter."""

    @staticmethod
    def forward(ctx, x, graph, norm, mode):
        ctx.graph, ctx.mode = graph, mode
        ctx.self_adjoint = graph.value is None and graph.is_symmetric()
        ctx.save_for_backward(norm)
        return _gcn_raw(graph, graph.value, norm, mode, x.float())

    @staticmethod
    def backward(ctx, g):
        (norm,) = ctx.saved_tensors
        if ctx.self_adjoint:
            return _gcn_raw(ctx.graph, None, norm, ctx.mode, g.contiguous().float()), None, None, None
        return _gcn_transpose_raw(ctx.graph, ctx.graph.value, norm, ctx.mode, g), None, None, None


def drop_adj(adj: Graph, dp: float, training: bool = True, doscale: bool = True,
             generator: Optional[torch.Generator] = None) -> Graph:
    """``DropAdj.forward`` (model.py:219-229): in training mode every stored entry survives with probability 1 - dp
    (``torch.rand_like(col) > dp``: the mask comes from torch's generator on the graph's device, as in the reference)
    and the survivors carry 1 / (1 - dp) (times their value).  The two directions of an edge are dropped
    independently, so the result is not symmetric."""
    if dp < 1e-6 or not training:
        return adj
    _require_cuda(adj.col)
    mask = torch.rand(adj.nnz, dtype=torch.float32, device=adj.device, generator=generator) > dp
    return adj.drop_entries(mask, 1.0 / (1.0 - dp) if doscale else 1.0)


def pure_conv(x: Tensor, adj: Graph, aggr: str = "gcn", norm: Optional[Tensor] = None) -> Tensor:
    """``PureConv.forward`` (model.py:42-55).  ``adj`` may carry values (DropAdj): they weigh the sums and the
    degree normalisation exactly as ``spmm_*`` / ``adj_t.sum(dim=-1)`` do on a valued SparseTensor."""
    _require_cuda(x)
    if aggr == "gcn":
        return _GcnFn.apply(x, adj, norm if norm is not None else gcn_norm(adj), 3)
    return spmm(adj, x, aggr)


def pure_conv3_gcn(x: Tensor, adj: Graph, norm: Optional[Tensor] = None) -> Tensor:
    """``PureConv3.forward`` aggr='gcn' before its Linear (model.py:135-141)."""
    _require_cuda(x)
    return _GcnFn.apply(x, adj, norm if norm is not None else gcn_norm(adj), 4)


def gcnconv_propagate(x: Tensor, adj: Graph, normalize: bool, add_self_loops: bool = True, aggr: str = "sum") -> Tensor:
    """Neighbour aggregation of PyG ``GCNConv`` as configured by ``convdict`` (model.py:58-71)."""
    if not normalize:
        return spmm(adj, x, aggr)
    if not add_self_loops:
        raise NotImplementedError("convdict never builds GCNConv(normalize=True, add_self_loops=False)")
    if adj.value is not None:
        # gcn_norm of PyG on a valued matrix adds unit self loops and normalises by the weighted degree: D (A_w + I) D
        return _GcnFn.apply(x, adj, gcn_norm(adj), 3)
    return _GcnFn.apply(x, adj, gcn_norm(adj), 3)


# ---- A^2 --------------------------------------------------------------------------------------

def spgemm_a2(adj: Graph, fold: int = 0, with_value: bool = False) -> Graph:
    """``spadj @ spadj`` (fold=0) or the reference's ``sparse_tensor_multiply(spadj, fold)``."""
    _require_cuda(adj.col)
    L = _lib.lib()
    dev = adj.device
    with torch.cuda.device(dev):
        st = _stream(dev)
        scratch = torch.zeros(L.ocn_spgemm_scratch_bytes(adj.n, adj.nnz, int(fold)), dtype=torch.uint8, device=dev)
        row_nnz = torch.zeros(adj.n, dtype=torch.int64, device=dev)
        _lib.check(L.ocn_spgemm_a2_symbolic(_lib.ptr(adj.rowptr), _lib.ptr(adj.col), adj.n, adj.nnz, int(fold),
                                            _lib.ptr(scratch), _lib.ptr(row_nnz), st), "ocn_spgemm_a2_symbolic")
        rowptr = torch.zeros(adj.n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_nnz, 0, out=rowptr[1:])
        nnz = int(rowptr[-1].item())
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev) if with_value else None
        if nnz:
            _lib.check(L.ocn_spgemm_a2_numeric(_lib.ptr(adj.rowptr), _lib.ptr(adj.col), adj.n, adj.nnz, int(fold),
                                               _lib.ptr(scratch), _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), st),
                       "ocn_spgemm_a2_numeric")
    return Graph(rowptr, col, adj.n, val)


def sparse_tensor_multiply(spadj: Graph, block_size: int = 1024) -> Graph:
    """utils.py:326-329 as written (folded)."""
    return spgemm_a2(spadj, fold=block_size, with_value=True)
