"""PyTorch custom ops (`torch.ops.ocn.*`) over the C ABI -- CUDA-only registrations, no CPU kernels.

north_star: "hand-written CUDA kernels for sm_100a behind a thin C-ABI exposed as PyTorch custom ops".
The ops are functional (tensors in, tensors out) so they compose with autograd and `torch.compile`
callers; the stateful session API (`ocn_b200.cn.CNSession`) stays available for training loops that
need the records twice (forward and backward).

    torch.ops.ocn.rows_intersect(rowptr1, col1, rowptr2, col2, edges)            -> (rowptr, col)
    torch.ops.ocn.spmm_csr(rowptr, col, val, x, reduce)                           -> out        (autograd)
    torch.ops.ocn.gcn_spmm(rowptr, col, norm, x, mode)                            -> out        (autograd)
    torch.ops.ocn.spgemm_a2(rowptr, col, fold, with_value)                        -> (rowptr, col, val)
    torch.ops.ocn.cn_aggregate(rowptr, col, edges, x, ip3, batch_size, order,
                               weighted, variant, fill)                           -> (xcn1, xcn2, xcn3, xij)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import cn as _cn
from . import sparse_ops as _sp
from .graph import Graph


def _g(rowptr: Tensor, col: Tensor) -> Graph:
    return Graph(rowptr, col, rowptr.numel() - 1)


@torch.library.custom_op("ocn::rows_intersect", mutates_args=(), device_types="cuda")
def rows_intersect(rowptr1: Tensor, col1: Tensor, rowptr2: Tensor, col2: Tensor, edges: Tensor) -> Tuple[Tensor, Tensor]:
    out = _cn.adjoverlap(_g(rowptr1, col1), _g(rowptr2, col2), edges)
    return out.rowptr, out.col


@rows_intersect.register_fake
def _(rowptr1, col1, rowptr2, col2, edges):
    nnz = torch.library.get_ctx().new_dynamic_size()
    return rowptr1.new_empty(edges.shape[1] + 1), rowptr1.new_empty(nnz)


@torch.library.custom_op("ocn::spmm_csr", mutates_args=(), device_types="cuda")
def spmm_csr(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, reduce: int) -> Tensor:
    return _sp._spmm_raw(rowptr, col.to(torch.int32), val, rowptr.numel() - 1, x.float(), reduce)


@spmm_csr.register_fake
def _(rowptr, col, val, x, reduce):
    return x.new_empty(rowptr.numel() - 1, x.shape[1], dtype=torch.float32)


@torch.library.custom_op("ocn::spmm_csr_bwd", mutates_args=(), device_types="cuda")
def spmm_csr_bwd(rowptr: Tensor, col: Tensor, val: Optional[Tensor], grad_out: Tensor, n_cols: int, reduce: int) -> Tensor:
    from . import _lib
    g = grad_out.contiguous().float()
    gx = torch.zeros(n_cols, g.shape[1], dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().ocn_spmm_csr_bwd(_lib.ptr(rowptr), _lib.ptr(col.to(torch.int32)), _lib.ptr(val),
                                               rowptr.numel() - 1, _lib.ptr(g), g.shape[1], reduce, _lib.ptr(gx),
                                               _cn._stream(g.device)), "ocn_spmm_csr_bwd")
    return gx


@spmm_csr_bwd.register_fake
def _(rowptr, col, val, grad_out, n_cols, reduce):
    return grad_out.new_empty(n_cols, grad_out.shape[1], dtype=torch.float32)


@torch.library.custom_op("ocn::spmm_csr_max_bwd", mutates_args=(), device_types="cuda")
def spmm_csr_max_bwd(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, grad_out: Tensor) -> Tensor:
    from . import _lib
    g, xf = grad_out.contiguous().float(), x.contiguous().float()
    gx = torch.zeros_like(xf)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().ocn_spmm_csr_max_bwd(_lib.ptr(rowptr), _lib.ptr(col.to(torch.int32)), _lib.ptr(val),
                                                   rowptr.numel() - 1, _lib.ptr(xf), _lib.ptr(g), g.shape[1],
                                                   _lib.ptr(gx), _cn._stream(g.device)), "ocn_spmm_csr_max_bwd")
    return gx


@spmm_csr_max_bwd.register_fake
def _(rowptr, col, val, x, grad_out):
    return x.new_empty(x.shape, dtype=torch.float32)


def _spmm_setup(ctx, inputs, output):
    rowptr, col, val, x, reduce = inputs
    ctx.save_for_backward(rowptr, col, val if val is not None else rowptr.new_empty(0),
                          x if reduce == 2 else rowptr.new_empty(0))
    ctx.has_val, ctx.reduce, ctx.n_cols = val is not None, reduce, x.shape[0]


def _spmm_backward(ctx, g):
    rowptr, col, val, x = ctx.saved_tensors
    if ctx.reduce == 2:
        gx = torch.ops.ocn.spmm_csr_max_bwd(rowptr, col, val if ctx.has_val else None, x, g)
        return None, None, None, gx, None
    gx = torch.ops.ocn.spmm_csr_bwd(rowptr, col, val if ctx.has_val else None, g, ctx.n_cols, ctx.reduce)
    return None, None, None, gx, None


spmm_csr.register_autograd(_spmm_backward, setup_context=_spmm_setup)


@torch.library.custom_op("ocn::gcn_spmm", mutates_args=(), device_types="cuda")
def gcn_spmm(rowptr: Tensor, col: Tensor, norm: Tensor, x: Tensor, mode: int, edge_w: Optional[Tensor] = None,
             self_adjoint: bool = True) -> Tensor:
    """``self_adjoint``: the matrix is a symmetric unit-weight adjacency (what to_symmetric() yields), so its
    backward is the same product; pass False for a directed graph or with ``edge_w`` (DropAdj values)."""
    return _sp._gcn_raw(_g(rowptr, col), edge_w, norm, mode, x.float())


@gcn_spmm.register_fake
def _(rowptr, col, norm, x, mode, edge_w=None, self_adjoint=True):
    return x.new_empty(rowptr.numel() - 1, x.shape[1], dtype=torch.float32)


def _gcn_setup(ctx, inputs, output):
    rowptr, col, norm, x, mode, edge_w, self_adjoint = inputs
    ctx.save_for_backward(rowptr, col, norm, edge_w if edge_w is not None else rowptr.new_empty(0))
    ctx.mode, ctx.has_w, ctx.self_adjoint = mode, edge_w is not None, bool(self_adjoint) and edge_w is None


def _gcn_backward(ctx, g):
    rowptr, col, norm, edge_w = ctx.saved_tensors
    if ctx.self_adjoint:  # A-hat == A-hat^T: grad_x = A-hat grad_out
        return None, None, None, torch.ops.ocn.gcn_spmm(rowptr, col, norm, g.contiguous(), ctx.mode), None, None, None
    gx = _sp._gcn_transpose_raw(_g(rowptr, col), edge_w if ctx.has_w else None, norm, ctx.mode, g)
    return None, None, None, gx, None, None, None


gcn_spmm.register_autograd(_gcn_backward, setup_context=_gcn_setup)


@torch.library.custom_op("ocn::spgemm_a2", mutates_args=(), device_types="cuda")
def spgemm_a2(rowptr: Tensor, col: Tensor, fold: int, with_value: bool) -> Tuple[Tensor, Tensor, Tensor]:
    out = _sp.spgemm_a2(_g(rowptr, col), fold, with_value)
    val = out.value if out.value is not None else torch.empty(0, dtype=torch.float32, device=col.device)
    return out.rowptr, out.col, val


@spgemm_a2.register_fake
def _(rowptr, col, fold, with_value):
    nnz = torch.library.get_ctx().new_dynamic_size()
    return rowptr.new_empty(rowptr.numel()), col.new_empty(nnz), rowptr.new_empty(nnz, dtype=torch.float32)


@torch.library.custom_op("ocn::cn_aggregate", mutates_args=(), device_types="cuda")
def cn_aggregate(rowptr: Tensor, col: Tensor, edges: Tensor, x: Tensor, ip3: Tensor, batch_size: int, order: int,
                 weighted: bool, variant: int, fill: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Inference-mode fused path over a stream of link batches (see cn.cn_aggregate_eval)."""
    xcn1, xcn2, xcn3, xij = _cn.cn_aggregate_eval(_g(rowptr, col), edges, x, batch_size, order, weighted, variant, fill, ip3)
    z = lambda t: t if t is not None else x.new_zeros(0, x.shape[1])
    return xcn1, z(xcn2), z(xcn3), xij


@cn_aggregate.register_fake
def _(rowptr, col, edges, x, ip3, batch_size, order, weighted, variant, fill):
    T, F = edges.shape[1], x.shape[1]
    e = lambda n: x.new_empty(n, F, dtype=torch.float32)
    return e(T), e(T if order >= 2 else 0), e(T if (order >= 3 and variant == 5) else 0), e(T)
