"""``pygho.backend.Spspmm``: ``spsphadamard`` and ``spspmm`` (NeighborOverlapCitation2.py:82-85, model.py:2243,
3153-3196) on libocn_b200.

``spsphadamard`` recognises the two shapes ``get_cn1_cn2`` produces and sends them to the fused kernels:

* ``adj[i] (.) adj[j]``                        -> ``ocn_rows_intersect_*`` on the graph (CN1)
* ``adj[i] (.) (adj[j] . adj [. adj])``        -> ``ocn_cn_plan / build / extract`` (CN2 / CN3 with walk counts;
  needs a symmetric unit-valued ``adj``, which is what the drivers build)

Everything else is the general case on two explicit matrices: ``ocn_rows_hadamard_*``.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import Product, RowSelect, SparseTensor, _need_cuda
from .... import _lib
from .... import cn as _cn


FUSED = {"cn1": 0, "cn_order": 0, "explicit": 0}   # which path served spsphadamard (read by the tests)


def _wrap_rows(rows, shape, dtype) -> SparseTensor:
    """SparseRows (rowptr, col int64, value fp32) -> pygho SparseTensor with CSR cached."""
    B = shape[0]
    rp = rows.rowptr
    row = torch.repeat_interleave(torch.arange(B, device=rp.device), rp[1:] - rp[:-1], output_size=int(rows.col.numel()))
    out = SparseTensor(torch.stack((row, rows.col)), rows.value.to(dtype), shape, is_coalesced=True)
    out._rowptr = rp
    return out


def _fusable(adj: SparseTensor, need_symmetric: bool) -> bool:
    if adj.shape[0] != adj.shape[1] or not adj.indices.is_cuda or not adj.unit_valued():
        return False
    return adj.graph().is_symmetric() if need_symmetric else True


def spsphadamard(A: SparseTensor, B: SparseTensor, *_, **__) -> SparseTensor:
    if A.shape[:2] != B.shape[:2]:
        raise AssertionError(f"spsphadamard: shapes differ ({A.shape} vs {B.shape})")
    dtype = torch.get_default_dtype()
    if isinstance(A, RowSelect):   # (materialised or not: it still knows its parent and rows)
        adj = A.parent
        vdt = adj.values.dtype if adj.values is not None else dtype
        if isinstance(B, RowSelect) and B.parent is adj and _fusable(adj, False):
            g = adj.graph()
            FUSED["cn1"] += 1
            if A.idx.numel() == 0:
                return _empty(A.shape, vdt, A.idx.device)
            rows = _cn.adjoverlap(g, g, torch.stack((A.idx, B.idx)))
            return _wrap_rows(rows, A.shape, vdt)
        if isinstance(B, Product):
            ch = B.chain()
            if ch is not None and ch[1] is adj and ch[2] <= 2 and _fusable(adj, True):
                sel, _, k = ch
                order = k + 1
                FUSED["cn_order"] += 1
                if A.idx.numel() == 0:
                    return _empty(A.shape, vdt, A.idx.device)
                sess = _cn.CNSession(adj.graph(), torch.stack((A.idx, sel.idx)), None, order)
                sess.build(order, True, with_stats=False)
                return _wrap_rows(sess.extract(order), A.shape, vdt)
    return _hadamard_explicit(A, B)


def _empty(shape, dtype, dev) -> SparseTensor:
    return SparseTensor(torch.zeros(2, 0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=dtype, device=dev), shape,
                        is_coalesced=True)


def _hadamard_explicit(A: SparseTensor, B: SparseTensor) -> SparseTensor:
    _need_cuda(A.indices, "spsphadamard")
    FUSED["explicit"] += 1
    dev = A.indices.device
    rows = A.shape[0]
    rpa, ca = A._csr()
    rpb, cb = B._csr()
    va, vb = A._fvalues(), B._fvalues()
    L = _lib.lib()
    counts = torch.zeros(rows, dtype=torch.int64, device=dev)
    rowptr = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.ocn_rows_hadamard_count(_lib.ptr(rpa), _lib.ptr(ca), _lib.ptr(rpb), _lib.ptr(cb), rows,
                                             _lib.ptr(counts), st), "ocn_rows_hadamard_count")
        torch.cumsum(counts, 0, out=rowptr[1:])
        nnz = int(rowptr[-1].item())
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        if nnz:
            _lib.check(L.ocn_rows_hadamard_fill(_lib.ptr(rpa), _lib.ptr(ca), _lib.ptr(va), _lib.ptr(rpb), _lib.ptr(cb),
                                                _lib.ptr(vb), rows, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), st),
                       "ocn_rows_hadamard_fill")
    row = torch.repeat_interleave(torch.arange(rows, device=dev), counts, output_size=nnz)
    dt = A.values.dtype if A.values is not None else (B.values.dtype if B.values is not None else torch.get_default_dtype())
    out = SparseTensor(torch.stack((row, col.to(torch.int64))), val.to(dt), A.shape, is_coalesced=True)
    out._rowptr, out._col32 = rowptr, col
    return out


def spspmm(A: SparseTensor, dim1: int, B: SparseTensor, dim2: int, aggr: str = "sum", *_, **__) -> SparseTensor:
    if dim1 != 1 or dim2 != 0 or aggr != "sum":
        raise NotImplementedError("the reference only calls spspmm(A, 1, B, 0) with the default sum")
    if A.shape[1] != B.shape[0]:
        raise AssertionError(f"spspmm: inner sizes differ ({A.shape} . {B.shape})")
    return Product(A, B)


def _is_diagonal(B: SparseTensor) -> bool:
    ind = B.indices
    return B.shape[0] == B.shape[1] and ind.shape[1] == B.shape[0] and bool((ind[0] == ind[1]).all()) and \
        bool((ind[0] == torch.arange(B.shape[0], device=ind.device)).all())


def _spspmm_explicit(A: SparseTensor, B: SparseTensor) -> SparseTensor:
    """General ``A . B`` with summed duplicates, sorted output.  ``B`` diagonal (the polynomial bases of cn7,
    model.py:3001-3016, 3153) scales the columns of ``A`` and keeps its pattern; otherwise the products are expanded,
    sorted and reduced with torch CUDA ops (off the hot path: the fused kernels cover the driver's products)."""
    _need_cuda(A.indices, "spspmm")
    ai, av = A.indices, A.values
    if _is_diagonal(B):
        d = B.values if B.values is not None else torch.ones(B.shape[0], device=ai.device)
        v = (av if av is not None else torch.ones(ai.shape[1], device=ai.device)) * d[ai[1]]
        out = SparseTensor(ai, v, (A.shape[0], B.shape[1]), is_coalesced=True)
        out._rowptr, out._col32 = A._rowptr, A._col32
        return out
    rpb, _ = B._csr()
    bi, bv = B.indices, B.values
    k = ai[1]
    start, cnt = rpb[k], rpb[k + 1] - rpb[k]
    total = int(cnt.sum())
    dev = ai.device
    owner = torch.repeat_interleave(torch.arange(ai.shape[1], device=dev), cnt, output_size=total)
    pos = torch.arange(total, device=dev) + torch.repeat_interleave(start - (torch.cumsum(cnt, 0) - cnt), cnt, output_size=total)
    w = B.shape[1]
    key = ai[0][owner] * w + bi[1][pos]
    one = torch.ones((), device=dev)
    val = (av[owner] if av is not None else one) * (bv[pos] if bv is not None else one)
    uk, inv = torch.unique(key, return_inverse=True)
    out = torch.zeros(uk.numel(), dtype=val.dtype, device=dev).index_add_(0, inv, val.expand(total) if val.dim() == 0 else val)
    return SparseTensor(torch.stack((torch.div(uk, w, rounding_mode="floor"), uk % w)), out, (A.shape[0], B.shape[1]),
                        is_coalesced=True)
