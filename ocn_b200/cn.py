"""Host side of the fused common-neighbour path (pieces 1, 1b and 2 of north_star).

Mirrors the reference call sites one-to-one:

* ``adjoverlap(adj1, adj2, tarei)``            utils.py:248-285 (calresadj=False branch)
* ``get_cn1_cn2(adj, tedge)``                  NeighborOverlapCitation2.py:78-104
* ``cn_aggregate(...)``                        the part of ``multidomainforward`` between the CN sets
                                               and the MLP heads (model.py:2261-2429 cn5, :3114-3216 cn7,
                                               :2546-2940 order 3)

Everything runs through the C ABI of ``include/ocn_b200.h`` on CUDA tensors; nothing here
computes on the CPU.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .graph import Graph, _require_cuda

PLAN_WORDS = 24
PLAN_HUB_DEGREE = 8  # include/ocn_b200.h OCN_PLAN_HUB_DEGREE
PLAN_BAD_LINKS = 17  # OCN_PLAN_BAD_LINKS
PLAN_DENSE = 18      # OCN_PLAN_DENSE
COLSTAT_BUDGET_BYTES = 4 << 30  # per-wave cap for the per-batch column statistics
HUB_WORKSPACE_FLOOR = 512 << 20  # first size of the per-stream scratch of the indexed order-3 path
RECORDS_BUCKET = 16 << 20        # record buffers are sized in these steps, so that the caching allocator re-serves them


@dataclasses.dataclass
class SparseRows:
    """A [B x N] sparse matrix in CSR, entries ascending by (row, col) -- the layout of the
    torch_sparse.SparseTensor that ``adjoverlap`` returns (utils.py:146-151)."""
    rowptr: Tensor          # int64 [B+1]
    col: Tensor             # int64 [nnz]
    value: Optional[Tensor]  # fp32 [nnz]
    shape: Tuple[int, int]

    def row(self) -> Tensor:
        return torch.repeat_interleave(torch.arange(self.shape[0], device=self.col.device),
                                       self.rowptr[1:] - self.rowptr[:-1])

    def coo(self):
        return self.row(), self.col, self.value

    def sizes(self):
        return self.shape

    def nnz(self) -> int:
        return int(self.col.numel())


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _edges(tarei: Tensor) -> Tuple[Tensor, Tensor]:
    if tarei.dim() != 2 or tarei.shape[0] != 2:
        raise ValueError("target links must be an int64 tensor of shape [2, B]")
    e = tarei.to(torch.int64).contiguous()
    return e[0], e[1]


def adjoverlap(adj1: Graph, adj2: Graph, tarei: Tensor, filled1: bool = False, calresadj: bool = False,
               cnsampledeg: int = -1, ressampledeg: int = -1):
    """``utils.adjoverlap``: row b = adj1[tarei[0,b]] (cap) adj2[tarei[1,b]], value 1.0.

    ``adj2`` may be any matrix over the same columns (A itself, an explicit A^2 from
    ``spgemm_a2`` or the folded adj2byblock matrix)."""
    if cnsampledeg > 0 or ressampledeg > 0:
        raise NotImplementedError("the random samplers (sparsesample_reweight, utils.py:109-143) are not on the hot path: "
                                  "no README command sets cndeg > 0")
    _require_cuda(tarei)
    if adj1.sizes() != adj2.sizes():
        raise AssertionError("adj1.sizes() == adj2.sizes()")  # utils.py:165
    src, dst = _edges(tarei)
    overlap = _rows_setop("intersect", adj1, adj2, src, dst)
    if not calresadj:
        return overlap
    # utils.py:260-274: (overlap, adj1[src] minus adj2[dst], adj2[dst] minus adj1[src])
    return overlap, _rows_setop("difference", adj1, adj2, src, dst), _rows_setop("difference", adj2, adj1, dst, src)


def _rows_setop(op: str, adj1: Graph, adj2: Graph, src: Tensor, dst: Tensor) -> SparseRows:
    B = src.numel()
    dev = adj1.device
    counts = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    L = _lib.lib()
    count_fn, fill_fn = getattr(L, f"ocn_rows_{op}_count"), getattr(L, f"ocn_rows_{op}_fill")
    with torch.cuda.device(dev):
        st = _stream(dev)
        _lib.check(count_fn(_lib.ptr(adj1.rowptr), _lib.ptr(adj1.col), adj1.n, _lib.ptr(adj2.rowptr), _lib.ptr(adj2.col),
                            adj2.n, _lib.ptr(src), _lib.ptr(dst), B, _lib.ptr(counts), st), f"ocn_rows_{op}_count")
        rowptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts[:B], 0, out=rowptr[1:])
        nnz, bad = torch.stack((rowptr[-1], counts[B])).tolist()  # one read-back: output size + out-of-range endpoints
        if bad:
            raise IndexError(f"{bad} target links have an endpoint outside the matrices' rows "
                             f"([0, {adj1.n}) for the first, [0, {adj2.n}) for the second end)")
        col = torch.empty(nnz, dtype=torch.int64, device=dev)
        if nnz:
            _lib.check(fill_fn(_lib.ptr(adj1.rowptr), _lib.ptr(adj1.col), adj1.n, _lib.ptr(adj2.rowptr), _lib.ptr(adj2.col),
                               adj2.n, _lib.ptr(src), _lib.ptr(dst), B, _lib.ptr(rowptr), _lib.ptr(col), st),
                       f"ocn_rows_{op}_fill")
    return SparseRows(rowptr, col, torch.ones(nnz, dtype=torch.float32, device=dev), (B, adj1.n_cols))


class CNSession:
    """One stream of target links against one graph: plan -> build -> stats -> aggregate / extract.

    ``batch_size`` cuts the stream into the reference's link batches; every batch is
    normalised independently (the column sums of model.py:2261 run over one batch)."""

    def __init__(self, graph: Graph, tarei: Tensor, batch_size: Optional[int] = None, order: int = 3,
                 hub_degree: int = 0, plan_stream: Optional["torch.cuda.Stream"] = None):
        """``hub_degree``: rows of at least this many columns are walked once per stream by the indexed
        order-3 build (0 = chosen from n and the stream length, -1 = per-run tables instead).

        ``plan_stream``: run the plan and its one size read-back on this CUDA stream instead of the
        current one, so the host only waits for the plan while the current stream keeps executing the
        previous session (``tarei`` must be ready on that stream).  Everything after the plan runs on
        the current stream, which is made to wait for the plan."""
        _require_cuda(graph.col)
        _require_cuda(tarei)
        self.g = graph
        if tarei.dim() != 2 or tarei.shape[0] != 2:
            raise ValueError("target links must be an int64 tensor of shape [2, B]")
        self.T = int(tarei.shape[1])
        if self.T == 0:
            raise ValueError("empty link batch")
        self.batch_size = int(batch_size or self.T)
        self.nb = (self.T + self.batch_size - 1) // self.batch_size
        self.dev = graph.device
        L = _lib.lib()
        self.L = L
        self.plan_bytes = L.ocn_cn_plan_bytes(self.T)
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev), torch.cuda.stream(plan_stream if plan_stream is not None else main):
            self.src, self.dst = _edges(tarei)  # may copy: on the stream the plan reads it from
            self.plan_scratch = torch.empty(self.plan_bytes, dtype=torch.uint8, device=self.dev)
            self.plan = torch.zeros(PLAN_WORDS, dtype=torch.int64, device=self.dev)
            _lib.check(L.ocn_cn_plan(_lib.ptr(graph.rowptr), _lib.ptr(graph.col), graph.n, _lib.ptr(self.src),
                                     _lib.ptr(self.dst), self.T, self.batch_size, int(order), int(hub_degree),
                                     _lib.ptr(self.plan_scratch), self.plan_bytes,
                                     _lib.ptr(self.plan), _stream(self.dev)), "ocn_cn_plan")
            host = self.plan.tolist()  # the one host sync of the session (of the plan stream only): buffer sizes
        if plan_stream is not None:
            main.wait_stream(plan_stream)
            for t in (self.plan_scratch, self.plan, self.src, self.dst):
                t.record_stream(main)
        if host[PLAN_BAD_LINKS]:
            raise IndexError(f"{host[PLAN_BAD_LINKS]} target links have an endpoint outside [0, {graph.n})")
        self.num_records, self.num_runs, self.num_units = host[0], host[1], host[2]
        self.plan_host = (ctypes.c_int64 * PLAN_WORDS)(*host)
        self.hub_degree = host[PLAN_HUB_DEGREE]
        self.dense = bool(host[PLAN_DENSE])
        self.hub_bytes = L.ocn_cn_hub_bytes(graph.n, graph.nnz, self.plan_host) if (self.hub_degree > 0 or self.dense) else 0
        rec_bytes = max(1, self.num_records) * L.ocn_cn_record_bytes()
        self.records = torch.empty(-(-rec_bytes // RECORDS_BUCKET) * RECORDS_BUCKET, dtype=torch.uint8, device=self.dev)
        self.colstat = None  # borrowed by build(with_stats=True)
        self.bscal = torch.zeros(self.nb * 8, dtype=torch.float32, device=self.dev)
        self.order = 0
        self.plan_order = int(order)
        self.weighted = True
        self._released = False

    # -- steps ------------------------------------------------------------------------------
    def build(self, order: int, weighted: bool, with_stats: bool = True, spd: bool = False) -> "CNSession":
        """``spd``: the shortest-path variant of SPD.py:65-126 -- 2-walk counts of nodes adjacent to the destination
        are zeroed before the column statistics are taken (weighted sets only)."""
        g = self.g
        if spd and not weighted:
            raise ValueError("the shortest-path variant masks walk counts: it needs weighted=True (SPD.py builds pygho matrices)")
        if order > self.plan_order:
            raise ValueError(f"the session was planned for order {self.plan_order}; cannot build order {order}")
        if with_stats and self.colstat is not None and not self._released and self.order > 0:
            # a second build (another order or weighting) would add its counts on top of the first one's: hand the
            # statistics back (release re-zeroes exactly the touched entries) and start from a zeroed buffer
            self.release()
        if with_stats and (self.colstat is None or self._released):
            self.colstat = _borrow_colstat(g, self.nb * self.L.ocn_cn_colstat_bytes(g.n))
            self._released = False
        hub_scratch = node_scratch = None
        if self.hub_bytes > 0 and (order >= 3 or self.dense):
            hub_scratch, node_scratch = _hub_workspace(g, self.hub_bytes)
        with torch.cuda.device(self.dev):
            try:
                _lib.check(self.L.ocn_cn_build(_lib.ptr(g.rowptr), _lib.ptr(g.col), g.n, _lib.ptr(self.src),
                                               _lib.ptr(self.dst), self.T, self.batch_size, int(order),
                                               2 if spd else int(bool(weighted)),
                                               _lib.ptr(self.plan_scratch), _lib.ptr(self.plan), _lib.ptr(self.records),
                                               self.num_records, _lib.ptr(self.colstat) if with_stats else None,
                                               g.nnz, self.plan_host, _lib.ptr(hub_scratch), self.hub_bytes,
                                               _lib.ptr(node_scratch), _stream(self.dev)), "ocn_cn_build")
            except _lib.OcnError:
                # the zero-on-entry scratch of a failed build is not trusted again: drop it (the next session on
                # this stream allocates fresh, zeroed buffers) together with the borrowed column statistics
                _drop_workspace(g, node_scratch)
                self.colstat = None
                raise
        self.order, self.weighted = int(order), bool(weighted)
        return self

    def stats(self, variant: int, fill: float, ip: Tensor, stage: int = 0) -> Tensor:
        g = self.g
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ocn_cn_stats(_lib.ptr(g.rowptr), _lib.ptr(g.col), g.n, _lib.ptr(self.src), self.T,
                                           self.batch_size, self.order, int(self.weighted), int(variant), float(fill),
                                           _lib.ptr(ip), int(stage), _lib.ptr(self.plan_scratch),
                                           _lib.ptr(self.records), _lib.ptr(self.colstat), _lib.ptr(self.bscal),
                                           self.plan_host, _stream(self.dev)), "ocn_cn_stats")
        return self.bscal.view(self.nb, 8)

    def set_batch_ip(self, ips: Tensor) -> None:
        """Per-batch inner-product coefficients (``[num_batches]`` or ``[num_batches, 3]``) into the batches' scalar
        slots 5..7; ``aggregate`` / ``aggregate_bwd`` / ``extract`` called with ``ip=None`` read them from there."""
        v = ips.detach().float().reshape(self.nb, -1)
        if v.shape[1] == 1:
            v = v.expand(self.nb, 3)
        self.bscal.view(self.nb, 8)[:, 5:8] = v

    def aggregate(self, x: Tensor, variant: int, fill: float, ip: Optional[Tensor], want_xij: bool = True):
        g = self.g
        x = _check_x(x, g)
        F = x.shape[1]
        mk = lambda: torch.empty(self.T, F, dtype=torch.float32, device=self.dev)
        xcn1, xcn2 = mk(), (mk() if self.order >= 2 else None)
        xcn3 = mk() if (self.order >= 3 and variant == 5) else None
        xij = mk() if want_xij else None
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ocn_cn_aggregate(
                _lib.ptr(g.rowptr), _lib.ptr(g.col), g.n, _lib.ptr(self.src), _lib.ptr(self.dst), self.T,
                self.batch_size, self.order, int(self.weighted), int(variant), float(fill), _lib.ptr(ip),
                _lib.ptr(self.plan_scratch), _lib.ptr(self.records), _lib.ptr(self.colstat), _lib.ptr(self.bscal),
                _lib.ptr(x), F, _lib.ptr(xcn1), _lib.ptr(xcn2), _lib.ptr(xcn3), _lib.ptr(xij), self.plan_host,
                _stream(self.dev)), "ocn_cn_aggregate")
        return xcn1, xcn2, xcn3, xij

    def aggregate_bwd(self, x: Tensor, variant: int, fill: float, ip: Tensor, g1, g2, g3, gij, grad_x: Tensor):
        if self._released:
            raise RuntimeError("the column statistics of this session were released (a backward through it has run, or "
                               "release() was called): the weights cannot be rebuilt -- build the session again, as "
                               "autograd asks for retain_graph on a second backward")
        g = self.g
        c = lambda t: None if t is None else t.contiguous()
        g1, g2, g3, gij = c(g1), c(g2), c(g3), c(gij)
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ocn_cn_aggregate_bwd(
                _lib.ptr(g.rowptr), _lib.ptr(g.col), g.n, _lib.ptr(self.src), _lib.ptr(self.dst), self.T,
                self.batch_size, self.order, int(self.weighted), int(variant), float(fill), _lib.ptr(ip),
                _lib.ptr(self.plan_scratch), _lib.ptr(self.records), _lib.ptr(self.colstat), _lib.ptr(self.bscal),
                _lib.ptr(x), x.shape[1], _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(g3), _lib.ptr(gij), _lib.ptr(grad_x),
                _stream(self.dev)), "ocn_cn_aggregate_bwd")
        return grad_x

    def extract(self, which: int, variant: int = 5, fill: float = 0.0, ip: Optional[Tensor] = None) -> SparseRows:
        """which 1/2/3: raw CN_k; 11/12/13: the normalised matrices the predictor feeds to spmm_add."""
        g = self.g
        counts = torch.zeros(self.T + 1, dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            st = _stream(self.dev)
            _lib.check(self.L.ocn_cn_extract_count(_lib.ptr(g.rowptr), g.n, _lib.ptr(self.src), self.T, int(which),
                                                   int(self.weighted), _lib.ptr(self.plan_scratch),
                                                   _lib.ptr(self.records), _lib.ptr(counts), st),
                       "ocn_cn_extract_count")
            rowptr = torch.zeros(self.T + 1, dtype=torch.int64, device=self.dev)
            torch.cumsum(counts[:self.T], 0, out=rowptr[1:])
            nnz = int(rowptr[-1].item())
            col = torch.empty(nnz, dtype=torch.int64, device=self.dev)
            val = torch.empty(nnz, dtype=torch.float32, device=self.dev)
            if ip is None:
                ip = torch.zeros(3, dtype=torch.float32, device=self.dev)
            if nnz:
                _lib.check(self.L.ocn_cn_extract_fill(
                    _lib.ptr(g.rowptr), _lib.ptr(g.col), g.n, _lib.ptr(self.src), self.T, self.batch_size, int(which),
                    int(self.weighted), int(variant), float(fill), _lib.ptr(ip), _lib.ptr(self.plan_scratch),
                    _lib.ptr(self.records), _lib.ptr(self.colstat), _lib.ptr(self.bscal), _lib.ptr(rowptr),
                    _lib.ptr(col), _lib.ptr(val), st), "ocn_cn_extract_fill")
        return SparseRows(rowptr, col, val, (self.T, g.n))

    def release(self):
        """Zero the column statistics this stream touched (the buffer is reusable afterwards)."""
        g = self.g
        if self.colstat is None or self._released:
            return
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ocn_cn_release(_lib.ptr(g.rowptr), _lib.ptr(g.col), g.n, _lib.ptr(self.src), self.T,
                                             self.batch_size, _lib.ptr(self.plan_scratch), _lib.ptr(self.records),
                                             _lib.ptr(self.colstat), self.plan_host, _stream(self.dev)), "ocn_cn_release")
        self._released = True
        _return_colstat(self.g, self.colstat)  # (the reference is kept for inspection; _released guards its further use)


def reserve_stream_pool(nbytes: int = 4 << 30, device=None) -> None:
    """Seed torch's caching allocator for the CURRENT stream with one block of ``nbytes``: the per-session buffers
    (plan scratch, records, outputs; a few tens of MB each, sizes varying from call to call) are then carved out of
    it instead of reaching cudaMalloc, whose latency in the middle of a stream of sessions was measured at 1 - 40 ms.
    Call once per stream before a latency-sensitive loop."""
    dev = torch.device(device if device is not None else torch.cuda.current_device())
    block = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    # the allocator serves requests of up to 1 MB (plan words, batch scalars, counts) from a separate pool of 2 MB
    # segments: seed that one too
    small = [torch.empty(1 << 20, dtype=torch.uint8, device=dev) for _ in range(64)]
    del block, small


def _hub_workspace(g: Graph, nbytes: int):
    """Scratch of the hub stage, kept with the graph: a byte buffer that only grows and the 32-byte-per-node
    key index (all zero between calls; ocn_cn_build restores it).  Calls on one stream are ordered by the stream;
    every stream has its own pair of buffers, so that two sessions may be in flight on two streams."""
    sid = torch.cuda.current_stream(g.device).cuda_stream  # one workspace per stream: sessions on different streams overlap
    buf = g._ws.get(("hub", sid))
    if buf is None or buf.numel() < nbytes:
        # grow rarely: a cudaMalloc in the middle of a stream of sessions costs 1 - 40 ms (measured), the scratch of a
        # 65 536-link call is ~40 MB
        g._ws.pop(("hub", sid), None)
        buf = torch.empty(max(2 * int(nbytes), HUB_WORKSPACE_FLOOR), dtype=torch.uint8, device=g.device)
        g._ws[("hub", sid)] = buf
    node = g._ws.get(("hub_node", sid))
    if node is None:
        node = torch.zeros((4 * g.n, 4), dtype=torch.int32, device=g.device)  # 32 B per node; second half: the pass over heavy sources
        g._ws[("hub_node", sid)] = node
    return buf, node


def _drop_workspace(g: Graph, node_scratch) -> None:
    sid = torch.cuda.current_stream(g.device).cuda_stream
    g._ws.pop(("hub", sid), None)
    g._ws.pop(("hub_node", sid), None)
    if node_scratch is not None:
        _lib.lib().ocn_cn_hub_scratch_reset(_lib.ptr(node_scratch))  # the address may be handed out again, zeroed


def _borrow_colstat(g: Graph, nbytes: int) -> Tensor:
    """A zeroed buffer for the per-batch column statistics.  Buffers handed back by
    ``CNSession.release`` (which re-zeroes exactly the touched entries) are reused, so a steady
    stream of sessions does not pay a multi-GB memset per call."""
    pool = g._ws.setdefault(("colstat", torch.cuda.current_stream(g.device).cuda_stream), [])
    for k, buf in enumerate(pool):
        if buf.numel() >= nbytes:
            pool.pop(k)
            return buf
    return torch.zeros(nbytes, dtype=torch.uint8, device=g.device)


def _return_colstat(g: Graph, buf: Tensor):
    # the pool is per stream: the re-zeroing enqueued by release() is ordered before the next borrower's kernels
    pool = g._ws.setdefault(("colstat", torch.cuda.current_stream(g.device).cuda_stream), [])
    if len(pool) < 2:
        pool.append(buf)


def _check_x(x: Tensor, g: Graph) -> Tensor:
    _require_cuda(x)
    if x.dim() != 2 or x.shape[0] < g.n:
        raise ValueError(f"x must be [N, F] with N >= {g.n}")
    if x.dtype != torch.float32:
        raise TypeError("x must be float32 (the reference computes the CN aggregation in fp32)")
    return x.contiguous()


def get_cn(adj: Graph, tedge: Tensor, order: int = 2, weighted: bool = True, hub_degree: int = 0,
           batch_size: Optional[int] = None) -> List[SparseRows]:
    """``get_cn1_cn2`` (NeighborOverlapCitation2.py:78-104) generalised to ``order`` sets.
    weighted=True keeps the pygho walk counts as values, False gives the 0/1 structure that
    ``adjoverlap(adj, adj2, e)`` yields in the _large drivers."""
    s = CNSession(adj, tedge, batch_size, order, hub_degree).build(order, weighted, with_stats=False)
    return [s.extract(k) for k in range(1, order + 1)]


def get_cn1_cn2(adj: Graph, tedge: Tensor):
    cn1, cn2 = get_cn(adj, tedge, 2, True)
    return cn1, cn2


def waves(num_links: int, batch_size: int, n_nodes: int, budget_bytes: int = COLSTAT_BUDGET_BYTES):
    """Split a long stream into waves whose per-batch column statistics fit ``budget_bytes``."""
    per_batch = 32 * n_nodes
    bpw = max(1, budget_bytes // per_batch)
    step = bpw * batch_size
    return [(s, min(num_links, s + step)) for s in range(0, num_links, step)]


def cn_aggregate_eval(graph: Graph, tarei: Tensor, x: Tensor, batch_size: int, order: int, weighted: bool,
                      variant: int, fill: float, ip: Tensor, want_xij: bool = True,
                      budget_bytes: int = COLSTAT_BUDGET_BYTES):
    """Inference-mode fused path for a whole stream (no autograd): returns xcn1, xcn2, xcn3, xij
    of shape [T, F].  ``ip`` is the predictor's ``innerprod`` buffer broadcast to 3 entries."""
    T = tarei.shape[1]
    outs = None
    for (s, e) in waves(T, batch_size, graph.n, budget_bytes):
        sess = CNSession(graph, tarei[:, s:e], batch_size, order).build(order, weighted)
        if variant == 5:
            sess.stats(variant, fill, ip, 0)
        part = sess.aggregate(x, variant, fill, ip, want_xij)
        sess.release()
        if s == 0 and e == T:
            return part
        if outs is None:
            outs = [None if p is None else torch.empty(T, p.shape[1], dtype=p.dtype, device=p.device) for p in part]
        for o, p in zip(outs, part):
            if o is not None:
                o[s:e] = p
    return tuple(outs)
