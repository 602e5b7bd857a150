"""The two routes of INTEGRATION.md at citation2 shape, one 2048-link batch of the evaluation stream (order 2, as the
reference's own get_cn1_cn2): (a) the shim route -- the call sequence of get_cn1_cn2's text (NeighborOverlapCitation2.py:78-104)
on the pygho / torch_sparse stand-ins, then the API calls of cn5's multidomainforward that touch the sparse matrices;
(b) the fused route (CNSession + predictor mirror)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
import ocn_b200.shim as shim
from ocn_b200 import synth

shim.install(force=True)
import torch_sparse  # noqa: E402  (the stand-ins)
from pygho import SparseTensor as pSparseTensor  # noqa: E402
from pygho.backend.Spspmm import spsphadamard, spspmm  # noqa: E402
from torch_sparse.matmul import spmm_add  # noqa: E402

dev = "cuda:0"
g = synth.make_graph("citation2", device=dev)
G = ob.Graph(g.rowptr, g.col, g.n)
x = g.features(32, device=dev)
row = torch.repeat_interleave(torch.arange(g.n, device=dev), g.rowptr[1:] - g.rowptr[:-1])
padj = pSparseTensor(torch.stack((row, g.col.long())), torch.ones(row.numel(), device=dev), (g.n, g.n), is_coalesced=True)
e = g.query_edges(2048, "stream", device=dev)


def shim_route():
    Ei = padj.index_select([0], e[0].unsqueeze(0))
    Ej = padj.index_select([0], e[1].unsqueeze(0))
    cn1 = spsphadamard(Ei, Ej)
    cn2 = spsphadamard(Ei, spspmm(Ej, 1, padj, 0))
    c1, c2 = cn1.to_torch_sparse_coo(), cn2.to_torch_sparse_coo()
    r1, k1 = c1.indices()
    r2, k2 = c2.indices()
    t1 = torch_sparse.SparseTensor(row=r1, col=k1, value=c1.values(), sparse_sizes=(2048, g.n))
    t2 = torch_sparse.SparseTensor(row=r2, col=k2, value=c2.values(), sparse_sizes=(2048, g.n))
    col_sum = t1.sum(dim=0)
    col_sum[col_sum == 0] = 1
    inv = 1 / col_sum
    inv[col_sum == 1] = 0
    n1 = t1.mul(inv.view(1, -1))
    return spmm_add(n1, x), spmm_add(t2, x)


ip3 = torch.zeros(3, device=dev)


def fused_route():
    s = ob.CNSession(G, e, 2048, 2).build(2, True)
    s.stats(5, 0.0, ip3, 0)
    out = s.aggregate(x, 5, 0.0, ip3)
    s.release()
    return out


for name, fn in (("shim route (get_cn1_cn2 text + the sparse calls of cn5)", shim_route), ("fused route (session)", fused_route)):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 10:.3f} ms per 2048-link batch", flush=True)
a_, _ = shim_route()
b_ = fused_route()[0]
print("xcn1 of the two routes agree:", bool(torch.allclose(a_, b_, rtol=1e-5, atol=1e-5)))
