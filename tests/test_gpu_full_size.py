"""BASELINE.json configs 2-4 at FULL size through size-independent properties (the oracle would need hours):

* CN2 built by the fused session (no A^2 materialised) == adjoverlap(adj, A^2, e) on the explicit A^2 from the
  SpGEMM kernel -- two independent device paths (NeighborOverlap_large.py:74,78-79) -- and the weighted values
  equal the A^2 entries (the walk counts of get_cn1_cn2, NeighborOverlapCitation2.py:78-104);
* the column sums of the aggregates are linear in x (xcn(x + y) == xcn(x) + xcn(y) within fp32 rounding) and the
  structure-only CN1 aggregate with unit features counts |CN1|;
* the folded adj2byblock matrix at ddi shape has the structure the definition gives on a probe of rows.
"""
import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _a2_values_at(a2: ob.Graph, rows: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
    """a2[rows, cols] for entries known to exist (binary search inside the CSR rows on the device)."""
    key = rows * a2.n + cols
    a2key = a2.row() * a2.n + a2.col.long()
    pos = torch.searchsorted(a2key, key)
    assert bool((a2key[pos] == key).all()), "a CN2 entry is missing from the explicit A^2"
    return a2.value[pos]


@pytest.mark.parametrize("name", ["pubmed", "collab", "ddi"])
def test_fused_cn2_equals_explicit_a2_path(name):
    g = synth.make_graph(name, device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    assert G.validate() == 0
    B = g.batch
    e = g.query_edges(B, "mixed", device=DEV)
    a2 = ob.spgemm_a2(G, 0, True)                           # spadj @ spadj with its 2-walk counts
    exp1, exp2 = ob.adjoverlap(G, G, e), ob.adjoverlap(G, a2, e)
    cn1, cn2 = ob.get_cn(G, e, 2, weighted=True)
    assert torch.equal(cn1.rowptr, exp1.rowptr) and torch.equal(cn1.col, exp1.col)
    assert torch.equal(cn2.rowptr, exp2.rowptr) and torch.equal(cn2.col, exp2.col)
    rows = e[1][cn2.row()]
    assert torch.equal(cn2.value, _a2_values_at(a2, rows, cn2.col))
    # structure-only flavour of the _large drivers (values 1)
    s1, s2 = ob.get_cn(G, e, 2, weighted=False)
    assert torch.equal(s2.col, exp2.col) and bool((s2.value == 1).all())


@pytest.mark.parametrize("name,variant", [("pubmed", 7), ("collab", 5), ("ddi", 7)])
def test_aggregates_linear_and_counting(name, variant):
    g = synth.make_graph(name, device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    B, F = g.batch, g.hidden
    e = g.query_edges(B, "mixed", device=DEV)
    x, y = g.features(F, device=DEV), torch.randn(g.n, F, device=DEV, generator=torch.Generator(DEV).manual_seed(9))
    ip3 = torch.full((3,), 0.3, device=DEV)
    sess = ob.CNSession(G, e, None, 2).build(2, False)
    if variant == 5:
        sess.stats(5, 0.0, ip3, 0)
    fill = 1.0 if variant == 7 else 0.0
    ax, ay, axy = (sess.aggregate(t, variant, fill, ip3) for t in (x, y, x + y))
    for k in (0, 1):
        ref = ax[k] + ay[k]
        err = (axy[k] - ref).abs().max().item()
        assert err <= 1e-4 * (1 + ref.abs().max().item()), (k, err)
    if variant == 7:  # cn7 with --sum 1: singleton columns weigh 1, the others 1/c1 -> a unit feature sums the weights
        ones = torch.ones(g.n, 4, device=DEV)
        w = sess.aggregate(ones, 7, 1.0, ip3)[0][:, 0]
        n1 = sess.extract(11, 7, 1.0, ip3)
        ref = torch.zeros(B, device=DEV).index_add_(0, n1.row(), n1.value)
        assert torch.allclose(w, ref, rtol=1e-5, atol=1e-5)
    sess.release()


def test_folded_a2_definition_at_ddi_shape():
    """sparse_tensor_multiply(spadj, 1024) as written (utils.py:287-329): block (bi, bj) of the product is
    accumulated at rows/cols folded modulo the block grid; probe rows against a dense recomputation."""
    g = synth.make_graph("ddi", device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    folded = ob.sparse_tensor_multiply(G, 1024)
    true = ob.spgemm_a2(G, 0, True)
    assert folded.n == G.n and folded.nnz > 0 and true.nnz >= folded.nnz
    # folded[r', c'] = sum of A^2[r, c] over r = r' (mod 1024), c = c' (mod 1024), recomputed with torch on the device
    bs = 1024
    key = (true.row() % bs) * g.n + (true.col.long() % bs)
    ukey, inv = torch.unique(key, return_inverse=True)
    vals = torch.zeros(ukey.numel(), dtype=torch.float64, device=DEV).index_add_(0, inv, true.value.double())
    fkey = folded.row() * g.n + folded.col.long()
    assert torch.equal(fkey, ukey)
    assert torch.equal(folded.value.double(), vals)


def _bench_slice(g, k, T=65536):
    return g.query_edges((k + 1) * T, "stream", device=DEV)[:, k * T:(k + 1) * T].contiguous()


def test_bench_slices_indexed_equals_tables_at_full_size(lib_options):
    """The benchmarked workload itself (bench.py: citation2 shape, 65 536 links = 32 batches of the evaluation stream,
    order 3): on slice 0 and on the first slice that holds a hub source (> 1024 neighbours: heavy pass, CTA-wide
    counter windows) the indexed path -- folded 64-bit run sets (default), exact run sets, and the segment walker -- gives the records of the per-run
    table path (hub_degree = -1) bit for bit, and the aggregates xcn1..3 / xij of the two paths agree."""
    g = synth.make_graph("citation2", device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    deg = g.rowptr[1:] - g.rowptr[:-1]
    x = g.features(32, device=DEV)
    ip3 = torch.zeros(3, device=DEV)
    T = 65536
    hub_slice = None
    all_e = g.query_edges(40 * T, "stream", device=DEV)
    for k in range(1, 40):
        if int(deg[all_e[0, k * T:(k + 1) * T]].max()) > 1024:
            hub_slice = k
            break
    assert hub_slice is not None, "the synthetic stream is expected to hold a hub source within 40 slices"
    del all_e
    for k in (0, hub_slice):
        e = _bench_slice(g, k)
        off = ob.CNSession(G, e, 2048, 3, hub_degree=-1).build(3, True)
        assert off.hub_degree == 0
        off.stats(5, 0.0, ip3, 0)
        ref = off.aggregate(x, 5, 0.0, ip3)
        nb = off.num_records * 8
        for mode in ({}, {"hub_exact": 1}, {"hub_walker": 1}):   # folded run sets (default), exact sets, segment walker
            lib_options(hub_walker=0, hub_exact=0)
            lib_options(**mode)
            on = ob.CNSession(G, e, 2048, 3).build(3, True)
            assert on.hub_degree > 0 and on.num_records == off.num_records
            if k == hub_slice:
                assert on.plan_host[14] > 0, "expected a heavy-source pass on this slice"
            assert torch.equal(on.records[:nb], off.records[:nb]), (k, mode)
            on.stats(5, 0.0, ip3, 0)
            got = on.aggregate(x, 5, 0.0, ip3)
            for a, b in zip(got, ref):
                assert torch.equal(a, b), "same records, same kernels: the aggregates must be identical"
            on.release()
        off.release()
