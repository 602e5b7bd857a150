"""GPU parity tests: every CUDA entry point (through the C ABI) against the CPU oracle.

Integer / index results are compared bit-exactly; floating-point results within
|delta| <= RTOL * (1 + L1 mass of the sum), the tolerance SURVEY.md §8c-5 states, RTOL = 1e-5
(1e-4 where a near-cancelling column sum is divided by).
"""
import numpy as np
import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def _graph(g):
    return ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)


def _sp(g):
    return R.sp_from_csr(g.rowptr.cpu(), g.col.cpu())


def _assert_rows_equal(got: ob.SparseRows, ref: R.Sp, values=True, vtol=0.0):
    rp = ref.rowptr()
    assert torch.equal(got.rowptr.cpu(), rp), "row pointers differ"
    assert torch.equal(got.col.cpu(), ref.col), "column indices differ"
    if values:
        gv, rv = got.value.cpu(), ref.values()
        if vtol == 0.0:
            assert torch.equal(gv, rv), "values differ"
        else:
            assert torch.allclose(gv, rv, rtol=vtol, atol=vtol), f"values differ by {(gv - rv).abs().max()}"


def _close(got, ref, mass, rtol=RTOL):
    err = (got.cpu().double() - ref.double()).abs()
    bound = rtol * (1.0 + mass.double())
    assert bool((err <= bound).all()), f"max err {err.max().item():.3e}, worst bound ratio {(err / bound).max().item():.2f}"


GRAPHS = {
    "tiny": lambda: synth.tiny_graph(60, 260, 3),
    "tiny_dense": lambda: synth.tiny_graph(40, 900, 4),
    "cora": lambda: synth.make_graph("cora"),
    "pubmed": lambda: synth.make_graph("pubmed"),
    "collab_s": lambda: synth.make_graph("collab", scale=0.02),
    "citation2_s": lambda: synth.make_graph("citation2", scale=0.002),
    "ddi_s": lambda: synth.make_graph("ddi", scale=0.08),
}


def test_library_loaded_and_validate():
    g = synth.tiny_graph(50, 200, 1)
    G = _graph(g)
    assert G.validate() == 0
    bad = ob.Graph(g.rowptr.to(DEV), torch.flip(g.col, [0]).to(DEV), g.n)
    assert bad.validate() != 0
    # asymmetric: keep the raw directed draws only
    asym = ob.Graph.from_edge_index(torch.stack((g.raw_src, g.raw_dst)).to(DEV), g.n, symmetric=False)
    assert asym.validate() & 8
    sym = ob.Graph.from_edge_index(torch.stack((g.raw_src, g.raw_dst)).to(DEV), g.n)
    keep = g.raw_src != g.raw_dst
    if bool(keep.all()):
        assert torch.equal(sym.rowptr.cpu(), g.rowptr) and torch.equal(sym.col.cpu(), g.col)


@pytest.mark.parametrize("name", ["tiny", "tiny_dense", "cora", "pubmed"])
def test_adjoverlap_generic_and_spgemm(name):
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(512, "mixed")
    ed = e.to(DEV)
    _assert_rows_equal(ob.adjoverlap(G, G, ed), R.adjoverlap(A, A, e))
    # calresadj=True (utils.py:260-274): overlap and the two residual sets
    for got, ref in zip(ob.adjoverlap(G, G, ed, calresadj=True), R.adjoverlap(A, A, e, calresadj=True)):
        _assert_rows_equal(got, ref)
    # true A^2 (structure and 2-walk counts), then adjoverlap(adj, adj2, e) as NeighborOverlap_large.py:78-79
    a2 = R.adj2_true(A, keep_value=True)
    G2 = ob.spgemm_a2(G, with_value=True)
    assert torch.equal(G2.rowptr.cpu(), a2.rowptr())
    assert torch.equal(G2.col.cpu().long(), a2.col)
    assert torch.equal(G2.value.cpu(), a2.values())
    _assert_rows_equal(ob.adjoverlap(G, G2, ed), R.adjoverlap(A, R.Sp(a2.row, a2.col, None, a2.shape), e))
    # the reference's folded adj2byblock matrix
    bs = 16 if g.n < 100 else 1024
    f = R.adj2_folded(A, bs)
    GF = ob.sparse_tensor_multiply(G, bs)
    assert torch.equal(GF.rowptr.cpu(), f.rowptr())
    assert torch.equal(GF.col.cpu().long(), f.col)
    assert torch.equal(GF.value.cpu(), f.values())
    _assert_rows_equal(ob.adjoverlap(G, GF, ed), R.adjoverlap(A, f, e))


@pytest.mark.parametrize("name,B", [("tiny", 128), ("tiny_dense", 64), ("cora", 1152), ("pubmed", 2048),
                                    ("collab_s", 4096), ("citation2_s", 2048), ("ddi_s", 256)])
def test_cn_sets_bit_exact(name, B):
    """CN_k index sets and walk counts, k = 1..3, against the pygho-style oracle."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    kind = "stream" if name.startswith("citation2") else "mixed"
    e = g.query_edges(B, kind)
    order = 3 if g.nnz < 60000 or name.startswith("citation2") else 2
    got = ob.get_cn(G, e.to(DEV), order, weighted=True)
    ref = R.get_cn(A, e, order)
    for k in range(order):
        _assert_rows_equal(got[k], ref[k])
    # unweighted structure == adjoverlap(adj, adj2, e)
    got_s = ob.get_cn(G, e.to(DEV), 2, weighted=False)
    ref_s = R.adjoverlap(A, R.adj2_true(A), e)
    _assert_rows_equal(got_s[1], ref_s)


@pytest.mark.parametrize("name,B,kind", [("tiny", 128, "mixed"), ("cora", 1152, "mixed"), ("pubmed", 2048, "pos"),
                                         ("collab_s", 4096, "pos"), ("citation2_s", 2048, "pos")])
def test_order2_direct_kernel_bit_exact(name, B, kind):
    """Orders 1-2 on streams of short runs (k_cn_build_direct: flat over the records, sorted-list intersections) against
    the oracle, walk counts included; "pos" = links drawn from the edges, as a training batch is."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(B, kind)
    got = ob.get_cn(G, e.to(DEV), 2, weighted=True)
    ref = R.get_cn(A, e, 2)
    for k in range(2):
        _assert_rows_equal(got[k], ref[k])
    got1 = ob.get_cn(G, e.to(DEV), 1, weighted=True)
    _assert_rows_equal(got1[0], ref[0])


def test_order2_direct_kernel_hub_rows():
    """A 9 000-neighbour hub as source, as destination and as a common neighbour: lane walks (shorter row <= 32) and the
    deferred whole-warp intersections of two long rows, against the oracle."""
    n = 20000
    hub = torch.arange(1, 9001)
    src = torch.cat((torch.zeros(9000, dtype=torch.int64), synth.hash_randint(40000, n, 3, 1, "cpu")))
    dst = torch.cat((hub, synth.hash_randint(40000, n, 3, 2, "cpu")))
    G = ob.Graph.from_edge_index(torch.stack((src, dst)).to(DEV), n)
    A = R.sp_from_csr(G.rowptr.cpu(), G.col.cpu())
    e = torch.stack((torch.tensor([5, 0, 17, 9000, 3, 0]), torch.tensor([0, 7, 0, 0, 4, 12345])))
    got = ob.get_cn(G, e.to(DEV), 2, weighted=True)
    ref = R.get_cn(A, e, 2)
    for k in range(2):
        _assert_rows_equal(got[k], ref[k])
    assert got[1].col.numel() > 9000


def _oracle_cns(A, e, order, weighted):
    cns = R.get_cn(A, e, order)
    if not weighted:
        cns = [R.Sp(c.row, c.col, torch.ones(c.nnz), c.shape) for c in cns]
    return cns


def _mass(sp: R.Sp, x):
    return R.spmm_add(R.Sp(sp.row, sp.col, sp.values().abs(), sp.shape), x.abs())


@pytest.mark.parametrize("name,B,F", [("tiny", 96, 8), ("cora", 1152, 256), ("pubmed", 2048, 64), ("collab_s", 4096, 32),
                                      ("citation2_s", 2048, 32)])
@pytest.mark.parametrize("weighted,ip", [(True, 0.0), (False, 0.0), (True, 0.731), (False, 1.9)])
def test_cn5_aggregate(name, B, F, weighted, ip):
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(B, "stream" if name.startswith("citation2") else "mixed")
    x = g.features(F)
    cns = _oracle_cns(A, e, 2, weighted)
    r1, r2, rij, n1, n2 = R.cn5_aggregate(cns[0], cns[1], x, e, R.InnerProdState(ip), training=False)
    ip3 = torch.full((3,), ip, dtype=torch.float32, device=DEV)
    sess = ob.CNSession(G, e.to(DEV)).build(2, weighted)
    sess.stats(5, 0.0, ip3, 0)
    x1, x2, x3, xij = sess.aggregate(x.to(DEV), 5, 0.0, ip3)
    assert x3 is None
    _close(x1, r1, _mass(n1, x))
    _close(x2, r2, _mass(n2, x), rtol=1e-4 if ip else RTOL)
    assert torch.equal(xij.cpu(), rij)
    # the normalised matrices themselves: pattern bit-exact (explicit zeros kept), values close
    _assert_rows_equal(sess.extract(11, 5, 0.0, ip3), n1, vtol=1e-6)
    _assert_rows_equal(sess.extract(12, 5, 0.0, ip3), n2, vtol=1e-4 if ip else 1e-6)
    sess.release()
    assert int(sess.colstat.count_nonzero()) == 0


@pytest.mark.parametrize("name,B,F", [("tiny", 96, 8), ("cora", 1152, 32), ("citation2_s", 2048, 32)])
@pytest.mark.parametrize("ip", [0.0, 0.43])
def test_cn6_order3_aggregate(name, B, F, ip):
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(B, "stream" if name.startswith("citation2") else "mixed")
    x = g.features(F)
    cns = _oracle_cns(A, e, 3, True)
    r1, r2, r3, rij, n1, n2, n3 = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, R.InnerProdState(ip), training=False)
    ip3 = torch.full((3,), ip, dtype=torch.float32, device=DEV)
    sess = ob.CNSession(G, e.to(DEV)).build(3, True)
    sess.stats(5, 0.0, ip3, 0)
    x1, x2, x3, xij = sess.aggregate(x.to(DEV), 5, 0.0, ip3)
    tol = 1e-4 if ip else RTOL
    _close(x1, r1, _mass(n1, x))
    _close(x2, r2, _mass(n2, x), rtol=tol)
    _close(x3, r3, _mass(n3, x), rtol=tol)
    _assert_rows_equal(sess.extract(13, 5, 0.0, ip3), n3, vtol=tol)


@pytest.mark.parametrize("name,B,F,fill", [("tiny", 96, 8, 1.0), ("pubmed", 2048, 256, 1.0), ("ddi_s", 512, 64, 0.0)])
def test_cn7_aggregate(name, B, F, fill):
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(B, "mixed")
    x = g.features(F)
    cns = _oracle_cns(A, e, 2, False)
    r1, r2, rij, n1 = R.cn7_aggregate(cns[0], cns[1], x, e, fill)
    ip3 = torch.zeros(3, device=DEV)
    sess = ob.CNSession(G, e.to(DEV)).build(2, False)
    x1, x2, _, xij = sess.aggregate(x.to(DEV), 7, fill, ip3)
    _close(x1, r1, _mass(n1, x))
    _close(x2, r2, _mass(cns[1], x))
    assert torch.equal(xij.cpu(), rij)


def test_stream_of_batches_equals_per_batch_calls():
    g = GRAPHS["citation2_s"]()
    G = _graph(g)
    e = g.query_edges(5 * 512 + 100, "stream").to(DEV)
    x = g.features(32).to(DEV)
    ip3 = torch.full((3,), 0.25, device=DEV)
    whole = ob.cn_aggregate_eval(G, e, x, 512, 3, True, 5, 0.0, ip3)
    tiny_budget = ob.cn_aggregate_eval(G, e, x, 512, 3, True, 5, 0.0, ip3, budget_bytes=2 * 32 * g.n)
    for s in range(0, e.shape[1], 512):
        part = ob.cn_aggregate_eval(G, e[:, s:s + 512], x, 512, 3, True, 5, 0.0, ip3)
        for w, t, p in zip(whole, tiny_budget, part):
            assert torch.equal(w[s:s + 512], p), "stream result differs from the single-batch call (must be bit-identical)"
            assert torch.equal(t[s:s + 512], p)


def test_training_running_inner_product_and_backward():
    """cn5 / cn6 in training mode: the running mean of model.py:2241-2250 over consecutive batches and
    grad_x against autograd through the oracle."""
    g = GRAPHS["cora"]()
    G, A = _graph(g), _sp(g)
    F = 16
    torch.manual_seed(0)
    for order, cls in ((2, ob.CNLinkPredictorOringin), (3, ob.CNLinkPredictor3hopCNs)):
        pred = cls(F, F, 1, 3, 0.0, weighted=True).to(DEV).train()
        st = R.InnerProdState()
        for step in range(3):
            e = torch.stack((synth.hash_randint(300, g.n, 50 + step, 1, "cpu"), synth.hash_randint(300, g.n, 50 + step, 2, "cpu")))
            pe = g.query_edges(300, "pos")
            e = torch.cat((e, pe), 1)
            x = g.features(F).clone().requires_grad_(True)
            cns = _oracle_cns(A, e, order, True)
            if order == 2:
                r1, r2, rij, n1, n2 = R.cn5_aggregate(cns[0], cns[1], x, e, st, training=True)
                ref_outs, masses = [r1, r2], [_mass(n1, x.detach()), _mass(n2, x.detach())]
            else:
                r1, r2, r3, rij, n1, n2, n3 = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, st, training=True)
                ref_outs = [r1, r2, r3]
                masses = [_mass(n1, x.detach()), _mass(n2, x.detach()), _mass(n3, x.detach())]
            xd = x.detach().to(DEV).requires_grad_(True)
            x1, x2, x3, xij, _ = pred.cn_stage(xd, G, e.to(DEV))
            assert pred.n == st.n
            assert abs(pred.innerprod.item() - st.innerprod.item()) <= 1e-4 * (1 + abs(st.innerprod.item()))
            got = [x1, x2] + ([x3] if order == 3 else [])
            for a, b, m in zip(got, ref_outs, masses):
                _close(a.detach(), b.detach(), m, rtol=2e-4)
            wts = [torch.randn_like(o) for o in ref_outs] + [torch.randn_like(rij)]
            loss_ref = sum((o * w).sum() for o, w in zip(ref_outs + [rij], wts))
            loss_ref.backward()
            loss = sum((o * w.to(DEV)).sum() for o, w in zip(got + [xij], wts))
            loss.backward()
            gm = x.grad.abs().max().item()
            assert (xd.grad.cpu() - x.grad).abs().max().item() <= 5e-4 * (1 + gm)


def test_predictor_end_to_end_matches_oracle_heads():
    """Full cn5 / cn7 forward (MLP heads included) == oracle aggregates fed through the same heads."""
    g = GRAPHS["pubmed"]()
    G, A = _graph(g), _sp(g)
    F = 32
    e = g.query_edges(1024, "mixed")
    x = g.features(F)
    torch.manual_seed(1)
    p5 = ob.CNLinkPredictorOringin(F, F, 1, 3, 0.0).to(DEV).eval()
    p7 = ob.CNLinkPredictorbaselearn(F, F, 1, 3, 0.0).to(DEV).eval()
    cns = _oracle_cns(A, e, 2, False)
    with torch.no_grad():
        out5 = p5(x.to(DEV), G, None, None, e.to(DEV)).cpu()
        r1, r2, rij, _, _ = R.cn5_aggregate(cns[0], cns[1], x, e, R.InnerProdState(), training=False)
        ref5 = p5.cpu()._head(r1, r2, None, rij)
        assert torch.allclose(out5, ref5, rtol=1e-4, atol=1e-4)

        class Args:
            sum = 1
        out7 = p7.multidomainforward(x.to(DEV), G, None, None, e.to(DEV), Args()).cpu()
        r1, r2, rij, _ = R.cn7_aggregate(cns[0], cns[1], x, e, 1.0)
        ref7 = p7.cpu()._head(r1, r2, None, rij)
        assert torch.allclose(out7, ref7, rtol=1e-4, atol=1e-4)
    pos, neg = out5[:512].flatten(), out5[512:].flatten()
    assert abs(R.hits_at_k(pos, neg, 20) - R.hits_at_k(ref5[:512].flatten(), ref5[512:].flatten(), 20)) <= 0.01


@pytest.mark.parametrize("name,F", [("tiny", 5), ("cora", 256), ("pubmed", 64), ("collab_s", 128), ("ddi_s", 64), ("citation2_s", 32)])
def test_gnn_aggregation(name, F, lib_options):
    lib_options(spmm_tma=2)          # the register gather (k_spmm)
    _check_gnn_aggregation(name, F)


@pytest.mark.parametrize("name,F", [("cora", 256), ("pubmed", 64), ("collab_s", 128), ("ddi_s", 64), ("citation2_s", 32)])
def test_gnn_aggregation_bulk_gather(name, F, lib_options):
    """The same checks with neighbour rows gathered by cp.async.bulk + mbarrier (k_spmm_tma), forced on."""
    from ocn_b200 import _lib
    lib_options(spmm_tma=1)
    before = _lib.lib().ocn_launch_count()
    _check_gnn_aggregation(name, F)
    assert _lib.lib().ocn_launch_count() > before


@pytest.mark.parametrize("name,F", [("cora", 256), ("pubmed", 64), ("collab_s", 128), ("ddi_s", 64), ("citation2_s", 32)])
def test_gnn_aggregation_cp_async_ring(name, F, lib_options):
    """The same checks on k_spmm_async (the per-warp shared-memory ring fed by per-thread cp.async), forced on."""
    lib_options(spmm_tma=4)
    _check_gnn_aggregation(name, F)


@pytest.mark.parametrize("name,F", [("cora", 256), ("pubmed", 64), ("collab_s", 128), ("ddi_s", 64), ("citation2_s", 32)])
def test_gnn_aggregation_lane_per_feature(name, F, lib_options):
    """The same checks on k_spmm_lane (a lane owns F / 32 features of every neighbour row), forced on."""
    lib_options(spmm_tma=3)
    _check_gnn_aggregation(name, F)


def _check_gnn_aggregation(name, F):
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    x = g.features(F)
    xd = x.to(DEV)
    ones = R.Sp(A.row, A.col, torch.ones(A.nnz), A.shape)
    deg = (g.rowptr[1:] - g.rowptr[:-1]).float().unsqueeze(1)
    mass_sum = R.spmm_add(ones, x.abs())
    _close(ob.pure_conv(xd, G, "sum"), R.pure_conv(x, A, "sum"), mass_sum)
    _close(ob.pure_conv(xd, G, "mean"), R.pure_conv(x, A, "mean"), mass_sum / deg.clamp(min=1))
    assert torch.equal(ob.pure_conv(xd, G, "max").cpu(), R.pure_conv(x, A, "max"))
    _close(ob.pure_conv(xd, G, "gcn"), R.pure_conv(x, A, "gcn"), mass_sum + x.abs())
    _close(ob.pure_conv3_gcn(xd, G), R.pure_conv3_gcn(x, A), mass_sum)
    _close(ob.gcnconv_propagate(xd, G, True, True), R.gcnconv_propagate(x, A, True, True), mass_sum + x.abs())
    _close(ob.gcnconv_propagate(xd, G, False), R.gcnconv_propagate(x, A, False, False), mass_sum)
    # backward of the sum / gcn aggregation (A-hat symmetric)
    xr = x.clone().requires_grad_(True)
    w = torch.randn(g.n, F)
    (R.pure_conv(xr, A, "gcn") * w).sum().backward()
    xg = xd.clone().requires_grad_(True)
    (ob.pure_conv(xg, G, "gcn") * w.to(DEV)).sum().backward()
    _close(xg.grad, xr.grad, R.spmm_add(ones, w.abs()) + w.abs(), rtol=1e-4)
    xr = x.clone().requires_grad_(True)
    (R.pure_conv(xr, A, "mean") * w).sum().backward()
    xg = xd.clone().requires_grad_(True)
    (ob.pure_conv(xg, G, "mean") * w.to(DEV)).sum().backward()
    _close(xg.grad, xr.grad, R.spmm_add(ones, w.abs()), rtol=1e-4)
    # backward of the max aggregation: the first arg-max of every (row, feature) receives the gradient
    xg = xd.clone().requires_grad_(True)
    (ob.pure_conv(xg, G, "max") * w.to(DEV)).sum().backward()
    _close(xg.grad, R.spmm_max_backward(A, x, w), R.spmm_add(ones, w.abs()), rtol=1e-5)
    xt = torch.round(x * 2) / 2  # many ties
    xg = xt.to(DEV).requires_grad_(True)
    (ob.pure_conv(xg, G, "max") * w.to(DEV)).sum().backward()
    _close(xg.grad, R.spmm_max_backward(A, xt, w), R.spmm_add(ones, w.abs()), rtol=1e-5)


def test_spmm_on_cn_matrix_with_values():
    """spmm_add(normalized_cn, x) on an explicit [B x N] matrix (model.py:2426-2427)."""
    g = GRAPHS["cora"]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(700, "mixed")
    x = g.features(24)
    cn2 = ob.get_cn(G, e.to(DEV), 2, True)[1]
    ref = R.get_cn(A, e, 2)[1]
    _close(ob.spmm_add(cn2, x.to(DEV)), R.spmm_add(ref, x), _mass(ref, x))


def test_full_size_citation2_walk_counts():
    """BASELINE.json configs[4] at full size: CN_k values of a few links against an independent
    torch-on-GPU propagation u_k = A^k e_j (size-independent property: C_k = u_k restricted to N(i))."""
    g = synth.make_graph("citation2", device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    assert G.validate() == 0
    e = g.query_edges(2048, "stream", device=DEV)
    cns = ob.get_cn(G, e, 3, weighted=True)
    row = G.row()
    colL = g.col.long()
    for b in (0, 777, 1500, 2047):
        i, j = int(e[0, b]), int(e[1, b])
        u = torch.zeros(g.n, dtype=torch.float64, device=DEV)
        u[j] = 1
        ni = colL[int(g.rowptr[i]):int(g.rowptr[i + 1])]
        for k in range(3):
            u = torch.zeros_like(u).index_add_(0, row, u[colL])
            vals = u[ni]
            keep = vals > 0
            s, t = int(cns[k].rowptr[b]), int(cns[k].rowptr[b + 1])
            assert torch.equal(cns[k].col[s:t], ni[keep])
            assert torch.equal(cns[k].value[s:t].double(), vals[keep])


@pytest.mark.parametrize("name,B", [("tiny", 128), ("tiny_dense", 64), ("cora", 1152), ("citation2_s", 2048)])
@pytest.mark.parametrize("hub", [2, 7, 40])
def test_hub_stage_bit_exact(name, B, hub):
    """Order-3 walk counts with rows of >= hub columns routed through the hub stage (cn_hub.cu)
    against the oracle; the per-link table kernel handles the rest."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(B, "stream" if name.startswith("citation2") else "mixed")
    ref = R.get_cn(A, e, 3)
    sess = ob.CNSession(G, e.to(DEV), None, 3, hub_degree=hub)
    assert sess.hub_degree == hub and sess.hub_bytes > 0
    if name == "cora":
        assert sess.plan_host[11] > 4096, "this case is meant to need several position windows"
    sess.build(3, True, with_stats=False)
    for k in range(3):
        _assert_rows_equal(sess.extract(k + 1), ref[k])
    nodes = [v for k, v in G._ws.items() if isinstance(k, tuple) and k[0] == "hub_node"]
    assert nodes and all(bool((v == 0).all()) for v in nodes), "node index not restored"
    # several batches in one stream (runs are cut at batch boundaries)
    got = ob.get_cn(G, e.to(DEV), 3, True, hub_degree=hub, batch_size=max(8, B // 5))
    for k in range(3):
        _assert_rows_equal(got[k], ref[k])


@pytest.mark.parametrize("warp_win,cta_win", [(64, 100000), (64, 96), (100000, 100000)])
def test_hub_stage_position_windows(lib_options, warp_win, cta_win):
    """Streams with more positions than a counter window holds: CTA-per-item counters and several passes
    (forced on a small graph through the library's test hooks) give the same records."""
    g = GRAPHS["cora"]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(300, "mixed")
    ref = R.get_cn(A, e, 3)
    lib_options(hub_window=warp_win, hub_cta_window=cta_win)
    got = ob.get_cn(G, e.to(DEV), 3, True, hub_degree=3, batch_size=128)
    for k in range(3):
        _assert_rows_equal(got[k], ref[k])


@pytest.mark.parametrize("n,links,whole", [(40, 400, True), (40, 64, False), (70, 900, True), (1500, 96, False), (1500, 20000, True),
                                           (3000, 96, False), (4267, 3000, False), (9000, 96, False)])
@pytest.mark.parametrize("order", [1, 2])
def test_dense_build_equals_the_walk_kernels(n, links, whole, order):
    """Orders 1-2 on a dense graph (mean degree >= n / 64) come from bit-vector rows: per-position row products
    (k_cn_build_dense), or -- ``whole``: streams with at least n^2 / 2 positions -- a tiled product for the whole matrix
    A^2 and a gather (k_dense_a2, k_cn_build_from_a2).  The same session built by the walk kernels (the plan's dense flag
    cleared) must give identical records and column statistics."""
    from ocn_b200.cn import PLAN_DENSE
    g = synth.tiny_graph(n, n * (n // 48 + 2), 5)
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    e = g.query_edges(links, "mixed").to(DEV)
    a = ob.CNSession(G, e, 32, order)
    assert a.dense, "this graph is meant to take the dense build"
    assert (n <= 8192 and 2 * a.num_records >= n * n) == whole
    a.build(order, True)
    b = ob.CNSession(G, e, 32, order)
    b.dense, b.hub_bytes = False, 0
    b.plan_host[PLAN_DENSE] = 0
    b.build(order, True)
    nb = a.num_records * 8
    assert nb > 0 and torch.equal(a.records[:nb], b.records[:nb])
    assert torch.equal(a.colstat, b.colstat)
    a.release(); b.release()


@pytest.mark.parametrize("name,B,kind", [("tiny_dense", 200, "mixed"), ("cora", 1500, "mixed"), ("citation2_s", 4096, "stream")])
@pytest.mark.parametrize("heavy_run", [1, 6, 20])
def test_hub_stage_heavy_sources_get_their_own_pass(lib_options, name, B, kind, heavy_run):
    """Runs whose source has more than `heavy_run` neighbours are indexed in a second pass (forced on small graphs
    through the plan's test hook; 1024 in production): every record must still equal the oracle's, for streams
    that are all light, mixed, and (heavy_run = 1) almost all heavy."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(B, kind)
    ref = R.get_cn(A, e, 3)
    lib_options(hub_heavy_run=heavy_run)
    sess = ob.CNSession(G, e.to(DEV), 512, 3, hub_degree=3)
    assert sess.hub_degree == 3
    deg = (g.rowptr[1:] - g.rowptr[:-1])
    if bool((deg[e[0]] > heavy_run).any()):
        assert sess.plan_host[14] > 0 and sess.plan_host[13] > 0, "expected a heavy pass"
    sess.build(3, True, with_stats=False)
    for k in range(3):
        _assert_rows_equal(sess.extract(k + 1), ref[k])
    # and with the CTA-wide counter window forced small in the same run (several window launches per pass)
    lib_options(hub_window=64, hub_cta_window=160)
    got = ob.get_cn(G, e.to(DEV), 3, True, hub_degree=3, batch_size=512)
    for k in range(3):
        _assert_rows_equal(got[k], ref[k])


def test_hub_stage_matches_table_kernel_at_scale():
    """citation2 shape at 5 % size, 16 batches of the evaluation stream: hub stage on (automatic
    threshold and a low one) == hub stage off, bit for bit, for every record."""
    g = synth.make_graph("citation2", scale=0.05, device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    e = g.query_edges(16 * 2048, "stream", device=DEV)
    off = ob.CNSession(G, e, 2048, 3, hub_degree=-1)
    assert off.hub_degree == 0
    off.build(3, True, with_stats=False)
    for hub in (0, 24):
        on = ob.CNSession(G, e, 2048, 3, hub_degree=hub)
        assert on.hub_degree > 0
        on.build(3, True, with_stats=False)
        nb = on.num_records * 8  # the buffers are sized in buckets: compare the records, not the slack behind them
        assert on.num_records == off.num_records and torch.equal(on.records[:nb], off.records[:nb])
    # a stream with one run per link (training shape) keeps the stage off
    many = ob.CNSession(G, g.query_edges(4096, "neg", device=DEV), 2048, 3)
    assert many.hub_degree == 0


def _grouped_stream(g, nsrc, per_src, seed=0):
    """nsrc random sources with neighbours, each against per_src uniform destinations (the evaluation-stream shape)."""
    deg = g.rowptr[1:] - g.rowptr[:-1]
    cand = torch.nonzero(deg > 0).flatten()
    gen = torch.Generator().manual_seed(seed)
    srcs = cand[torch.randint(0, cand.numel(), (nsrc,), generator=gen)]
    src = srcs.repeat_interleave(per_src)
    dst = torch.randint(0, g.n, (nsrc * per_src,), generator=gen)
    return torch.stack((src, dst))


@pytest.mark.parametrize("name,nsrc,per_src,hub", [("tiny_dense", 30, 12, 2), ("tiny", 100, 6, 2), ("cora", 90, 30, 3),
                                                   ("cora", 128, 8, 8), ("citation2_s", 60, 50, 4), ("tiny_dense", 129, 3, 2)])
def test_hub_run_segment_index(lib_options, name, nsrc, per_src, hub):
    """Opt-in index layout for streams of up to 128 runs: 32-byte node entries with exact run sets and run-segment
    starts (search-free per-link look-ups), with the whole-list walk of the shared rows or k_cn_hub_count_seg as
    walker.  The records equal the oracle's under the default (folded 64-bit sets) and equal it bit for bit under
    both options.  tiny_dense gives lists with several entries per run (the sidx path); 129 sources fall back to
    the folded layout."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = _grouped_stream(g, nsrc, per_src, seed=nsrc)
    ref = R.get_cn(A, e, 3)
    sess = ob.CNSession(G, e.to(DEV), None, 3, hub_degree=hub)
    assert sess.hub_degree == hub
    if nsrc <= 128 and sess.plan_host[11] <= 4096:
        assert sess.num_runs <= 128
    sess.build(3, True, with_stats=False)
    for k in range(3):
        _assert_rows_equal(sess.extract(k + 1), ref[k])
    nodes = [v for k, v in G._ws.items() if isinstance(k, tuple) and k[0] == "hub_node"]
    assert nodes and all(bool((v == 0).all()) for v in nodes), "node index not restored"
    nb = sess.num_records * 8
    for mode in ({"hub_exact": 1}, {"hub_walker": 1}):
        lib_options(hub_walker=0, hub_exact=0)
        lib_options(**mode)
        other = ob.CNSession(G, e.to(DEV), None, 3, hub_degree=hub).build(3, True, with_stats=False)
        assert torch.equal(sess.records[:nb], other.records[:nb]), mode


@pytest.mark.parametrize("name,nsrc,per_src,batch,F", [("cora", 20, 40, 64, 32), ("tiny_dense", 12, 50, 100, 8), ("citation2_s", 9, 300, 512, 64),
                                                       ("cora", 6, 200, 2048, 256), ("pubmed", 30, 17, 96, 32)])
@pytest.mark.parametrize("order,variant,weighted", [(3, 5, True), (2, 5, False), (2, 7, True)])
def test_run_grouped_kernels(lib_options, name, nsrc, per_src, batch, F, order, variant, weighted):
    """Streams of long runs (>= 16 links per source) take the run-grouped kernels of cn_grouped.cu for the column
    statistics, batch scalars, aggregation and release: integer statistics and batch scalars equal the per-link
    kernels' bit for bit (=> identical normalised matrices), the aggregates agree within the fp32 tolerance (the
    sums run in a different order) and with the oracle; batches that cut runs, position tiles (> 32 neighbours) and
    windows that hold several runs are all in the cases."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    e = _grouped_stream(g, nsrc, per_src, seed=per_src)
    ed, x = e.to(DEV), g.features(F)
    ip = 0.37
    ip3 = torch.full((3,), ip, dtype=torch.float32, device=DEV)
    fill = 1.0 if variant == 7 else 0.0
    res = {}
    for off in (0, 1):
        lib_options(grouped_off=off)
        sess = ob.CNSession(G, ed, batch, order).build(order, weighted)
        assert sess.T >= 16 * sess.num_runs
        if variant == 5:
            sess.stats(5, fill, ip3, 0)
        outs = sess.aggregate(x.to(DEV), variant, fill, ip3)
        mats = [sess.extract(10 + k, variant, fill, ip3) for k in range(1, (order if variant == 5 else 1) + 1)]
        res[off] = (sess.bscal.clone(), outs, mats)
        sess.release()
        assert int(sess.colstat.count_nonzero()) == 0, "release left statistics behind"
    assert torch.equal(res[0][0].view(torch.int32), res[1][0].view(torch.int32)), "batch scalars differ between the grouped and the per-link kernels"
    for a, b in zip(res[0][2], res[1][2]):
        assert torch.equal(a.col, b.col) and torch.equal(a.value, b.value)
    for a, b in zip(res[0][1], res[1][1]):
        if a is not None:
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (a - b).abs().max()
    assert torch.equal(res[0][1][3], res[1][1][3])   # pair term
    # against the oracle, batch by batch (every batch is normalised on its own)
    T = e.shape[1]
    for s in range(0, T, batch):
        eb = e[:, s:s + batch]
        cns = _oracle_cns(A, eb, order, weighted)
        if variant == 7:
            r = R.cn7_aggregate(cns[0], cns[1], x, eb, fill)
            refs, masses = [r[0], r[1]], [_mass(r[3], x), _mass(cns[1], x)]
        elif order == 3:
            r = R.cn6_aggregate(cns[0], cns[1], cns[2], x, eb, R.InnerProdState(ip), training=False)
            refs, masses = [r[0], r[1], r[2]], [_mass(r[4], x), _mass(r[5], x), _mass(r[6], x)]
        else:
            r = R.cn5_aggregate(cns[0], cns[1], x, eb, R.InnerProdState(ip), training=False)
            refs, masses = [r[0], r[1]], [_mass(r[3], x), _mass(r[4], x)]
        for k, (ref, mass) in enumerate(zip(refs, masses)):
            _close(res[0][1][k][s:s + batch], ref, mass, rtol=1e-4)


def test_torch_custom_ops():
    """torch.ops.ocn.* (north_star: the C ABI exposed as PyTorch custom ops), incl. autograd registration."""
    g = GRAPHS["cora"]()
    G, A = _graph(g), _sp(g)
    e = g.query_edges(600, "mixed")
    ed = e.to(DEV)
    x = g.features(32)
    rp, col = torch.ops.ocn.rows_intersect(G.rowptr, G.col, G.rowptr, G.col, ed)
    ref = R.adjoverlap(A, A, e)
    assert torch.equal(rp.cpu(), ref.rowptr()) and torch.equal(col.cpu(), ref.col)
    # spmm with autograd
    xr = x.clone().requires_grad_(True)
    w = torch.randn(g.n, 32)
    (R.pure_conv(xr, A, "sum") * w).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    out = torch.ops.ocn.spmm_csr(G.rowptr, G.col, None, xg, 0)
    (out * w.to(DEV)).sum().backward()
    ones = R.Sp(A.row, A.col, torch.ones(A.nnz), A.shape)
    _close(out.detach(), R.pure_conv(x, A, "sum"), R.spmm_add(ones, x.abs()))
    _close(xg.grad, xr.grad, R.spmm_add(ones, w.abs()), rtol=1e-4)
    # gcn with autograd
    norm = ob.gcn_norm(G)
    xg2 = x.to(DEV).requires_grad_(True)
    o2 = torch.ops.ocn.gcn_spmm(G.rowptr, G.col, norm, xg2, 3)
    _close(o2.detach(), R.pure_conv(x, A, "gcn"), R.spmm_add(ones, x.abs()) + x.abs())
    (o2 * w.to(DEV)).sum().backward()
    xr2 = x.clone().requires_grad_(True)
    (R.pure_conv(xr2, A, "gcn") * w).sum().backward()
    _close(xg2.grad, xr2.grad, R.spmm_add(ones, w.abs()) + w.abs(), rtol=1e-4)
    # A^2 and the fused stream op
    rp2, c2, v2 = torch.ops.ocn.spgemm_a2(G.rowptr, G.col, 0, True)
    a2 = R.adj2_true(A, keep_value=True)
    assert torch.equal(rp2.cpu(), a2.rowptr()) and torch.equal(c2.cpu().long(), a2.col) and torch.equal(v2.cpu(), a2.values())
    ip3 = torch.zeros(3, device=DEV)
    x1, x2, x3, xij = torch.ops.ocn.cn_aggregate(G.rowptr, G.col, ed, x.to(DEV), ip3, 600, 2, True, 5, 0.0)
    cns = R.get_cn(A, e, 2)
    r1, r2, rij, n1, n2 = R.cn5_aggregate(cns[0], cns[1], x, e, R.InnerProdState(), training=False)
    _close(x1, r1, _mass(n1, x))
    _close(x2, r2, _mass(n2, x))
    assert x3.shape[0] == 0 and torch.equal(xij.cpu(), rij)


@pytest.mark.parametrize("name,fold", [("cora", 0), ("pubmed", 1024), ("ddi_s", 64)])
def test_explicit_large_driver_flow(name, fold):
    """NeighborOverlap_large.py:68-82 as written: adj2 = A@A (or the folded adj2byblock matrix), cn1/cn2 from
    adjoverlap, then cn5 / cn7 on the explicit matrices -- against the oracle fed with the same matrices."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    F = 16
    e = g.query_edges(512, "mixed")
    ed, x = e.to(DEV), g.features(F)
    G2 = ob.sparse_tensor_multiply(G, fold) if fold else ob.spgemm_a2(G)
    a2 = R.adj2_folded(A, fold) if fold else R.adj2_true(A)
    cn1, cn2 = ob.adjoverlap(G, G, ed), ob.adjoverlap(G, G2, ed)
    r1, r2 = R.adjoverlap(A, A, e), R.adjoverlap(A, a2, e)
    _assert_rows_equal(cn1, r1)
    _assert_rows_equal(cn2, r2)
    torch.manual_seed(3)
    p5 = ob.CNLinkPredictorOringin(F, F, 1, 3, 0.0).to(DEV).train()
    st = R.InnerProdState()
    for _ in range(2):  # two training calls: the running inner product moves
        with torch.no_grad():
            out = p5.multidomainforward(x.to(DEV), G, cn1, cn2, ed)
        o1, o2, oij, n1, n2 = R.cn5_aggregate(r1, r2, x, e, st, training=True)
        ref = p5._head(o1.to(DEV), o2.to(DEV), None, oij.to(DEV))
        assert torch.allclose(out, ref, rtol=1e-3, atol=1e-4)
        assert abs(p5.innerprod.item() - st.innerprod.item()) <= 1e-4 * (1 + abs(st.innerprod.item()))
    p7 = ob.CNLinkPredictorbaselearn(F, F, 1, 3, 0.0).to(DEV).eval()

    class Args:
        sum = 1
    with torch.no_grad():
        out7 = p7.multidomainforward(x.to(DEV), G, cn1, cn2, ed, Args())
    o1, o2, oij, _ = R.cn7_aggregate(r1, r2, x, e, 1.0)
    assert torch.allclose(out7, p7._head(o1.to(DEV), o2.to(DEV), None, oij.to(DEV)), rtol=1e-3, atol=1e-4)


def _sp_of_graph(G):
    row, col, val = G.coo()
    v = val.cpu() if val is not None else torch.ones(col.numel())
    return R.Sp(row.cpu(), col.cpu(), v, (G.n, G.n_cols))


@pytest.mark.parametrize("name,F,dp", [("cora", 32, 0.51), ("pubmed", 64, 0.25), ("citation2_s", 32, 0.2), ("tiny_dense", 8, 0.07)])
def test_dropadj_weighted_aggregation_and_transpose_backward(name, F, dp):
    """DropAdj (model.py:211-229) on the GNN's adjacency as GCN.forward applies it per layer (model.py:312): a torch
    Bernoulli mask over the stored entries, survivors rescaled by 1/(1-dp).  The dropped matrix is NOT symmetric (the
    two directions of an edge are dropped independently), so forward AND backward of every aggregation the convs
    use (puregcn with its self term, PureConv3's gcn, sum = gin, mean, max) are checked against autograd through the
    oracle on the very same mask."""
    g = GRAPHS[name]()
    G = ob.Graph.from_edge_index(torch.stack((g.raw_src, g.raw_dst)).to(DEV), g.n)
    assert G.is_symmetric()
    torch.manual_seed(5)
    drop = ob.predictor.DropAdj(dp).to(DEV).train()
    Gd = drop(G)
    assert Gd.nnz < G.nnz and Gd.value is not None and not Gd.is_symmetric()
    assert abs(Gd.nnz / G.nnz - (1 - dp)) < 0.05
    assert torch.allclose(Gd.value, torch.full_like(Gd.value, 1 / (1 - dp)))
    assert drop.eval()(G) is G                        # no drop outside training
    # the kept entries are a subset of the original rows, in order
    key_all = G.row() * G.n + G.col.long()
    key_kept = Gd.row() * G.n + Gd.col.long()
    assert bool(torch.isin(key_kept, key_all).all()) and bool((key_kept[1:] > key_kept[:-1]).all())
    A = _sp_of_graph(Gd)
    x = g.features(F)
    w = torch.randn(g.n, F, generator=torch.Generator().manual_seed(1))
    ones = R.Sp(A.row, A.col, A.values().abs(), A.shape)
    # gradient mass: |w| pushed back through |A|^T
    At = R.Sp(A.col, A.row, A.values().abs(), (A.shape[1], A.shape[0]))
    order = torch.argsort(At.row * At.shape[1] + At.col)
    At = R.Sp(At.row[order], At.col[order], At.values()[order], At.shape)
    gmass = R.spmm_add(At, w.abs()) + w.abs()
    mass = R.spmm_add(ones, x.abs()) + x.abs()
    for aggr in ("gcn", "gcn3", "sum", "mean", "max"):
        xg = x.to(DEV).requires_grad_(True)
        out = ob.pure_conv3_gcn(xg, Gd) if aggr == "gcn3" else ob.pure_conv(xg, Gd, aggr)
        (out * w.to(DEV)).sum().backward()
        if aggr == "max":   # the oracle's max is not differentiable by autograd: its backward is restated separately
            _close(out.detach(), R.pure_conv(x, A, "max"), mass, rtol=1e-5)
            _close(xg.grad, R.spmm_max_backward(A, x, w), gmass, rtol=1e-5)
            continue
        xr = x.clone().requires_grad_(True)
        ref = R.pure_conv3_gcn(xr, A) if aggr == "gcn3" else R.pure_conv(xr, A, aggr)
        (ref * w).sum().backward()
        _close(out.detach(), ref.detach(), mass, rtol=1e-4)
        _close(xg.grad, xr.grad, gmass, rtol=1e-4)


def test_directed_graph_gcn_backward_uses_the_transpose():
    """a directed (non-symmetric) unit-weight adjacency must not take the self-adjoint backward."""
    g = GRAPHS["cora"]()
    G = ob.Graph.from_edge_index(torch.stack((g.raw_src, g.raw_dst)).to(DEV), g.n, symmetric=False)
    assert not G.is_symmetric()
    A = _sp_of_graph(G)
    x = g.features(16)
    w = torch.randn(g.n, 16, generator=torch.Generator().manual_seed(2))
    for mode, ref_fn in ((3, lambda t: R.pure_conv(t, A, "gcn")), (4, lambda t: R.pure_conv3_gcn(t, A))):
        xr = x.clone().requires_grad_(True)
        (ref_fn(xr) * w).sum().backward()
        xg = x.to(DEV).requires_grad_(True)
        out = ob.pure_conv(xg, G, "gcn") if mode == 3 else ob.pure_conv3_gcn(xg, G)
        (out * w.to(DEV)).sum().backward()
        assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-5)
        # the custom op with the flag
        xo = x.to(DEV).requires_grad_(True)
        o2 = torch.ops.ocn.gcn_spmm(G.rowptr, G.col, ob.gcn_norm(G), xo, mode, None, False)
        (o2 * w.to(DEV)).sum().backward()
        assert torch.allclose(xo.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-5)


def test_out_of_range_links_raise():
    """a target link outside [0, n) raises (the reference: IndexError) instead of reading out of bounds,
    on the fused path (the plan counts them and stops) and on the generic set operations; the next call works."""
    g = GRAPHS["cora"]()
    G = _graph(g)
    e = g.query_edges(64, "mixed").to(DEV)
    for bad_val, end in ((g.n, 0), (-1, 1), (g.n + 12345, 1)):
        bad = e.clone()
        bad[end, 7] = bad_val
        with pytest.raises(IndexError):
            ob.CNSession(G, bad, None, 3)
        with pytest.raises(IndexError):
            ob.get_cn(G, bad, 2)
        with pytest.raises(IndexError):
            ob.adjoverlap(G, G, bad)
        with pytest.raises(IndexError):
            ob.adjoverlap(G, G, bad, calresadj=True)
    A = _sp(g)
    ref = R.get_cn(A, e.cpu(), 2)
    got = ob.get_cn(G, e, 2)
    for k in range(2):
        _assert_rows_equal(got[k], ref[k])


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("name,fold", [("tiny", 0), ("tiny", 32), ("tiny_dense", 0), ("tiny_dense", 16), ("cora", 0), ("cora", 1024),
                                       ("ddi_s", 0), ("ddi_s", 64), ("ddi_s", 96), ("pubmed", 0), ("pubmed", 1024)])
def test_spgemm_kernels_agree_with_the_oracle(lib_options, name, fold, mode):
    """The four A^2 kernels -- global-scratch accumulator (1), shared-memory row accumulator (2), dense bit-matrix
    rows (3), the whole dense count matrix + compaction (4: true A^2 only, folded calls fall back) -- forced in turn:
    structure and 2-walk counts bit-exact against torch.sparse on the CPU, true and folded (a fold that is no
    multiple of 32 leaves the dense kernel and falls back), structure-only too."""
    g = GRAPHS[name]()
    G, A = _graph(g), _sp(g)
    lib_options(spgemm_mode=mode)
    ref = R.adj2_folded(A, fold) if fold else R.adj2_true(A, keep_value=True)
    got = ob.spgemm_a2(G, fold, True)
    assert torch.equal(got.rowptr.cpu(), ref.rowptr())
    assert torch.equal(got.col.cpu().long(), ref.col)
    assert torch.equal(got.value.cpu(), ref.values())
    s = ob.spgemm_a2(G, fold, False)
    assert s.value is None and torch.equal(s.rowptr, got.rowptr) and torch.equal(s.col, got.col)


def test_spgemm_whole_matrix_numeric_without_symbolic():
    """The numeric call of the whole-matrix mode reuses the symbolic call's count matrix through a stamp in the scratch;
    on a scratch that does not carry the stamp (zeroed, or stamped for another graph) it must compute the matrix itself."""
    from ocn_b200 import _lib
    from ocn_b200.cn import _stream
    g = GRAPHS["ddi_s"]()
    G = _graph(g)
    ref = ob.spgemm_a2(G, 0, True)
    L = _lib.lib()
    st = _stream(G.device)
    scratch = torch.zeros(L.ocn_spgemm_scratch_bytes(G.n, G.nnz, 0), dtype=torch.uint8, device=DEV)
    for dirty in (False, True):
        if dirty:  # a stamp of some other graph, stale matrix contents
            scratch.random_(0, 255)
        col = torch.empty_like(ref.col)
        val = torch.empty_like(ref.value)
        _lib.check(L.ocn_spgemm_a2_numeric(_lib.ptr(G.rowptr), _lib.ptr(G.col), G.n, G.nnz, 0, _lib.ptr(scratch),
                                           _lib.ptr(ref.rowptr), _lib.ptr(col), _lib.ptr(val), st), "ocn_spgemm_a2_numeric")
        assert torch.equal(col, ref.col) and torch.equal(val, ref.value)
