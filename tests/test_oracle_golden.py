"""Pin the CPU oracle: hand-derived vectors of SURVEY.md §8c + an independent dense brute force."""
import numpy as np
import pytest
import torch

from oracle import brute
from oracle import ref_ops as R
from ocn_b200 import synth


def _sp_from_pairs(pairs, n):
    pairs = sorted(pairs)
    r = torch.tensor([p[0] for p in pairs], dtype=torch.int64)
    c = torch.tensor([p[1] for p in pairs], dtype=torch.int64)
    return R.Sp(r, c, None, (n, n))


def _pairs(s):
    return list(zip(s.row.tolist(), s.col.tolist()))


def test_utils_main_vectors():
    # utils.py:332-335 inputs; expected sets from SURVEY.md §8c-1
    adj1 = _sp_from_pairs([(0, 0), (0, 1), (1, 1), (2, 2), (3, 3)], 4)
    adj2 = _sp_from_pairs([(0, 0), (3, 1), (1, 1), (2, 2), (3, 3)], 4)
    assert _pairs(R.spmoverlap_(adj1, adj2)) == [(0, 0), (1, 1), (2, 2), (3, 3)]
    only1, only2 = R.spmnotoverlap_(adj1, adj2)
    assert _pairs(only1) == [(0, 1)] and _pairs(only2) == [(3, 1)]
    ov, o1, o2 = R.spmoverlap_notoverlap_(adj1, adj2)
    assert _pairs(ov) == [(0, 0), (1, 1), (2, 2), (3, 3)] and _pairs(o1) == [(0, 1)] and _pairs(o2) == [(3, 1)]


def _hand_graph():
    und = [(0, 2), (0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 3), (4, 5)]
    pairs = und + [(b, a) for a, b in und]
    return _sp_from_pairs(pairs, 6)


def _rows(s, B):
    out = [dict() for _ in range(B)]
    for r, c, v in zip(s.row.tolist(), s.col.tolist(), s.values().tolist()):
        out[r][c] = v
    return out


def test_hand_worked_cn5():
    # SURVEY.md §8c-3
    A = _hand_graph()
    e = torch.tensor([[0, 2, 0, 1], [1, 3, 5, 0]])
    cn1 = R.adjoverlap(A, A, e)
    assert [sorted(d) for d in _rows(cn1, 4)] == [[2, 3, 4], [0, 1], [4], [2, 3, 4]]
    a2 = R.adj2_true(A)
    cn2s = R.adjoverlap(A, a2, e)
    assert [sorted(d) for d in _rows(cn2s, 4)] == [[2, 3], [0, 1, 3], [], [2, 3]]
    w1, w2 = R.get_cn(A, e, 2)
    assert [sorted(d) for d in _rows(w1, 4)] == [[2, 3, 4], [0, 1], [4], [2, 3, 4]]
    rows2 = _rows(w2, 4)
    assert rows2[1] == {0: 1.0, 1: 1.0, 3: 3.0} and rows2[0] == {2: 1.0, 3: 1.0}
    x = torch.eye(6)
    # training, first call, structure-only CN2
    st = R.InnerProdState()
    xcn1, xcn2, xij, n1, n2 = R.cn5_aggregate(cn1, cn2s, x, e, st, training=True)
    assert st.innerprod.item() == pytest.approx(2.0) and st.n == 1
    r1 = _rows(n1, 4)
    assert r1[0] == pytest.approx({2: 0.5, 3: 0.5, 4: 1 / 3}) and r1[1] == {0: 0.0, 1: 0.0}
    r2 = _rows(n2, 4)
    assert r2[0] == pytest.approx({2: 0.5, 3: 1.0, 4: 1 / 3})
    assert r2[1] == pytest.approx({0: 1.0, 1: 1.0, 3: -1.0})
    assert r2[2] == pytest.approx({4: 1 / 3})
    assert torch.allclose(xcn2, n2.to_dense())
    # eval, untrained (ip = 0), pygho-weighted CN2
    st = R.InnerProdState()
    _, _, _, n1, n2 = R.cn5_aggregate(w1, w2, x, e, st, training=False)
    r2 = _rows(n2, 4)
    assert r2[0] == pytest.approx({2: 0.5, 3: 1 / 5, 4: 0.0})
    # structure-only eval vector of SURVEY §8c-3
    _, _, _, n1, n2 = R.cn5_aggregate(cn1, cn2s, x, e, R.InnerProdState(), training=False)
    r2 = _rows(n2, 4)
    assert r2[0] == pytest.approx({2: 0.5, 3: 1 / 3, 4: 0.0})
    assert r2[1] == pytest.approx({0: 1.0, 1: 1.0, 3: 1 / 3})
    assert r2[2] == pytest.approx({4: 0.0})
    # cn7 --sum 1
    xcn1, xcn2, _, n1 = R.cn7_aggregate(cn1, cn2s, x, e, 1.0)
    assert _rows(n1, 4)[1] == {0: 1.0, 1: 1.0}
    assert _rows(n1, 4)[0] == pytest.approx({2: 0.5, 3: 0.5, 4: 1 / 3})
    assert torch.equal(xcn2, cn2s.to_dense())


@pytest.mark.parametrize("seed,n,m", [(0, 40, 120), (1, 90, 500), (2, 150, 400)])
def test_oracle_vs_bruteforce_sets(seed, n, m):
    g = synth.tiny_graph(n, m, seed)
    A = R.sp_from_csr(g.rowptr, g.col)
    e = g.query_edges(64, "mixed")
    a = brute.dense_adj(g.rowptr.numpy(), g.col.numpy(), n)
    en = e.numpy()
    # order 1 via the searchsorted path and via python sets
    cn1 = R.adjoverlap(A, A, e)
    sets = brute.cn1_python_sets(g.rowptr.numpy(), g.col.numpy(), en)
    assert [sorted(d) for d in _rows(cn1, 64)] == sets
    # orders 1..3 with walk counts via the pygho-style path
    cns = R.get_cn(A, e, 3)
    for k in (1, 2, 3):
        ref = brute.cn_sets(a, en, k)
        rows = _rows(cns[k - 1], 64)
        for b in range(64):
            assert sorted(rows[b]) == ref[b][0].tolist()
            assert [rows[b][c] for c in sorted(rows[b])] == ref[b][1].astype(float).tolist()
    # structure of A^2 through adjoverlap(adj, adj2)
    cn2s = R.adjoverlap(A, R.adj2_true(A), e)
    assert [sorted(d) for d in _rows(cn2s, 64)] == [r[0].tolist() for r in brute.cn_sets(a, en, 2)]


@pytest.mark.parametrize("order,weighted,ip", [(2, True, 0.0), (2, False, 0.37), (3, True, 0.0), (3, True, 0.81)])
def test_oracle_vs_bruteforce_aggregate(order, weighted, ip):
    g = synth.tiny_graph(70, 300, 5)
    A = R.sp_from_csr(g.rowptr, g.col)
    e = g.query_edges(48, "mixed")
    x = g.features(8)
    a = brute.dense_adj(g.rowptr.numpy(), g.col.numpy(), g.n)
    cns = R.get_cn(A, e, order)
    if not weighted:
        cns = [R.Sp(c.row, c.col, torch.ones(c.nnz), c.shape) for c in cns]
    st = R.InnerProdState(ip)
    if order == 2:
        outs = R.cn5_aggregate(cns[0], cns[1], x, e, st, training=False)[:2]
    else:
        outs = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, st, training=False)[:3]
    ref = brute.cn5_dense(a, e.numpy(), x.double().numpy(), float(np.float32(ip)), order, weighted)
    for o, r in zip(outs, ref):
        scale = 1.0 + np.abs(r).max()
        assert np.abs(o.double().numpy() - r).max() <= 2e-4 * scale


def test_folded_adj2_matches_definition():
    g = synth.tiny_graph(50, 200, 7)
    A = R.sp_from_csr(g.rowptr, g.col)
    a = brute.dense_adj(g.rowptr.numpy(), g.col.numpy(), 50)
    f = R.adj2_folded(A, block_size=16).to_dense().numpy()
    full = a @ a
    exp = np.zeros((50, 50))
    for i in range(0, 50, 16):
        for j in range(0, 50, 16):
            blk = full[i:i + 16, j:j + 16]
            exp[:blk.shape[0], :blk.shape[1]] += blk
    assert np.array_equal(f, exp)
    # block_size >= n degenerates to the true A^2
    assert np.array_equal(R.adj2_folded(A, 64).to_dense().numpy(), full)


def test_metrics():
    pos = torch.tensor([0.9, 0.2, 0.5])
    neg = torch.tensor([0.1, 0.3, 0.6, 0.05])
    assert R.hits_at_k(pos, neg, 2) == pytest.approx(2 / 3)
    assert R.hits_at_k(pos, neg, 10) == 1.0
    m = R.mrr(torch.tensor([0.5, 0.1]), torch.tensor([[0.4, 0.6, 0.5], [0.0, 0.0, 0.0]]))
    assert m.tolist() == pytest.approx([1 / 2.5, 1.0])


def test_pure_conv_gcn_dense():
    g = synth.tiny_graph(30, 80, 9)
    A = R.sp_from_csr(g.rowptr, g.col)
    a = torch.tensor(brute.dense_adj(g.rowptr.numpy(), g.col.numpy(), 30), dtype=torch.float32)
    x = g.features(5)
    d = torch.rsqrt(1 + a.sum(1)).unsqueeze(1)
    assert torch.allclose(R.pure_conv(x, A, "gcn"), d * (a @ (d * x) + d * x), atol=1e-5)
    assert torch.allclose(R.pure_conv(x, A, "sum"), a @ x, atol=1e-5)
    ahat = a + torch.eye(30)
    dd = ahat.sum(1).pow(-0.5)
    assert torch.allclose(R.gcnconv_propagate(x, A, True, True), (dd[:, None] * ahat * dd[None, :]) @ x, atol=1e-5)
    assert torch.allclose(R.pure_conv3_gcn(x, A), (d * a * d.t()) @ x, atol=1e-5)


def test_adjoverlap_calresadj_and_masked_adjacency_vs_sets():
    """The oracle restatements of the steps either side of the path against python sets: adjoverlap(calresadj=True)
    (utils.py:260-274) and the --maskinput adjacency (NeighborOverlap_large.py:49-63)."""
    g = synth.tiny_graph(60, 300, 9)
    A = R.sp_from_csr(g.rowptr, g.col)
    nbr = [set(g.col[int(g.rowptr[v]):int(g.rowptr[v + 1])].tolist()) for v in range(g.n)]
    e = g.query_edges(40, "mixed")
    ov, r1, r2 = R.adjoverlap(A, A, e, calresadj=True)
    for b in range(e.shape[1]):
        i, j = int(e[0, b]), int(e[1, b])
        got = [set(s.col[s.row == b].tolist()) for s in (ov, r1, r2)]
        assert got == [nbr[i] & nbr[j], nbr[i] - nbr[j], nbr[j] - nbr[i]]
    # masked adjacency: duplicates and reversed copies of an edge keep it alive until every copy is masked
    el = torch.tensor([[0, 0, 1, 2, 3, 3], [1, 1, 0, 3, 2, 4]])
    full = R.masked_adjacency(el, 5)
    assert _pairs(full) == [(0, 1), (1, 0), (2, 3), (3, 2), (3, 4), (4, 3)]
    assert _pairs(R.masked_adjacency(el, 5, torch.tensor([0, 1]))) == _pairs(full)          # (1, 0) still lists the pair
    assert _pairs(R.masked_adjacency(el, 5, torch.tensor([0, 1, 2]))) == [(2, 3), (3, 2), (3, 4), (4, 3)]
    assert _pairs(R.masked_adjacency(el, 5, torch.tensor([5]), symmetric=False)) == [(0, 1), (1, 0), (2, 3), (3, 2)]


def test_spmm_max_backward_matches_autograd():
    """First-arg-max gradient routing of spmm_max (no ties in random data: equals torch's amax backward)."""
    g = synth.make_graph("cora")
    A = R.sp_from_csr(g.rowptr, g.col)
    x = g.features(6)
    w = torch.randn(g.n, 6, generator=torch.Generator().manual_seed(4))
    xr = x.clone().requires_grad_(True)
    out = torch.full((g.n, 6), float("-inf")).index_reduce(0, A.row, A.values().unsqueeze(1) * xr[A.col], "amax", include_self=True)
    out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
    (out * w).sum().backward()
    assert torch.allclose(R.spmm_max_backward(A, x, w), xr.grad, rtol=1e-5, atol=1e-5)
    # ties: the first maximum in row order takes the whole gradient
    xt = torch.zeros(g.n, 1)
    gx = R.spmm_max_backward(A, xt, torch.ones(g.n, 1))
    deg = g.rowptr[1:] - g.rowptr[:-1]
    first = g.col[g.rowptr[:-1][deg > 0]].long()
    ref = torch.zeros(g.n, 1).index_add_(0, first, torch.ones(first.numel(), 1))
    assert torch.equal(gx, ref)


def test_oracle_edge_cases_vs_bruteforce():
    """Empty batch, endpoints without neighbours, i == j, a link repeated inside one batch (its CN columns
    count twice in the column statistics, model.py:2261) -- the cases the driver loops can produce
    (utils.py:218 is the reference's own empty-input guard; negative sampling draws isolated nodes and repeats)."""
    n = 30
    g = synth.tiny_graph(n, 60, 11)
    rp, col = g.rowptr.numpy(), g.col.numpy()
    deg = np.diff(rp)
    A = R.sp_from_csr(g.rowptr, g.col)
    a = brute.dense_adj(rp, col, n)
    # empty batch: every CN matrix is [0 x n] with no entries, aggregates are [0 x F]
    e0 = torch.zeros(2, 0, dtype=torch.int64)
    cns = R.get_cn(A, e0, 3)
    assert [c.nnz for c in cns] == [0, 0, 0] and all(c.shape == (0, n) for c in cns)
    assert R.adjoverlap(A, A, e0).nnz == 0
    x = g.features(4)
    out = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e0, R.InnerProdState(), training=False)
    assert all(o.shape == (0, 4) for o in out[:4])
    # isolated endpoints (made by removing a node's row and column), self link, repeated link
    iso = int(np.argmax(deg))
    keep = (A.row != iso) & (A.col != iso)
    A2 = R.Sp(A.row[keep], A.col[keep], None, A.shape)
    a2 = a.copy()
    a2[iso, :] = 0
    a2[:, iso] = 0
    other = int(np.argsort(deg)[-2])
    e = torch.tensor([[iso, other, other, 3, 3, 5], [other, iso, other, 7, 7, 5]], dtype=torch.int64)
    cns = R.get_cn(A2, e, 3)
    for k in (1, 2, 3):
        ref = brute.cn_sets(a2, e.numpy(), k)
        rows = _rows(cns[k - 1], e.shape[1])
        for b in range(e.shape[1]):
            assert sorted(rows[b]) == ref[b][0].tolist()
            assert [rows[b][c] for c in sorted(rows[b])] == ref[b][1].astype(float).tolist()
        assert rows[0] == {} and rows[1] == {}          # no neighbours on one side: empty set at every order
        assert rows[3] == rows[4]                       # the repeated link
    # i == j at order 1 is the whole row
    assert sorted(_rows(cns[0], 6)[2]) == np.nonzero(a2[other])[0].tolist()
    outs = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, R.InnerProdState(0.25), training=False)[:3]
    ref = brute.cn5_dense(a2, e.numpy(), x.double().numpy(), float(np.float32(0.25)), 3, True)
    for o, r in zip(outs, ref):
        assert np.abs(o.double().numpy() - r).max() <= 2e-4 * (1.0 + np.abs(r).max())
