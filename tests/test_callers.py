"""the callers of the path -- SPD's shortest-path CN2, the ppa / citation2 scoring loops, the
training step.  CPU: the oracle's SPD restatement against python sets.  GPU: ocn_b200.callers against the oracle
and against the per-batch loop the drivers run."""
import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import callers, synth
from oracle import ref_ops as R

DEV = "cuda:0"


def _neighbours(g):
    rp, col = g.rowptr.tolist(), g.col.tolist()
    return [set(col[rp[i]:rp[i + 1]]) for i in range(g.n)]


def _links(g, B, seed):
    neg = torch.stack((synth.hash_randint(B - B // 2, g.n, 500 + seed, 1, "cpu"), synth.hash_randint(B - B // 2, g.n, 500 + seed, 2, "cpu")))
    return torch.cat((g.query_edges(B // 2, "pos"), neg), 1)


def test_oracle_spd_against_python_sets():
    g = synth.tiny_graph(70, 300, 5)
    N = _neighbours(g)
    e = _links(g, 40, 0)
    A = R.sp_from_csr(g.rowptr, g.col)
    cn1, cn2 = R.get_cn_spd(A, e)
    plain = R.get_cn(A, e, 2)
    assert torch.equal(cn1.row, plain[0].row) and torch.equal(cn1.col, plain[0].col)
    assert torch.equal(cn2.row, plain[1].row) and torch.equal(cn2.col, plain[1].col)      # the masked entries stay
    for b, k, v in zip(cn2.row.tolist(), cn2.col.tolist(), cn2.values().tolist()):
        i, j = int(e[0, b]), int(e[1, b])
        walks = len(N[j] & N[k])
        assert k in N[i] and walks > 0
        assert v == (0 if k in N[j] else walks)
    lit = R.get_cn_spd(A, e, literal=True)[1]
    for b, k, v in zip(lit.row.tolist(), lit.col.tolist(), lit.values().tolist()):
        j = int(e[1, b])
        assert v == (0 if k in N[b] else len(N[j] & N[k]))     # masked by NODE b: the reference's loop as written


@pytest.mark.gpu
@pytest.mark.parametrize("shape,scale,B", [("cora", 0.5, 512), ("citation2", 0.002, 1024)])
def test_spd_sets_and_fused_aggregate(shape, scale, B):
    g = synth.make_graph(shape, scale=scale)
    e = _links(g, B, 1) if shape == "cora" else g.query_edges(B, "stream")
    A = R.sp_from_csr(g.rowptr, g.col)
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    want1, want2 = R.get_cn_spd(A, e)
    got1, got2 = callers.get_cn1_cn2_spd(G, e.to(DEV))
    for got, want in ((got1, want1), (got2, want2)):
        assert torch.equal(got.rowptr.cpu(), want.rowptr()) and torch.equal(got.col.cpu(), want.col)
        assert torch.equal(got.value.cpu(), want.values())
    # the fused session with the flag: same aggregates as the oracle's cn5 fed with the masked matrices
    x = g.features(16)
    r = R.cn5_aggregate(want1, want2, x, e, R.InnerProdState(0.37), training=False)
    ip3 = torch.full((3,), 0.37, device=DEV)
    sess = ob.CNSession(G, e.to(DEV), None, 2).build(2, True, spd=True)
    sess.stats(5, 0.0, ip3, 0)
    xcn1, xcn2, _, xij = sess.aggregate(x.to(DEV), 5, 0.0, ip3)
    sess.release()
    for a, b in ((xcn1, r[0]), (xcn2, r[1]), (xij, r[2])):
        assert torch.allclose(a.cpu(), b, rtol=1e-4, atol=1e-4 * (1 + b.abs().max().item())), (a.cpu() - b).abs().max()
    # and the plain session differs (the flag does something)
    s2 = ob.CNSession(G, e.to(DEV), None, 2).build(2, True)
    s2.stats(5, 0.0, ip3, 0)
    p2 = s2.aggregate(x.to(DEV), 5, 0.0, ip3)[1]
    s2.release()
    assert not torch.allclose(p2, xcn2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cn5", "cn6", "cn7"])
def test_score_links_equals_the_per_batch_loop(name):
    import types
    g = synth.make_graph("citation2", scale=0.002)
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    torch.manual_seed(1)
    pred = ob.predictor_dict[name](32, 32, 1, 3, 0.0, weighted=True).to(DEV).eval()
    h = g.features(32).to(DEV)
    e = g.query_edges(5 * 256 + 77, "stream").to(DEV)          # a ragged last batch
    args = types.SimpleNamespace(sum=1.0)
    got = callers.score_links(pred, h, G, e, 256, args, batches_per_session=2)
    ref = []
    with torch.no_grad():
        for s in range(0, e.shape[1], 256):
            eb = e[:, s:s + 256].contiguous()
            if name == "cn6":
                ref.append(pred(h, G, None, None, None, eb, args).reshape(-1))
            elif name == "cn7":
                ref.append(pred.multidomainforward(h, G, None, None, eb, args).reshape(-1))
            else:
                ref.append(pred(h, G, None, None, eb).reshape(-1))
    assert torch.allclose(got, torch.cat(ref), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_ppa_and_citation2_test_loops_give_the_evaluator_numbers():
    g = synth.make_graph("collab", scale=0.01)
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    torch.manual_seed(2)
    pred = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0, weighted=True).to(DEV).eval()
    h = g.features(32).to(DEV)
    rnd = lambda m, s: torch.stack((synth.hash_randint(m, g.n, s, 1, "cpu"), synth.hash_randint(m, g.n, s, 2, "cpu"))).t().contiguous()
    split = {"valid": {"edge": g.query_edges(300, "pos").t().contiguous(), "edge_neg": rnd(700, 11)},
             "test": {"edge": g.query_edges(280, "pos").t().contiguous(), "edge_neg": rnd(650, 12)}}
    res = callers.test_ppa(pred, h, G, split, 128)
    sc = lambda ed: callers.score_links(pred, h, G, ed.t().contiguous().to(DEV), 128).cpu()
    pv, nv, pt, nt = sc(split["valid"]["edge"]), sc(split["valid"]["edge_neg"]), sc(split["test"]["edge"]), sc(split["test"]["edge_neg"])
    for K in (20, 50, 100):
        want = (R.hits_at_k(pv, nv, K), R.hits_at_k(pv, nv, K), R.hits_at_k(pt, nt, K))
        assert res[f"Hits@{K}"] == pytest.approx(want, abs=1e-6)
    S, K = 37, 50
    src = synth.hash_randint(S, g.n, 21, 1, "cpu")
    tgt = synth.hash_randint(S, g.n, 21, 2, "cpu")
    neg = synth.hash_randint(S * K, g.n, 22, 1, "cpu").view(S, K)
    mrr = callers.test_citation2_split(pred, h, G, src, tgt, neg, 256)
    sc256 = lambda ed: callers.score_links(pred, h, G, ed.to(DEV), 256).cpu()
    pos = sc256(torch.stack((src, tgt)))
    negs = sc256(torch.stack((src.view(-1, 1).repeat(1, K).view(-1), neg.reshape(-1)))).view(S, K)
    # NB: the two legs normalise different batches (256 links cut from different streams) exactly as the driver does
    assert float(mrr) == pytest.approx(float(R.mrr(pos, negs).mean()), abs=1e-6)


@pytest.mark.gpu
def test_train_step_equals_the_drivers_sequential_loop():
    import copy
    import torch.nn.functional as F
    g = synth.make_graph("cora", scale=0.5)
    rp, col = g.rowptr, g.col.long()
    row = torch.repeat_interleave(torch.arange(g.n), rp[1:] - rp[:-1])
    und = torch.stack((row, col))[:, row < col].to(DEV)                 # the training edge list, one entry per edge
    G = ob.Graph.from_edge_index(und, g.n, symmetric=True, with_multiplicity=True)
    torch.manual_seed(4)
    lin = torch.nn.Linear(16, 16).to(DEV)
    model = lambda x, adj: ob.pure_conv(lin(x), adj, "gcn")
    pred = ob.CNLinkPredictorOringin(16, 16, 1, 3, 0.0, weighted=True).to(DEV).train()
    seq_lin, seq_pred = copy.deepcopy(lin), copy.deepcopy(pred)
    x = g.features(16).to(DEV)
    perm = torch.randperm(und.shape[1], device=DEV)[:96]
    pos = und[:, perm]
    negs = torch.stack((synth.hash_randint(96, g.n, 31, 1, DEV), synth.hash_randint(96, g.n, 31, 2, DEV)))
    loss = callers.train_step(model, pred, x, G, pos, negs, 32, maskinput=True)
    # the driver's loop (NeighborOverlap_large_ppa.py:69-141) on the same operators, one sub-batch at a time
    adj = G.masked(pos)
    h0 = ob.pure_conv(seq_lin(x), adj, "gcn")
    h = h0.detach().requires_grad_(True)
    want = 0.0
    for edge, sign in ((pos, 1.0), (negs, -1.0)):
        for s in range(0, 96, 32):
            out = seq_pred.multidomainforward(h, adj, None, None, edge[:, s:s + 32].contiguous())
            l = -(1 / 96) * F.logsigmoid(sign * out).sum()
            l.backward()
            want += l.item()
    h0.backward(h.grad)
    assert float(loss) == pytest.approx(want, rel=1e-5)
    assert torch.allclose(pred.innerprod, seq_pred.innerprod, rtol=1e-5) and pred.n == seq_pred.n == 6
    for (n1, p1), (_, p2) in zip(pred.named_parameters(), seq_pred.named_parameters()):
        if p2.grad is not None:
            assert torch.allclose(p1.grad, p2.grad, rtol=1e-3, atol=1e-6), n1
    assert torch.allclose(lin.weight.grad, seq_lin.weight.grad, rtol=1e-3, atol=1e-6)
