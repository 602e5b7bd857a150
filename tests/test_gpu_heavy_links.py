"""Links whose source has more than 1024 neighbours (kHeavyLink) are walked by a whole CTA in the per-link
kernels (column statistics, batch statistics, aggregation forward / backward, release) and, for orders 1-2 on
training-shape batches, by many warps of the record-flattened direct kernel.  A graph with two hub nodes makes
those paths run at test size; everything is compared with the oracle exactly as in test_gpu_parity.py."""
import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _hub_graph():
    n = 2600
    base = synth.tiny_graph(n, 9000, 11)
    el = torch.stack((base.raw_src, base.raw_dst)).cpu()
    gen = torch.Generator().manual_seed(5)
    h0 = torch.randperm(n, generator=gen)[:1500]   # hub 0: 1500 neighbours
    h1 = torch.randperm(n, generator=gen)[:1100]   # hub 1: 1100 neighbours, overlapping
    star = torch.cat((torch.stack((torch.zeros_like(h0), h0)), torch.stack((torch.ones_like(h1), h1))), 1)
    el = torch.cat((el, star), 1)
    el = el[:, el[0] != el[1]]
    A = R.masked_adjacency(el, n)
    G = ob.Graph(A.rowptr().to(DEV), A.col.to(DEV), n)
    assert G.validate() == 0
    deg = G.degree()
    assert int(deg[0]) > 1024 and int(deg[1]) > 1024
    # links: hub sources against assorted destinations (incl. the other hub), light sources, and hub destinations
    d = torch.randint(0, n, (40,), generator=gen)
    e = torch.cat((torch.stack((torch.zeros(20, dtype=torch.long), d[:20])), torch.stack((torch.ones(20, dtype=torch.long), d[20:])),
                   torch.tensor([[0, 1, 7, 9], [1, 0, 0, 1]]), torch.stack((d[:30], d[10:40]))), 1)
    return G, A, e, n


def _close(got, ref, mass, rtol):
    err = (got.cpu().double() - ref.double()).abs()
    bound = rtol * (1.0 + mass.double())
    assert bool((err <= bound).all()), f"max err {err.max().item():.3e}, worst bound ratio {(err / bound).max().item():.2f}"


def _mass(sp, x):
    return R.spmm_add(R.Sp(sp.row, sp.col, sp.values().abs(), sp.shape), x.abs())


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("ip", [0.0, 0.37])
def test_heavy_links_forward_stats_release(order, ip):
    G, A, e, n = _hub_graph()
    F = 32
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(1))
    cns = R.get_cn(A, e, order)
    ip3 = torch.full((3,), ip, dtype=torch.float32, device=DEV)
    sess = ob.CNSession(G, e.to(DEV), None, order, hub_degree=-1 if order == 2 else 0).build(order, True)
    for k in range(order):  # CN sets bit-exact (order 2: record-flattened direct kernel; order 3: indexed / table path)
        got = sess.extract(k + 1)
        assert torch.equal(got.rowptr.cpu(), cns[k].rowptr()) and torch.equal(got.col.cpu(), cns[k].col)
        assert torch.equal(got.value.cpu(), cns[k].values())
    bs = sess.stats(5, 0.0, ip3, 0)
    if order == 2:
        r1, r2, rij, n1, n2 = R.cn5_aggregate(cns[0], cns[1], x, e, R.InnerProdState(ip), training=False)
        refs, ns = [r1, r2], [n1, n2]
    else:
        r1, r2, r3, rij, n1, n2, n3 = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, R.InnerProdState(ip), training=False)
        refs, ns = [r1, r2, r3], [n1, n2, n3]
    outs = sess.aggregate(x.to(DEV), 5, 0.0, ip3)
    for got, ref, nm in zip(outs[:order], refs, ns):
        _close(got, ref, _mass(nm, x), 2e-4 if ip else 2e-5)
    assert torch.equal(outs[3].cpu(), rij)
    # the batch inner product s = sum(C2 * C1-hat) (model.py:2244), summed per link then per batch
    s_ref = float((R.spsphadamard(cns[1], n1)).values().sum())
    assert abs(float(bs[0, 1]) - s_ref) <= 1e-4 * (1 + abs(s_ref))
    # run-to-run determinism of the CTA-per-link path
    outs2 = sess.aggregate(x.to(DEV), 5, 0.0, ip3)
    for a, b in zip(outs, outs2):
        assert a is None or torch.equal(a, b)
    sess.release()
    assert int(sess.colstat.count_nonzero()) == 0


def test_heavy_links_backward():
    G, A, e, n = _hub_graph()
    F = 16
    torch.manual_seed(0)
    pred = ob.CNLinkPredictorOringin(F, F, 1, 3, 0.0, weighted=True).to(DEV).train()
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    cns = R.get_cn(A, e, 2)
    st = R.InnerProdState()
    r1, r2, rij, n1, n2 = R.cn5_aggregate(cns[0], cns[1], x, e, st, training=True)
    xd = x.detach().to(DEV).requires_grad_(True)
    x1, x2, _, xij, _ = pred.cn_stage(xd, G, e.to(DEV))
    assert abs(pred.innerprod.item() - st.innerprod.item()) <= 1e-4 * (1 + abs(st.innerprod.item()))
    wts = [torch.randn_like(o) for o in (r1, r2, rij)]
    sum((o * w).sum() for o, w in zip((r1, r2, rij), wts)).backward()
    sum((o * w.to(DEV)).sum() for o, w in zip((x1, x2, xij), wts)).backward()
    gm = x.grad.abs().max().item()
    assert (xd.grad.cpu() - x.grad).abs().max().item() <= 5e-4 * (1 + gm)
