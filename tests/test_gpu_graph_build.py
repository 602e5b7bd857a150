"""The step before the path: device CSR build from an edge list and the per-batch
--maskinput adjacency, bit-exact against the oracle's restatement of
``SparseTensor.from_edge_index(tei).to_symmetric()`` (NeighborOverlap_large.py:56-63)."""
import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _same(G: ob.Graph, S: R.Sp):
    assert torch.equal(G.rowptr.cpu(), S.rowptr()), "row pointers differ"
    assert torch.equal(G.col.cpu().long(), S.col), "columns differ"


def _edge_list(name, dup=True):
    g = synth.make_graph(name) if name in ("cora", "pubmed") else synth.tiny_graph(70, 400, 5)
    el = torch.stack((g.raw_src, g.raw_dst)).cpu()
    if dup:  # duplicates, both directions of some links and a few self loops, as a raw ogb edge list may hold
        k = el.shape[1] // 5
        el = torch.cat((el, el[:, :k], el[:, k:2 * k].flip(0), torch.tensor([[3, 9], [3, 9]])), dim=1)
    return g.n, el


@pytest.mark.parametrize("name", ["tiny", "cora", "pubmed"])
@pytest.mark.parametrize("symmetric", [True, False])
def test_device_build_matches_reference_construction(name, symmetric):
    n, el = _edge_list(name)
    G = ob.Graph.from_edge_index(el.to(DEV), n, symmetric=symmetric, with_multiplicity=True)
    S = R.masked_adjacency(el, n, None, symmetric)
    _same(G, S)
    assert int(G.mult.sum()) == el.shape[1] * (2 if symmetric else 1)
    if symmetric:
        assert G.validate() == 0
    # keep mask = the reference's adjmask
    keep = torch.rand(el.shape[1], generator=torch.Generator().manual_seed(1)) < 0.7
    Gk = ob.Graph.from_edge_index(el.to(DEV), n, symmetric=symmetric, keep=keep.to(DEV))
    _same(Gk, R.masked_adjacency(el, n, torch.nonzero(~keep).squeeze(1), symmetric))
    with pytest.raises(ValueError):
        ob.Graph.from_edge_index(torch.tensor([[0, n], [1, 2]], device=DEV), n)


def test_empty_and_degenerate_edge_lists():
    G = ob.Graph.from_edge_index(torch.zeros(2, 0, dtype=torch.int64, device=DEV), 17)
    assert G.nnz == 0 and bool((G.rowptr == 0).all())
    el = torch.tensor([[2, 2, 5], [5, 5, 2]])
    G = ob.Graph.from_edge_index(el.to(DEV), 8, with_multiplicity=True)
    assert G.col.tolist() == [5, 2] and G.mult.tolist() == [3, 3]
    Gm = G.masked(el[:, :2].to(DEV))
    assert Gm.col.tolist() == [5, 2] and Gm.mult.tolist() == [1, 1]
    Gm = G.masked(el.to(DEV))
    assert Gm.nnz == 0


@pytest.mark.parametrize("name", ["tiny", "cora", "pubmed"])
@pytest.mark.parametrize("batch", [1, 64, 1152])
def test_masked_adjacency_equals_rebuild(name, batch):
    """adjmask[perm] = 0 -> rebuild == multiplicity decrement + row compaction, for consecutive batches
    of one permutation (the work array must come back clean every time)."""
    n, el = _edge_list(name)
    G = ob.Graph.from_edge_index(el.to(DEV), n, with_multiplicity=True)
    perm_all = torch.randperm(el.shape[1], generator=torch.Generator().manual_seed(7))
    for k in range(3):
        perm = perm_all[k * batch:(k + 1) * batch]
        if perm.numel() == 0:
            break
        Gm = G.masked(el[:, perm].to(DEV))
        _same(Gm, R.masked_adjacency(el, n, perm))
        assert Gm.validate() == 0
    decs = [v for k, v in G._ws.items() if isinstance(k, tuple) and k[0] == "mask_dec"]
    assert decs and all(bool((v == 0).all()) for v in decs)
    with pytest.raises(ValueError):
        G.masked(torch.tensor([[0], [0]], device=DEV) if not bool(((el[0] == 0) & (el[1] == 0)).any())
                 else torch.tensor([[n - 1], [n - 1]], device=DEV))


def test_masked_graph_feeds_the_cn_path():
    """CN sets of a training batch on the masked adjacency (the target links themselves are gone)."""
    n, el = _edge_list("cora", dup=False)
    G = ob.Graph.from_edge_index(el.to(DEV), n, with_multiplicity=True)
    perm = torch.randperm(el.shape[1], generator=torch.Generator().manual_seed(2))[:256]
    e = el[:, perm]
    Gm = G.masked(e.to(DEV))
    A = R.masked_adjacency(el, n, perm)
    got = ob.get_cn(Gm, e.to(DEV), 3, True)
    ref = R.get_cn(A, e, 3)
    for k in range(3):
        assert torch.equal(got[k].rowptr.cpu(), ref[k].rowptr())
        assert torch.equal(got[k].col.cpu(), ref[k].col)
        assert torch.equal(got[k].value.cpu(), ref[k].values())
