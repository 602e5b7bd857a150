"""cn2 (IncompleteCN1Predictor) and its sampler on the GPU against the oracle restatement at sizes the
reference-executed fixtures do not reach (tests/test_golden_reference.py holds the fixture checks)."""
import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import synth
from ocn_b200.cn import SparseRows
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _shared_draws():
    """One sequence of uniform draws served to both implementations (generated on the CPU, moved on demand)."""
    g = torch.Generator().manual_seed(123)
    log = []

    def record(shape, device=None):
        d = torch.rand(shape, generator=g)
        log.append(d)
        return d if device is None else d.to(device)

    def replay():
        it = iter(list(log))
        return lambda shape, device=None: (lambda d: d if device is None else d.to(device))(next(it))
    return record, replay


@pytest.mark.parametrize("deg", [1, 3, 8, 200])
def test_sparsesample_reweight_matches_the_oracle(deg):
    g = synth.make_graph("cora", scale=0.3)
    A = R.sp_from_csr(g.rowptr, g.col)
    e = g.query_edges(300, "pos")
    _, res1, _ = R.adjoverlap(A, A, e, calresadj=True)
    record, replay = _shared_draws()
    want = R.sparsesample_reweight(res1, deg, record)
    rows = SparseRows(res1.rowptr().to(DEV), res1.col.to(DEV), None, res1.shape)
    got = ob.sparsesample_reweight(rows, deg, replay())
    assert torch.equal(got.rowptr.cpu(), want.rowptr())
    assert torch.equal(got.col.cpu(), want.col)
    assert torch.allclose(got.value.cpu(), want.values(), rtol=1e-6, atol=0)
    # every sampled row keeps its total mass: deg draws of rowcount / deg
    cnt = (res1.rowptr()[1:] - res1.rowptr()[:-1]).float()
    mass = torch.zeros(res1.shape[0]).index_add_(0, want.row, want.values())
    assert torch.allclose(mass, cnt, rtol=1e-5)


@pytest.mark.parametrize("mode,resdeg", [("eval", 16), ("eval", 4), ("train", 4)])
def test_cn2_forward_matches_the_oracle(mode, resdeg, learnablept=False):
    g = synth.make_graph("cora", scale=0.5)
    torch.manual_seed(3)
    pred = ob.IncompleteCN1Predictor(64, 64, 1, 3, 0.0, trainresdeg=resdeg, testresdeg=resdeg, learnablept=learnablept)
    pred.train() if mode == "train" else pred.eval()
    x = g.features(64)
    A = R.sp_from_csr(g.rowptr, g.col)
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    state = R.InnerProdState()
    import copy
    cpu_mod = copy.deepcopy(pred)
    pred = pred.to(DEV)
    with torch.no_grad():
        for s in range(2):
            neg = torch.stack((synth.hash_randint(128, g.n, 900 + s, 1, "cpu"), synth.hash_randint(128, g.n, 900 + s, 2, "cpu")))
            e = torch.cat((g.query_edges(128, "pos"), neg), 1)
            record, replay = _shared_draws()
            want = R.cn2_forward(cpu_mod, x, A, e, state, mode == "train", 1, record)
            pred.rand_fn = replay()
            got = pred(x.to(DEV), G, e.to(DEV))
            assert got.shape == want.shape == (256, 1)
            assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4), (got.cpu() - want).abs().max()
            assert torch.allclose(pred.innerprod.cpu(), state.innerprod, rtol=1e-4, atol=1e-6)


def test_cn2_learnablept_is_refused():
    with pytest.raises(NotImplementedError):
        ob.IncompleteCN1Predictor(64, 64, 1, 3, 0.0, learnablept=True)


def test_cn2_trains_through_the_fused_operators():
    """Gradients reach x and the heads (the sparse weights carry none: they are scores computed under no_grad)."""
    g = synth.tiny_graph(80, 400, 2)
    torch.manual_seed(0)
    pred = ob.IncompleteCN1Predictor(64, 64, 1, 3, 0.0, trainresdeg=4).to(DEV).train()
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    x = g.features(64).to(DEV).requires_grad_(True)
    e = g.query_edges(32, "pos").to(DEV)
    pred(x, G, e).sum().backward()
    assert x.grad is not None and bool(torch.isfinite(x.grad).all()) and float(x.grad.abs().sum()) > 0
    assert all(p.grad is not None for n, p in pred.named_parameters() if n.startswith(("xcnlin", "xijlin", "lin.")))
