"""Golden fixtures produced by EXECUTING the reference's own model.py / utils.py / get_cn1_cn2
(oracle/make_golden.py, on a pure-torch stand-in of torch_sparse / pygho).

* CPU: the oracle restatement (oracle/ref_ops.py) must reproduce them  -> pins the oracle.
* GPU: the CUDA path (through the C ABI and the predictor mirror, loading the reference's
  state_dict) must reproduce them                                       -> parity with the reference code.
"""
import glob
import os

import pytest
import torch

import ocn_b200 as ob
from oracle import ref_ops as R

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_model_*.pt")))
assert len(GOLDEN) >= 9


def _load(path):
    return torch.load(path, weights_only=False)


def _sp(d):
    return R.Sp(d["row"], d["col"], d["val"], tuple(d["shape"]))


def _same_sparse(a: R.Sp, d, values=True):
    assert torch.equal(a.row, d["row"]) and torch.equal(a.col, d["col"])
    if values:
        assert torch.equal(a.values(), d["val"] if d["val"] is not None else torch.ones(a.nnz))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-3] for p in GOLDEN])
def test_oracle_reproduces_reference_execution(path):
    fx = _load(path)
    A = R.sp_from_csr(fx["rowptr"], fx["col"])
    x = fx["x"]
    st = R.InnerProdState()
    training = fx["mode"] == "train"
    a2 = R.adj2_true(A) if fx["style"] == "large" else None
    for call in fx["calls"]:
        e = call["edges"]
        if fx["style"] == "large":
            cn1, cn2 = R.adjoverlap(A, A, e), R.adjoverlap(A, a2, e)
            cns = [cn1, cn2]
        else:
            cns = R.get_cn(A, e, 3 if fx["predictor"] == "cn6" else 2)
        _same_sparse(cns[0], call["cn1"])
        _same_sparse(cns[1], call["cn2"])
        if fx["predictor"] == "cn6":
            _same_sparse(cns[2], call["cn3"])
            out = R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, st, training)
            assert torch.allclose(out[2], call["xcn3lin"], rtol=1e-5, atol=1e-6)
        elif fx["predictor"] == "cn7":
            out = R.cn7_aggregate(cns[0], cns[1], x, e, fx["fill"])
        else:
            out = R.cn5_aggregate(cns[0], cns[1], x, e, st, training)
        assert torch.allclose(out[0], call["xcn1lin"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(out[1], call["xcn2lin"], rtol=1e-5, atol=1e-6)
        assert torch.equal(x[e[0]] * x[e[1]], call["xijlin"])
        if fx["predictor"] != "cn7":
            assert torch.allclose(st.innerprod, call["innerprod"], rtol=1e-6, atol=1e-7) and st.n == call["n"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-3] for p in GOLDEN])
def test_cuda_path_reproduces_reference_execution(path):
    fx = _load(path)
    dev = "cuda:0"
    G = ob.Graph(fx["rowptr"].to(dev), fx["col"].to(dev), fx["n"])
    x = fx["x"].to(dev)
    F = fx["F"]
    cls = {"cn5": ob.CNLinkPredictorOringin, "cn6": ob.CNLinkPredictor3hopCNs, "cn7": ob.CNLinkPredictorbaselearn}[fx["predictor"]]
    pred = cls(F, F, 1, 3, 0.0, ln=fx["ln"], weighted=(fx["style"] == "pygho")).to(dev)
    pred.load_state_dict({k: v.to(dev) for k, v in fx["state_dict"].items()})
    pred.train() if fx["mode"] == "train" else pred.eval()
    order = 3 if fx["predictor"] == "cn6" else 2

    class Args:
        sum = fx["fill"]

    for call in fx["calls"]:
        e = call["edges"].to(dev)
        # CN index sets / values, bit-exact
        got = ob.get_cn(G, e, order, weighted=(fx["style"] == "pygho"))
        for k in range(order):
            ref = call[f"cn{k + 1}"]
            rp = torch.zeros(e.shape[1] + 1, dtype=torch.long)
            torch.cumsum(torch.bincount(ref["row"], minlength=e.shape[1]), 0, out=rp[1:])
            assert torch.equal(got[k].rowptr.cpu(), rp) and torch.equal(got[k].col.cpu(), ref["col"])
            assert torch.equal(got[k].value.cpu(), ref["val"] if ref["val"] is not None else torch.ones(ref["row"].numel()))
        with torch.no_grad():
            fill = float(fx["fill"] or 0.0) if fx["predictor"] == "cn7" else 0.0
            xcn1, xcn2, xcn3, xij, _ = pred.cn_stage(x, G, e, fill)
            scale = lambda t: 1.0 + t.abs().max().item()
            assert (xcn1.cpu() - call["xcn1lin"]).abs().max().item() <= 2e-5 * scale(call["xcn1lin"])
            assert (xcn2.cpu() - call["xcn2lin"]).abs().max().item() <= 1e-4 * scale(call["xcn2lin"])
            if order == 3:
                assert (xcn3.cpu() - call["xcn3lin"]).abs().max().item() <= 1e-4 * scale(call["xcn3lin"])
            assert torch.equal(xij.cpu(), call["xijlin"])
            out = pred._head(xcn1, xcn2, xcn3, xij)
            assert torch.allclose(out.cpu(), call["out"], rtol=1e-3, atol=1e-4)
        if fx["predictor"] != "cn7":
            assert pred.n == call["n"]
            assert abs(pred.innerprod.item() - call["innerprod"].item()) <= 1e-4 * (1 + abs(call["innerprod"].item()))


UTILS_FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "ref_utils_adj2byblock_calresadj.pt")


def test_oracle_reproduces_reference_adj2byblock_and_calresadj():
    """The reference's own ``sparse_tensor_multiply`` (--adj2byblock, utils.py:287-329), the ``adjoverlap(adj, adj2,
    edge)`` that follows it in the driver (NeighborOverlap_large.py:79) and ``adjoverlap(..., calresadj=True)``
    (utils.py:260-274), executed by oracle/make_golden.py: pins ``adj2_folded`` (every block product lands
    in the top-left corner), and the residual sets.  The CUDA path is compared with the same oracle functions in
    tests/test_gpu_parity.py::test_adjoverlap_generic_and_spgemm."""
    cases = _load(UTILS_FIXTURE)
    assert len(cases) >= 4
    for c in cases:
        A = R.sp_from_csr(c["rowptr"], c["col"])
        f = R.adj2_folded(A, c["block"])
        want = c["adj2"]
        assert tuple(want["shape"]) == (c["n"], c["n"])
        assert torch.equal(f.row, want["row"]) and torch.equal(f.col, want["col"]), c["name"]
        assert torch.equal(f.values(), want["val"].float()), c["name"]
        if c["block"] < c["n"]:   # the fold is visible: nothing outside the top-left block
            assert int(f.row.max()) < c["block"] and int(f.col.max()) < c["block"]
        e = c["edges"]
        _same_sparse(R.adjoverlap(A, f, e), c["cn2"])
        for got, key in zip(R.adjoverlap(A, A, e, calresadj=True), ("overlap", "res1", "res2")):
            _same_sparse(got, c[key])


# ---- cn2 = IncompleteCN1Predictor: fixtures produced by the reference's own class ----------------------

CN2_FIXTURES = ["cn2_eval_cora", "cn2_train_tiny", "cn4_train_tiny", "cn3_train_tiny", "cn3_eval_tiny"]


def _replay_draws(draws):
    it = iter(draws)

    def rand_fn(shape, device=None):
        d = next(it)
        assert tuple(d.shape) == tuple(shape), f"sampler asked for {tuple(shape)}, the reference drew {tuple(d.shape)}"
        return d if device is None else d.to(device)
    return rand_fn


def _cn2_module(fx, device="cpu"):
    from ocn_b200 import completion
    cls = getattr(completion, fx.get("cls", "IncompleteCN1Predictor"))
    pred = cls(64, 64, 1, 3, 0.0, trainresdeg=fx["trainresdeg"], testresdeg=fx["testresdeg"], depth=1)
    missing = pred.load_state_dict(fx["state_dict"], strict=True)     # same parameter / buffer names as the reference
    assert not missing.missing_keys and not missing.unexpected_keys
    pred = pred.to(device)
    pred.train() if fx["mode"] == "train" else pred.eval()
    return pred


@pytest.mark.parametrize("name", CN2_FIXTURES)
def test_oracle_completion_predictors_match_the_reference_classes(name):
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", f"ref_{name}.pt"))
    mod = _cn2_module(fx)
    A = R.sp_from_csr(fx["rowptr"], fx["col"])
    state = R.InnerProdState()
    with torch.no_grad():
        for call in fx["calls"]:
            if fx.get("cls", "").endswith("highorder"):
                out = R.cn3_forward(mod, fx["x"], A, R.adj2_true(A), call["edges"], state, fx["mode"] == "train", 1,
                                    _replay_draws(call["draws"]))
            else:
                out = R.cn2_forward(mod, fx["x"], A, call["edges"], state, fx["mode"] == "train", 1, _replay_draws(call["draws"]),
                                    fill=mod.residual_fill, xij_passes=mod.xij_passes)
            assert torch.allclose(out, call["out"], rtol=1e-4, atol=1e-5), (out - call["out"]).abs().max()
            assert torch.allclose(state.innerprod, call["innerprod"], rtol=1e-5, atol=1e-6)
            assert state.n == call["n"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CN2_FIXTURES)
def test_cuda_completion_predictors_match_the_reference_classes(name):
    import ocn_b200 as ob
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", f"ref_{name}.pt"))
    dev = "cuda:0"
    pred = _cn2_module(fx, dev)
    G = ob.Graph(fx["rowptr"].to(dev), fx["col"].to(dev), fx["n"])
    x = fx["x"].to(dev)
    with torch.no_grad():
        for call in fx["calls"]:
            pred.rand_fn = _replay_draws(call["draws"])
            e = call["edges"].to(dev)
            out = pred(x, G, None, None, e) if fx.get("cls", "").endswith("highorder") else pred(x, G, e)
            assert torch.allclose(out.cpu(), call["out"], rtol=1e-4, atol=1e-5), (out.cpu() - call["out"]).abs().max()
            assert torch.allclose(pred.innerprod.cpu(), call["innerprod"], rtol=1e-5, atol=1e-6)
            assert pred.n == call["n"]
