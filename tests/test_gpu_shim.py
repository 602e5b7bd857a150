"""The CUDA import shim (ocn_b200/shim) against the recorded API traces of the reference's own text
(tests/golden/ref_trace_*.pt: every torch_sparse / pygho call that utils.adjoverlap, get_cn1_cn2 and the cn5 / cn6 /
cn7 multidomainforward bodies made, with arguments and results, produced by oracle/make_trace.py), against the
reference-executed fixtures of utils.adjoverlap, and -- where the reference's files are present -- under the
reference's unmodified model.py / utils.py."""
import glob
import importlib
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

from oracle import trace as T  # noqa: E402
import shim_replay  # noqa: E402

pytestmark = pytest.mark.gpu
TRACES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_trace_*.pt")))
DEV = "cuda:0"


def _shim():
    return [importlib.import_module("ocn_b200.shim." + n) for n in
            ("torch_sparse", "pygho", "pygho.backend.Spspmm", "pygho.backend.Spmm")]


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p)[10:-3] for p in TRACES])
def test_replay_of_the_reference_calls_on_the_cuda_shim(path):
    ts, pg, spspmm_mod, spmm_mod = _shim()
    fx = torch.load(path)
    before = dict(spspmm_mod.FUSED)
    n = T.replay(fx["calls"], shim_replay.api_table(ts, pg, spspmm_mod, spmm_mod), torch.device(DEV), shim_replay.check)
    assert n == len(fx["calls"])
    if fx.get("style") == "pygho":
        # get_cn1_cn2's text went through the fused kernels: Ej.A was never materialised
        assert spspmm_mod.FUSED["cn1"] > before["cn1"] and spspmm_mod.FUSED["cn_order"] > before["cn_order"]


def test_accelerated_adjoverlap_matches_the_reference_fixture():
    import ocn_b200.shim as shim
    ts = importlib.import_module("ocn_b200.shim.torch_sparse")
    for case in torch.load(os.path.join(HERE, "golden", "ref_utils_adj2byblock_calresadj.pt")):
        n = case["n"]
        adj = ts.SparseTensor(rowptr=case["rowptr"].to(DEV), col=case["col"].to(DEV), sparse_sizes=(n, n), is_sorted=True)
        a2 = case["adj2"]
        adj2 = ts.SparseTensor(row=a2["row"].to(DEV), col=a2["col"].to(DEV), sparse_sizes=a2["shape"], is_sorted=True)
        e = case["edges"].to(DEV)
        # the folded adj2byblock matrix has fewer rows than n: only links whose destination is inside it can be asked
        ok = e[1] < a2["shape"][0]
        if a2["shape"][1] == n and bool(ok.all()):
            got = shim.adjoverlap(adj, adj2, e)
            r, c, v = got.coo()
            assert torch.equal(r.cpu(), case["cn2"]["row"]) and torch.equal(c.cpu(), case["cn2"]["col"])
        ov, r1, r2 = shim.adjoverlap(adj, adj, e, calresadj=True)
        for got, key in ((ov, "overlap"), (r1, "res1"), (r2, "res2")):
            r, c, v = got.coo()
            assert torch.equal(r.cpu(), case[key]["row"]), (case["name"], key)
            assert torch.equal(c.cpu(), case[key]["col"]), (case["name"], key)
            assert got.sizes() == list(case[key]["shape"])
            assert v.dtype == torch.float32 and bool((v == 1).all())


def test_gcnconv_shim_matches_the_dense_formula_forward_and_backward():
    from ocn_b200 import synth
    tg = importlib.import_module("ocn_b200.shim.torch_geometric.nn")
    ts = importlib.import_module("ocn_b200.shim.torch_sparse")
    g = synth.tiny_graph(50, 200, 9)
    n = g.n
    adj = ts.SparseTensor(rowptr=g.rowptr.to(DEV), col=g.col.to(DEV), sparse_sizes=(n, n), is_sorted=True)
    A = adj.to_dense().double()
    torch.manual_seed(0)
    x = torch.randn(n, 6, device=DEV)
    for kw, dense in (({}, "gcn"), ({"aggr": "sum", "normalize": False, "add_self_loops": False}, "sum"),
                      ({"aggr": "mean", "normalize": False, "add_self_loops": False}, "mean")):
        conv = tg.GCNConv(6, 5, **kw).to(DEV)
        assert sorted(conv.state_dict()) == ["bias", "lin.weight"]
        with torch.no_grad():
            conv.bias.normal_()
        xs = x.clone().requires_grad_(True)
        out = conv(xs, adj)
        out.square().sum().backward()
        xd = x.double().clone().requires_grad_(True)
        h = xd @ conv.lin.weight.double().t()
        if dense == "gcn":
            Ah = A + torch.eye(n, device=DEV, dtype=torch.double)
            d = Ah.sum(1).rsqrt()
            ref = d[:, None] * (Ah @ (d[:, None] * h))
        elif dense == "sum":
            ref = A @ h
        else:
            ref = (A @ h) / A.sum(1).clamp(min=1)[:, None]
        ref = ref + conv.bias.double()
        ref.square().sum().backward()
        assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(xs.grad.double(), xd.grad, rtol=1e-4, atol=1e-4)


def test_cooview_matmul_runs_through_the_library_and_reads_like_a_coo_tensor():
    pg = importlib.import_module("ocn_b200.shim.pygho")
    from ocn_b200 import _lib, synth
    g = synth.tiny_graph(40, 150, 4)
    n = g.n
    row = torch.repeat_interleave(torch.arange(n), g.rowptr[1:] - g.rowptr[:-1])
    ind = torch.stack((row, g.col.long())).to(DEV)
    val = torch.rand(ind.shape[1], device=DEV)
    A = pg.SparseTensor(ind, val, (n, n), is_coalesced=True)
    x = torch.randn(n, 7, device=DEV, requires_grad=True)
    view = A.to_torch_sparse_coo()
    launches = _lib.lib().ocn_launch_count()
    y = view @ x                                          # PureConv3 'sum' / 'gcn' (model.py:134-140)
    assert _lib.lib().ocn_launch_count() > launches
    ref = torch.sparse_coo_tensor(ind, val, (n, n)).to_dense() @ x.detach()
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5)
    y.sum().backward()
    assert torch.allclose(x.grad, torch.sparse_coo_tensor(ind, val, (n, n)).to_dense().t() @ torch.ones(n, 7, device=DEV),
                          rtol=1e-5, atol=1e-5)
    assert torch.equal(view.indices(), ind) and view.shape == (n, n) and view.coalesce() is view
    assert torch.allclose(view.to_dense(), torch.sparse_coo_tensor(ind, val, (n, n)).to_dense())


REF = os.environ.get("OCN_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model.py")),
                    reason="the reference's files are not on this machine (they never travel to the GPU boxes)")
def test_reference_text_runs_unmodified_on_the_shim():
    """Where a maintainer has the reference checked out next to a GPU: its model.py / utils.py, imported unmodified
    over shim.install(), reproduce the golden scores (tests/golden/ref_model_*.pt)."""
    import types
    import ocn_b200.shim as shim
    shim.install(force=True)
    sys.path.insert(0, REF)
    for m in ("model", "utils"):
        sys.modules.pop(m, None)
    try:
        import model as ref_model
        import utils as ref_utils
        import torch_sparse
        shim.accelerate(ref_utils, ref_model)
        for name in ("cn5_large_eval_tiny", "cn7_large_sum1_cora"):
            fx = torch.load(os.path.join(HERE, "golden", f"ref_model_{name}.pt"))
            n = fx["n"]
            adj = torch_sparse.SparseTensor(rowptr=fx["rowptr"].to(DEV), col=fx["col"].to(DEV), sparse_sizes=(n, n), is_sorted=True)
            adj2 = shim.a2(adj)
            cls = {"cn5": ref_model.CNLinkPredictorOringin, "cn7": ref_model.CNLinkPredictorbaselearn}[fx["predictor"]]
            pred = cls(fx["F"], fx["F"], 1, 3, 0.0, ln=fx["ln"]).to(DEV)
            pred.load_state_dict(fx["state_dict"])
            pred.eval()
            x = fx["x"].to(DEV)
            with torch.no_grad():
                for call in fx["calls"]:
                    e = call["edges"].to(DEV)
                    cn1 = ref_utils.adjoverlap(adj, adj, e, False)
                    cn2 = ref_utils.adjoverlap(adj, adj2, e, False)
                    out = pred.multidomainforward(x, adj, cn1, cn2, e, types.SimpleNamespace(sum=fx["fill"]))
                    assert torch.allclose(out.cpu(), call["out"], rtol=1e-4, atol=1e-5)
    finally:
        shim.uninstall()
        sys.path.remove(REF)
