"""TEST INFRASTRUCTURE ONLY -- record the sparse-API calls of the reference's own text into tests/golden/ref_trace_*.pt.

    python oracle/make_trace.py            # needs /root/reference (read-only); run in the authoring container

The reference's ``utils.py`` and ``model.py`` are imported unmodified on top of ``oracle/emul`` with every stand-in
function wrapped by ``oracle/trace.Recorder``; ``get_cn1_cn2`` is exec'd from the text of
``NeighborOverlapCitation2.py``.  Each fixture is the ordered list of API calls (arguments + results) that
``adjoverlap`` / ``get_cn1_cn2`` / ``multidomainforward`` made.  ``tests/test_gpu_shim.py`` replays them against the
CUDA import shim (``ocn_b200/shim``) call by call.
"""
import os
import re
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "emul"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import pygho  # noqa: E402  (oracle/emul)
import pygho.backend.Spmm as emul_spmm  # noqa: E402
import pygho.backend.Spspmm as emul_spspmm  # noqa: E402
import torch_sparse  # noqa: E402  (oracle/emul)

from oracle.trace import Recorder  # noqa: E402

REC = Recorder()
REC.install(torch_sparse, pygho, emul_spspmm, emul_spmm)      # BEFORE the reference binds the names

import model as ref_model  # noqa: E402  (the reference's model.py)
import utils as ref_utils  # noqa: E402  (the reference's utils.py)

from ocn_b200 import synth  # noqa: E402


def _ref_get_cn1_cn2():
    text = open(os.path.join(REF, "NeighborOverlapCitation2.py")).read()
    m = re.search(r"^def get_cn1_cn2\(adj,tedge\):\n(?:.*\n)*?    return cn1,cn2\n", text, flags=re.M)
    ns = {"torch": torch, "torch_sparse": torch_sparse, "spsphadamard": emul_spspmm.spsphadamard, "spspmm": emul_spspmm.spspmm}
    exec(m.group(0), ns)
    return ns["get_cn1_cn2"]


get_cn1_cn2 = _ref_get_cn1_cn2()


def get_cn3(adj, tedge):
    """Order-3 extension in the reference's idiom, as oracle/make_golden.py."""
    Ei = adj.index_select([0], tedge[0].unsqueeze(0))
    Ej = adj.index_select([0], tedge[1].unsqueeze(0))
    Ej3 = emul_spspmm.spspmm(emul_spspmm.spspmm(Ej, 1, adj, 0), 1, adj, 0)
    cn3 = emul_spspmm.spsphadamard(Ei, Ej3).to_torch_sparse_coo()
    r, c = cn3.indices()
    return torch_sparse.SparseTensor(row=r, col=c, value=cn3.values(), sparse_sizes=tuple(cn3.shape))


def links(g, B, seed):
    neg = torch.stack((synth.hash_randint(B - B // 2, g.n, 170 + seed, 1, "cpu"),
                       synth.hash_randint(B - B // 2, g.n, 170 + seed, 2, "cpu")))
    return torch.cat((g.query_edges(B // 2, "pos"), neg), 1)


def case(name, graph, F, predictor, mode, style, fill=None, B=48, seed=0, calres=False):
    torch.manual_seed(seed)
    n = graph.n
    rowptr, col = graph.rowptr, graph.col.long()
    row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    e = links(graph, B, seed)
    x = graph.features(F)
    cls = {"cn5": ref_model.CNLinkPredictorOringin, "cn6": ref_model.CNLinkPredictor3hopCNs,
           "cn7": ref_model.CNLinkPredictorbaselearn}[predictor]
    pred = cls(F, F, 1, 3, 0.0)
    pred.train() if mode == "train" else pred.eval()
    args = types.SimpleNamespace(sum=fill)
    REC.take()
    with torch.no_grad():
        # ---- everything from here on is logged: the driver's graph construction included
        ei = torch.stack((row, col))
        half = ei[:, ei[0] < ei[1]]
        adj = torch_sparse.SparseTensor.from_edge_index(half, sparse_sizes=(n, n)).to_symmetric()  # NeighborOverlap_large.py:59-63
        if style == "large":
            spadj = adj.to_torch_sparse_coo_tensor()
            adj2 = torch_sparse.SparseTensor.from_torch_sparse_coo_tensor(spadj @ spadj, False)    # :74
            cn1 = ref_utils.adjoverlap(adj, adj, e, False)                                         # :78
            cn2 = ref_utils.adjoverlap(adj, adj2, e, False)                                        # :79
            if calres:
                ref_utils.adjoverlap(adj, adj, e, False, calresadj=True)                           # utils.py:260-274
            fadj = adj
        else:
            r, c, v = adj.coo()
            v = torch.ones_like(r, dtype=torch.float)
            padj = pygho.SparseTensor(torch.stack((r, c)), v, adj.sizes(), is_coalesced=True)      # NeighborOverlapCitation2.py:147-151
            cn1, cn2 = get_cn1_cn2(padj, e)                                                        # :169
            fadj = padj
        if predictor == "cn6":
            out = pred.multidomainforward(x, fadj, cn1, cn2, get_cn3(padj, e), e, args)
        else:
            out = pred.multidomainforward(x, fadj, cn1, cn2, e, args)
    calls = REC.take()
    fx = {"name": name, "n": n, "predictor": predictor, "mode": mode, "style": style, "calls": calls,
          "out": out.detach().clone()}
    path = os.path.join(ROOT, "tests", "golden", f"ref_trace_{name}.pt")
    torch.save(fx, path)
    names = sorted({c["fn"] for c in calls})
    print(f"wrote {path}: {len(calls)} calls, {os.path.getsize(path) / 1024:.0f} KiB; API: {', '.join(names)}")


def completion_case(name, cls_name, mode, seed, B=16):
    """The NCNC-style completion predictors cn2 / cn3 / cn4 (model.py:843-1886) at depth 1: adjoverlap with
    calresadj=True, the residual links scored by the depth-0 pass, sparsesample_reweight where a residual row is
    longer than the sampling degree, then the cn5-style normalisation / orthogonalisation.  Their xijlin is a
    Linear(64, hidden) applied twice (model.py:576, 902, 1126), so they only run at in = hidden = 64."""
    g = synth.tiny_graph(40, 150, 11)
    n = g.n
    torch.manual_seed(seed)
    row = torch.repeat_interleave(torch.arange(n), g.rowptr[1:] - g.rowptr[:-1])
    adj = torch_sparse.SparseTensor(row=row, col=g.col.long(), sparse_sizes=(n, n), is_sorted=True)
    e = links(g, B, seed)
    x = g.features(64)
    pred = getattr(ref_model, cls_name)(64, 64, 1, 3, 0.0, trainresdeg=4, testresdeg=6, depth=1)
    pred.train() if mode == "train" else pred.eval()
    REC.take()
    with torch.no_grad():
        if cls_name == "IncompleteCN1Predictorhighorder":
            out = pred(x, adj, None, None, e)
        else:
            out = pred(x, adj, e)
    calls = REC.take()
    path = os.path.join(ROOT, "tests", "golden", f"ref_trace_{name}.pt")
    torch.save({"name": name, "n": n, "predictor": cls_name, "mode": mode, "calls": calls, "out": out.detach().clone()}, path)
    print(f"wrote {path}: {len(calls)} calls, {os.path.getsize(path) / 1024:.0f} KiB; API: {', '.join(sorted({c['fn'] for c in calls}))}")


def conv_case():
    """PureConv (model.py:42-55) in its four modes + DropAdj (model.py:219-229): the GNN side of the API."""
    g = synth.tiny_graph(60, 260, 3)
    n = g.n
    row = torch.repeat_interleave(torch.arange(n), g.rowptr[1:] - g.rowptr[:-1])
    adj = torch_sparse.SparseTensor(row=row, col=g.col.long(), sparse_sizes=(n, n), is_sorted=True)
    x = g.features(8)
    torch.manual_seed(5)
    REC.take()
    with torch.no_grad():
        for aggr in ("gcn", "sum", "mean", "max"):
            ref_model.PureConv(8, 8, aggr)(x, adj)
        drop = ref_model.DropAdj(0.3)
        drop.train()
        dropped = drop(adj)
        ref_model.PureConv(8, 8, "gcn")(x, dropped)
    calls = REC.take()
    path = os.path.join(ROOT, "tests", "golden", "ref_trace_conv.pt")
    torch.save({"name": "conv", "n": n, "calls": calls}, path)
    print(f"wrote {path}: {len(calls)} calls; API: {', '.join(sorted({c['fn'] for c in calls}))}")


def main():
    tiny = synth.tiny_graph(60, 260, 3)
    cora = synth.make_graph("cora", scale=0.06)
    case("cn5_large_eval", tiny, 8, "cn5", "eval", "large", calres=True)
    case("cn5_large_train", cora, 8, "cn5", "train", "large", B=64, seed=1)
    case("cn7_large_sum1", tiny, 8, "cn7", "eval", "large", fill=1, seed=2)
    case("cn5_pygho_train", tiny, 8, "cn5", "train", "pygho", seed=3)
    case("cn6_pygho_eval", tiny, 8, "cn6", "eval", "pygho", seed=4)
    conv_case()
    completion_case("cn2_eval", "IncompleteCN1Predictor", "eval", 6)
    completion_case("cn2_train", "IncompleteCN1Predictor", "train", 7)
    completion_case("cn3_eval", "IncompleteCN1Predictorhighorder", "eval", 8, B=4)
    completion_case("cn4_eval", "IncompleteCN1PredictorSaveMemory", "eval", 9, B=12)


if __name__ == "__main__":
    main()
