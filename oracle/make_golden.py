"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/ref_model_*.pt by EXECUTING the reference.

    python oracle/make_golden.py            # needs /root/reference (read-only); run in the authoring container

The reference's own ``utils.py`` and ``model.py`` are imported unmodified and ``get_cn1_cn2`` is exec'd
from the text of ``NeighborOverlapCitation2.py`` (the driver module itself cannot be imported: it needs
ogb / tensorboard at import time).  The third-party packages they import (torch_sparse, pygho,
torch_geometric) are not installable here; ``oracle/emul/`` provides pure-torch stand-ins for exactly the
subset touched.  What the fixtures pin is therefore the reference's *own* Python logic
(adjoverlap's searchsorted path, the 170-line cn5 / cn7 / cn6 multidomainforward bodies, the running inner
product) -- the library semantics stay "[recalled]" and are cross-checked by oracle/brute.py.

Each fixture holds: graph CSR, links, features, predictor state_dict, the CN matrices the reference built,
the inputs of the xcn1lin / xcn2lin / xcn3lin / xijlin heads captured by forward pre-hooks (= the
aggregates of model.py:2426-2429), the final scores, and the inner-product buffer after every call.
"""
import argparse
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

import model as ref_model  # noqa: E402  (the reference's model.py)
import utils as ref_utils  # noqa: E402  (the reference's utils.py)
import torch_sparse  # noqa: E402  (oracle/emul)
from pygho import SparseTensor as pSparseTensor  # noqa: E402
from pygho.backend.Spspmm import spsphadamard, spspmm  # noqa: E402
from torch_sparse import SparseTensor  # noqa: E402

from ocn_b200 import synth  # noqa: E402


def _ref_get_cn1_cn2():
    text = open(os.path.join(REF, "NeighborOverlapCitation2.py")).read()
    m = re.search(r"^def get_cn1_cn2\(adj,tedge\):\n(?:.*\n)*?    return cn1,cn2\n", text, flags=re.M)
    ns = {"torch": torch, "torch_sparse": torch_sparse, "spsphadamard": spsphadamard, "spspmm": spspmm}
    exec(m.group(0), ns)
    return ns["get_cn1_cn2"]


get_cn1_cn2 = _ref_get_cn1_cn2()


def get_cn3(adj, tedge):
    """Order-3 extension in the reference's own idiom: Ej3 = spspmm(Ej2, 1, adj, 0)."""
    Ei = adj.index_select([0], tedge[0].unsqueeze(0))
    Ej = adj.index_select([0], tedge[1].unsqueeze(0))
    Ej3 = spspmm(spspmm(Ej, 1, adj, 0), 1, adj, 0)
    cn3 = spsphadamard(Ei, Ej3).to_torch_sparse_coo()
    r, c = cn3.indices()
    return torch_sparse.SparseTensor(row=r, col=c, value=cn3.values(), sparse_sizes=tuple(cn3.shape))


def _sp_dump(s):
    r, c, v = s.coo()
    return {"row": r.clone(), "col": c.clone(), "val": None if v is None else v.clone(), "shape": tuple(s.sizes())}


class Capture:
    def __init__(self, pred):
        self.store = {}
        self.handles = []
        for name in ("xcn1lin", "xcn2lin", "xcn3lin", "xijlin"):
            mod = getattr(pred, name, None)
            if isinstance(mod, torch.nn.Module):
                self.handles.append(mod.register_forward_pre_hook(self._hook(name)))

    def _hook(self, name):
        def fn(_m, inp):
            self.store[name] = inp[0].detach().clone()
        return fn

    def pop(self):
        out, self.store = self.store, {}
        return out


def run_case(name, graph, F, batches, predictor, mode, style, fill=None, ln=False, seed=0):
    torch.manual_seed(seed)
    n = graph.n
    rowptr, col = graph.rowptr, graph.col.long()
    row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    adj = SparseTensor(row=row, col=col, sparse_sizes=(n, n), is_sorted=True)
    x = graph.features(F)
    cls = {"cn5": ref_model.CNLinkPredictorOringin, "cn6": ref_model.CNLinkPredictor3hopCNs,
           "cn7": ref_model.CNLinkPredictorbaselearn}[predictor]
    pred = cls(F, F, 1, 3, 0.0, ln=ln)
    pred.train() if mode == "train" else pred.eval()
    cap = Capture(pred)
    args = types.SimpleNamespace(sum=fill)
    if style == "large":
        spadj = adj.to_torch_sparse_coo_tensor()
        adj2 = SparseTensor.from_torch_sparse_coo_tensor(spadj @ spadj, False)   # NeighborOverlap_large.py:74
    else:
        padj = pSparseTensor(torch.stack((row, col)), torch.ones(row.numel()), (n, n), is_coalesced=True)
    calls = []
    with torch.set_grad_enabled(False):
        for e in batches:
            if style == "large":
                cn1 = ref_utils.adjoverlap(adj, adj, e, False)                      # NeighborOverlap_large.py:78
                cn2 = ref_utils.adjoverlap(adj, adj2, e, False)                     # :79
                fadj = adj
            else:
                cn1, cn2 = get_cn1_cn2(padj, e)                                     # NeighborOverlapCitation2.py:169
                fadj = padj
            rec = {"edges": e.clone(), "cn1": _sp_dump(cn1), "cn2": _sp_dump(cn2)}
            if predictor == "cn6":
                cn3 = get_cn3(padj, e)
                rec["cn3"] = _sp_dump(cn3)
                out = pred.multidomainforward(x, fadj, cn1, cn2, cn3, e, args)
            elif predictor == "cn7":
                out = pred.multidomainforward(x, fadj, cn1, cn2, e, args)
            else:
                out = pred.multidomainforward(x, fadj, cn1, cn2, e, args)
            rec.update(cap.pop())
            rec["out"] = out.detach().clone()
            rec["innerprod"] = pred.innerprod.detach().clone()
            rec["n"] = pred.n
            calls.append(rec)
    fx = {"name": name, "n": n, "rowptr": rowptr.clone(), "col": graph.col.clone(), "x": x, "F": F, "predictor": predictor,
          "mode": mode, "style": style, "fill": fill, "ln": ln, "state_dict": {k: v.clone() for k, v in pred.state_dict().items()},
          "calls": calls}
    path = os.path.join(ROOT, "tests", "golden", f"ref_model_{name}.pt")
    torch.save(fx, path)
    print(f"wrote {path}: {len(calls)} call(s), {os.path.getsize(path) / 1024:.0f} KiB")


def run_utils_case(graphs):
    """The reference's own ``utils.py`` on the two remaining callers of the path: ``sparse_tensor_multiply``
    (--adj2byblock, NeighborOverlap_large.py:68-71; block sizes that do and do not divide n) followed by the driver's
    ``adjoverlap(adj, adj2, edge)`` (:79), and ``adjoverlap(..., calresadj=True)`` (utils.py:260-274)."""
    cases = []
    for name, graph, block, e in graphs:
        n = graph.n
        rowptr, col = graph.rowptr, graph.col.long()
        row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
        adj = SparseTensor(row=row, col=col, sparse_sizes=(n, n), is_sorted=True)
        spadj = SparseTensor.from_torch_sparse_coo_tensor(adj.to_torch_sparse_coo_tensor())   # NeighborOverlap_large.py:67-70
        adj2 = ref_utils.sparse_tensor_multiply(spadj, block_size=block)                       # :71
        cn2 = ref_utils.adjoverlap(adj, adj2, e, False)                                        # :79
        ov, r1, r2 = ref_utils.adjoverlap(adj, adj, e, False, calresadj=True)                  # utils.py:260-274
        cases.append({"name": name, "n": n, "rowptr": rowptr.clone(), "col": graph.col.clone(), "block": block,
                      "edges": e.clone(), "adj2": _sp_dump(adj2), "cn2": _sp_dump(cn2),
                      "overlap": _sp_dump(ov), "res1": _sp_dump(r1), "res2": _sp_dump(r2)})
    path = os.path.join(ROOT, "tests", "golden", "ref_utils_adj2byblock_calresadj.pt")
    torch.save(cases, path)
    print(f"wrote {path}: {len(cases)} case(s), {os.path.getsize(path) / 1024:.0f} KiB")


def run_completion_case(name, graph, B, mode, ncalls, trainresdeg, testresdeg, seed, cls_name="IncompleteCN1Predictor"):
    """cn2 = IncompleteCN1Predictor (model.py:843-1146) at depth 1, in = hidden = 64 (its xijlin is a Linear(64, .)
    applied twice).  The draws of sparsesample_reweight's torch.rand are recorded so that a replay is deterministic."""
    torch.manual_seed(seed)
    n = graph.n
    rowptr, col = graph.rowptr, graph.col.long()
    row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    adj = SparseTensor(row=row, col=col, sparse_sizes=(n, n), is_sorted=True)
    x = graph.features(64)
    pred = getattr(ref_model, cls_name)(64, 64, 1, 3, 0.0, trainresdeg=trainresdeg, testresdeg=testresdeg, depth=1)
    pred.train() if mode == "train" else pred.eval()
    draws = []
    real_rand = torch.rand

    def rand(*a, **k):
        out = real_rand(*a, **k)
        draws.append(out.clone())
        return out
    calls = []
    torch.rand = rand
    try:
        with torch.no_grad():
            for s in range(ncalls):
                neg = torch.stack((synth.hash_randint(B - B // 2, n, 270 + s, 1, "cpu"), synth.hash_randint(B - B // 2, n, 270 + s, 2, "cpu")))
                e = torch.cat((graph.query_edges(B // 2, "pos"), neg), 1)
                first = len(draws)
                out = pred(x, adj, None, None, e) if cls_name.endswith("highorder") else pred(x, adj, e)
                calls.append({"edges": e.clone(), "out": out.detach().clone(), "innerprod": pred.innerprod.detach().clone(),
                              "n": pred.n, "draws": draws[first:]})
    finally:
        torch.rand = real_rand
    fx = {"name": name, "n": n, "rowptr": rowptr.clone(), "col": graph.col.clone(), "x": x, "mode": mode, "cls": cls_name,
          "trainresdeg": trainresdeg, "testresdeg": testresdeg,
          "state_dict": {k: v.clone() for k, v in pred.state_dict().items()}, "calls": calls}
    path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.pt")
    torch.save(fx, path)
    print(f"wrote {path}: {len(calls)} call(s), {sum(len(c['draws']) for c in calls)} recorded draws, {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="regex: write only the fixtures whose name matches")
    only = re.compile(ap.parse_args().only)
    write_case = globals()["run_case"]

    def run_case(name, *a, **k):
        if only.search(name):
            write_case(name, *a, **k)
    tiny = synth.tiny_graph(60, 260, 3)
    cora = synth.make_graph("cora", scale=0.12)
    cit = synth.make_graph("citation2", scale=0.0002)

    def links(g, B, k):
        """k batches: half positive edges of the graph, half uniform random pairs (fresh per batch)."""
        out = []
        for s in range(k):
            neg = torch.stack((synth.hash_randint(B - B // 2, g.n, 70 + s, 1, "cpu"),
                               synth.hash_randint(B - B // 2, g.n, 70 + s, 2, "cpu")))
            out.append(torch.cat((g.query_edges(B // 2, "pos"), neg), 1))
        return out

    if only.search("utils_adj2byblock_calresadj"):
        mid = synth.tiny_graph(150, 700, 8)
        run_utils_case([("tiny_b16", tiny, 16, links(tiny, 48, 1)[0]), ("tiny_b60", tiny, 60, links(tiny, 48, 1)[0]),
                        ("tiny_b1024", tiny, 1024, links(tiny, 48, 1)[0]), ("mid_b64", mid, 64, links(mid, 64, 1)[0])])
    if only.search("cn2_eval_cora"):
        run_completion_case("cn2_eval_cora", synth.make_graph("cora", scale=0.06), 32, "eval", 1, 8, 6, 11)
    if only.search("cn2_train_tiny"):
        run_completion_case("cn2_train_tiny", tiny, 24, "train", 3, 3, 128, 12)
    if only.search("cn3_train_tiny"):
        run_completion_case("cn3_train_tiny", synth.tiny_graph(40, 150, 11), 10, "train", 2, 3, 128, 14,
                            cls_name="IncompleteCN1Predictorhighorder")
    if only.search("cn3_eval_tiny"):
        run_completion_case("cn3_eval_tiny", synth.tiny_graph(40, 150, 11), 10, "eval", 1, 8, 5, 15,
                            cls_name="IncompleteCN1Predictorhighorder")
    if only.search("cn4_train_tiny"):
        run_completion_case("cn4_train_tiny", tiny, 24, "train", 2, 3, 128, 13, cls_name="IncompleteCN1PredictorSaveMemory")
    run_case("cn5_large_eval_tiny", tiny, 8, links(tiny, 48, 1), "cn5", "eval", "large")
    run_case("cn5_large_train_cora", cora, 16, links(cora, 96, 3), "cn5", "train", "large")
    run_case("cn5_large_eval_ln_cora", cora, 16, links(cora, 96, 1), "cn5", "eval", "large", ln=True)
    run_case("cn7_large_sum1_cora", cora, 16, links(cora, 96, 2), "cn7", "eval", "large", fill=1)
    run_case("cn7_large_sum0_tiny", tiny, 8, links(tiny, 48, 1), "cn7", "train", "large", fill=0)
    run_case("cn5_pygho_eval_cit", cit, 8, [cit.query_edges(64, "stream")], "cn5", "eval", "pygho")
    run_case("cn5_pygho_train_cit", cit, 8, links(cit, 64, 3), "cn5", "train", "pygho")
    run_case("cn6_pygho_eval_cit", cit, 8, [cit.query_edges(64, "stream")], "cn6", "eval", "pygho")
    run_case("cn6_pygho_train_tiny", tiny, 8, links(tiny, 48, 3), "cn6", "train", "pygho")
    # widths served by the fused inference head (csrc/head.cu): 32 (the citation2 config) and 64
    run_case("cn6_pygho_eval_cit_f32", cit, 32, [cit.query_edges(64, "stream")], "cn6", "eval", "pygho", seed=1)
    run_case("cn5_large_eval_ln_cora_f32", cora, 32, links(cora, 96, 1), "cn5", "eval", "large", ln=True, seed=2)
    run_case("cn7_large_sum1_tiny_f64", tiny, 64, links(tiny, 48, 1), "cn7", "eval", "large", fill=1, seed=3)


if __name__ == "__main__":
    main()
