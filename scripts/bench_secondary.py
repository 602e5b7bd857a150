"""Secondary measurements of SURVEY.md §8(d): GNN SpMM GB/s, A^2 SpGEMM products/s, generic intersect,
CN aggregate kernel -- device-timed with CUDA events, algorithmic bytes (no cache credit)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth

DEV = "cuda:0"
PEAK = 6537.0


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = []
    for name, feats in (("citation2", (32, 128)), ("collab", (256,)), ("pubmed", (256,)), ("ddi", (64,))):
        g = synth.make_graph(name, device=DEV)
        G = ob.Graph(g.rowptr, g.col, g.n)
        norm = ob.gcn_norm(G)
        for F in feats:
            x = g.features(F, device=DEV)
            byt = 8 * (g.n + 1) + 4 * g.nnz + 4 * F * g.nnz + 4 * F * g.n
            for mode, fn in (("sum", lambda: ob.pure_conv(x, G, "sum")), ("gcn", lambda: ob.pure_conv(x, G, "gcn", norm))):
                ms = timeit(fn)
                out.append({"op": f"gnn_spmm_{mode}", "graph": name, "F": F, "ms": ms, "alg_GBs": byt / ms / 1e6,
                            "frac_of_measured_hbm": byt / ms / 1e6 / PEAK})
                print(out[-1], flush=True)
        if name in ("collab", "pubmed", "ddi"):
            deg = G.degree()
            prods = int((deg[g.col.long()]).sum())
            ms = timeit(lambda: ob.spgemm_a2(G, 0, True), reps=3, warm=1)
            a2 = ob.spgemm_a2(G, 0, True)
            out.append({"op": "spgemm_a2", "graph": name, "ms": ms, "products": prods, "nnz_out": a2.nnz,
                        "Gproducts_per_s": prods / ms / 1e6})
            print(out[-1], flush=True)
            e = g.query_edges(g.batch, "mixed", device=DEV)
            ms = timeit(lambda: ob.adjoverlap(G, a2, e))
            out.append({"op": "adjoverlap(adj, adj2)", "graph": name, "B": g.batch, "ms": ms, "Mlinks_per_s": g.batch / ms / 1e3})
            print(out[-1], flush=True)
            ip3 = torch.zeros(3, device=DEV)
            x = g.features(g.hidden, device=DEV)

            def fused():
                s = ob.CNSession(G, e, g.batch, 2).build(2, False)
                s.stats(5, 0.0, ip3, 0)
                r = s.aggregate(x, 5 if g.predictor == "cn5" else 7, 0.0, ip3)
                s.release()
                return r
            ms = timeit(fused, reps=3, warm=1)
            out.append({"op": f"fused_{g.predictor}_order2", "graph": name, "B": g.batch, "F": g.hidden, "ms": ms,
                        "Mlinks_per_s": g.batch / ms / 1e3})
            print(out[-1], flush=True)
        del G, g
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/secondary.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
