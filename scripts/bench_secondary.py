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


def before_after(out):
    """The steps either side of the path at citation2 shape: device CSR build, per-batch
    --maskinput adjacency (vs rebuilding with torch sort/unique on the same GPU, the reference's op sequence),
    the fused inference head (vs the torch modules) and the ranking metrics."""
    g = synth.make_graph("citation2", device=DEV)
    n = g.n
    el = torch.stack((g.raw_src, g.raw_dst)).to(DEV)
    E = el.shape[1]
    ms = timeit(lambda: ob.Graph.from_edge_index(el, n, with_multiplicity=True), reps=3, warm=1)
    out.append({"op": "graph_build (from_edge_index + to_symmetric)", "edges": E, "ms": ms, "Medges_per_s": E / ms / 1e3})
    print(out[-1], flush=True)
    G = ob.Graph.from_edge_index(el, n, with_multiplicity=True)
    perm = torch.randperm(E, device=DEV)[:16384]
    links = el[:, perm].contiguous()
    ms = timeit(lambda: G.masked(links), reps=5, warm=2)
    byt = 2 * 8 * (n + 1) + 3 * 4 * G.nnz + 4 * G.nnz  # rowptr in/out, col + mult + dec read twice-ish, col written
    out.append({"op": "graph_mask (adjmask[perm]=0 rebuild) by multiplicity decrement", "masked_links": 16384, "nnz": G.nnz,
                "ms": ms, "alg_GBs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / PEAK})
    print(out[-1], flush=True)

    def torch_rebuild():  # the reference's per-batch op sequence on the same GPU (sort-based)
        keep = torch.ones(E, dtype=torch.bool, device=DEV)
        keep[perm] = False
        s, d = el[0][keep], el[1][keep]
        key = torch.unique(torch.cat((s, d)) * n + torch.cat((d, s)))
        return key
    ms2 = timeit(torch_rebuild, reps=3, warm=1)
    out.append({"op": "graph_mask reference op sequence (torch.unique of the remaining list, same GPU)", "ms": ms2,
                "speedup_of_graph_mask": ms2 / ms})
    print(out[-1], flush=True)
    del G, el
    torch.cuda.empty_cache()
    for F in (32, 64):
        torch.manual_seed(0)
        pred = ob.CNLinkPredictor3hopCNs(F, F, 1, 3, 0.0).to(DEV).eval()
        xs = [torch.randn(65536, F, device=DEV) for _ in range(4)]
        with torch.no_grad():
            pred.fuse_head = True
            ms = timeit(lambda: pred._head(*xs))
            pred.fuse_head = False
            ms2 = timeit(lambda: pred._head(*xs))
        out.append({"op": "predictor head, 65536 links", "F": F, "fused_ms": ms, "torch_modules_ms": ms2,
                    "Mlinks_per_s": 65536 / ms / 1e3, "speedup": ms2 / ms})
        print(out[-1], flush=True)
    pos, neg = torch.randn(86596, device=DEV), torch.randn(86596, 1000, device=DEV)
    ms = timeit(lambda: ob.metrics.mrr_list(pos, neg))
    out.append({"op": "mrr (86596 sources x 1000 negatives)", "ms": ms, "alg_GBs": 4 * neg.numel() / ms / 1e6,
                "frac_of_measured_hbm": 4 * neg.numel() / ms / 1e6 / PEAK})
    print(out[-1], flush=True)
    negf = torch.randn(3_000_000, device=DEV)
    ms = timeit(lambda: ob.metrics.hits_at_k(pos, negf, 100))
    out.append({"op": "hits@100 (3M negatives)", "ms": ms})
    print(out[-1], flush=True)


def training(out):
    """One optimiser step of the citation2 driver's predictor loop (NeighborOverlapCitation2.py:131-209): 16 384
    positive links + 16 384 negatives in sub-batches of 2048 (one source per link: the per-run table path),
    forward (CN sets, statistics, aggregation, heads) + backward, through ocn_b200.dist.sharded_train_step."""
    from ocn_b200.dist import sharded_train_step
    g = synth.make_graph("citation2", device=DEV)
    G = ob.Graph(g.rowptr, g.col, g.n)
    for order, cls in ((2, ob.CNLinkPredictorOringin),):
        torch.manual_seed(0)
        pred = cls(32, 32, 1, 3, 0.0, weighted=True).to(DEV).train()
        h = g.features(32, device=DEV).requires_grad_(True)
        pos = g.query_edges(16384, "pos", device=DEV)
        neg = torch.stack((pos[0], synth.hash_randint(16384, g.n, 5, 9, DEV)))
        subs = [pos[:, k:k + 2048] for k in range(0, 16384, 2048)] + [neg[:, k:k + 2048] for k in range(0, 16384, 2048)]
        signs = [1.0] * 8 + [-1.0] * 8

        def step():
            pred.zero_grad(set_to_none=True)
            h.grad = None
            return sharded_train_step(pred, h, G, subs, signs, 16384, 0, 1)
        ms = timeit(step, reps=3, warm=1)
        out.append({"op": f"training step, cn5 order {order}, 32768 links in 16 sub-batches (fwd + bwd)", "ms": ms,
                    "Mlinks_per_s": 32768 / ms / 1e3})
        print(out[-1], flush=True)


def main():
    out = []
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in ("", "steps"):
        before_after(out)
    if only in ("", "train"):
        training(out)
    graphs = (("citation2", (32, 128)), ("collab", (256,)), ("pubmed", (256,)), ("ddi", (64,)))
    if only in ("steps", "train"):
        graphs = ()
    elif only:
        graphs = tuple(gf for gf in graphs if gf[0] == only)
    for name, feats in graphs:
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
    with open(f"gpurun_out/secondary{('_' + only) if only else ''}.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
