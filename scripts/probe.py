"""Quick device-side timing probe of the fused CN path (not the bench; used while tuning)."""
import argparse
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="citation2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--kind", default="stream")
    ap.add_argument("--feat", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--hub", type=int, default=0, help="hub_degree: 0 auto, -1 off")
    a = ap.parse_args()
    dev = "cuda:0"
    t0 = time.time()
    g = synth.make_graph(a.graph, device=dev, scale=a.scale)
    G = ob.Graph(g.rowptr, g.col, g.n)
    torch.cuda.synchronize()
    print(f"graph {a.graph} n={g.n} nnz={g.nnz} built in {time.time()-t0:.1f}s", flush=True)
    T = a.batches * a.batch
    e = g.query_edges(T, a.kind, device=dev)
    x = g.features(a.feat, device=dev)
    ip3 = torch.zeros(3, device=dev)
    deg = G.degree()
    F = torch.zeros(g.n, dtype=torch.float64, device=dev).index_add_(0, G.row(), deg[g.col.long()].double())
    print(f"links {T}: mean d(i) {deg[e[0]].double().mean():.1f} mean d(j) {deg[e[1]].double().mean():.1f} "
          f"mean F_i {F[e[0]].mean():.0f} mean F_j {F[e[1]].mean():.0f} max F_j {F[e[1]].max():.0f}", flush=True)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for rep in range(a.reps):
        marks = [ev() for _ in range(6)]
        marks[0].record()
        sess = ob.CNSession(G, e, a.batch, a.order, a.hub)
        marks[1].record()
        sess.build(a.order, True)
        marks[2].record()
        sess.stats(5, 0.0, ip3, 0)
        marks[3].record()
        out = sess.aggregate(x, 5, 0.0, ip3)
        marks[4].record()
        sess.release()
        marks[5].record()
        torch.cuda.synchronize()
        names = ["plan+alloc", "build", "stats", "aggregate", "release"]
        ts = [marks[i].elapsed_time(marks[i + 1]) for i in range(5)]
        tot = sum(ts)
        print(f"rep {rep}: " + "  ".join(f"{n} {t:.3f} ms" for n, t in zip(names, ts)) +
              f"  | total {tot:.3f} ms  {T / tot / 1e3:.3f} M links/s  units {sess.num_units} runs {sess.num_runs} "
              f"records {sess.num_records} hub_d {sess.hub_degree} pairs {sess.plan_host[9]} entries {sess.plan_host[10]}", flush=True)


if __name__ == "__main__":
    main()
