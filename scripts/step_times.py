"""Per-step device time of the fused path over consecutive 65536-link slices of the evaluation stream
(finds slices that hit a slow case, e.g. a hub source)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth

dev = "cuda:0"
n_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 26
hub = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = synth.make_graph("citation2", device=dev)
G = ob.Graph(g.rowptr, g.col, g.n)
x = g.features(32, device=dev)
T = 65536
e_all = g.query_edges(n_slices * T, "stream", device=dev)
ip3 = torch.zeros(3, device=dev)
deg = G.degree()
for rep in range(2):
    for s in range(n_slices):
        e = e_all[:, s * T:(s + 1) * T]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        sess = ob.CNSession(G, e, 2048, 3, hub)
        ev[1].record()
        sess.build(3, True)
        sess.stats(5, 0.0, ip3, 0)
        sess.aggregate(x, 5, 0.0, ip3)
        sess.release()
        ev[2].record()
        torch.cuda.synchronize()
        if rep == 1:
            print(f"slice {s:2d}: plan {ev[0].elapsed_time(ev[1]):6.3f} ms  rest {ev[1].elapsed_time(ev[2]):7.3f} ms  hub_d {sess.hub_degree} "
                  f"runs {sess.num_runs} positions {sess.plan_host[11]} pairs {sess.plan_host[9]} entries {sess.plan_host[10]} "
                  f"max d(src) {int(deg[e[0]].max())} max d(dst) {int(deg[e[1]].max())}", flush=True)
