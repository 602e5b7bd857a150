"""Order-3 build + cn5 aggregation on UNGROUPED links (one source per link: the training shape) at citation2 shape:
links/s for several session sizes, automatic path choice against the per-run tables (hub_degree = -1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth

dev = "cuda:0"
g = synth.make_graph("citation2", device=dev)
G = ob.Graph(g.rowptr, g.col, g.n)
x = g.features(32, device=dev)
ip3 = torch.zeros(3, device=dev)
for kind in ("uniform", "pos"):
    for T, bs in ((2048, 2048), (8192, 2048), (65536, 2048)):
        if kind == "uniform":
            e = torch.stack((synth.hash_randint(T, g.n, 41, 1, dev), synth.hash_randint(T, g.n, 41, 2, dev)))
        else:
            e = g.query_edges(T, "pos", device=dev)
        for hub in (0, -1):
            def run():
                s = ob.CNSession(G, e, bs, 3, hub)
                s.build(3, True); s.stats(5, 0.0, ip3, 0); out = s.aggregate(x, 5, 0.0, ip3); s.release()
                return s
            s = run(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            print(f"{kind:8s} T={T:6d} hub={hub:2d}: {ms:8.3f} ms  {T / ms / 1e3:7.2f} M links/s  runs={s.num_runs} hub_degree={s.hub_degree}", flush=True)
