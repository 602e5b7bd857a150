"""Device time of one fused order-2 session of a BASELINE config (CUDA events, median of 20)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "collab"
dev = "cuda:0"
gg = synth.make_graph(name, device=dev)
GG = ob.Graph(gg.rowptr, gg.col, gg.n)
e = gg.query_edges(gg.batch, "mixed", device=dev)
xx = gg.features(gg.hidden, device=dev)
ip3 = torch.zeros(3, device=dev)
variant = 5 if gg.predictor == "cn5" else 7
s = ob.CNSession(GG, e, gg.batch, 2).build(2, False)
if variant == 5:
    s.stats(5, 0.0, ip3, 0)
ts = []
for rep in range(25):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = s.aggregate(xx, variant, 1.0 if variant == 7 else 0.0, ip3)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts = sorted(ts[5:])
print(f"{name:8s} F={gg.hidden} links={gg.batch} OCN_WIDE_DIV={os.environ.get('OCN_WIDE_DIV', '-')}: aggregate {ts[len(ts) // 2] * 1e3:8.1f} us")
s.release()
