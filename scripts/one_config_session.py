"""One fused order-2 session of a BASELINE config (collab / ddi / pubmed / cora) for an ncu launch list."""
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
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    s = ob.CNSession(GG, e, gg.batch, 2).build(2, False)
    if variant == 5:
        s.stats(5, 0.0, ip3, 0)
    r = s.aggregate(xx, variant, 1.0 if variant == 7 else 0.0, ip3)
    s.release()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
