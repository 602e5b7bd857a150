"""One order-2 build of training-shaped links (citation2 shape: 16 384 positives drawn from the edges + 16 384 negatives
with the same sources) for ncu captures of k_cn_build_direct."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth

DEV = "cuda:0"
g = synth.make_graph("citation2", device=DEV)
G = ob.Graph(g.rowptr, g.col, g.n)
pos = g.query_edges(16384, "pos", device=DEV)
neg = torch.stack((pos[0], synth.hash_randint(16384, g.n, 5, 9, DEV)))
e = torch.cat((pos, neg), 1).contiguous()
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    ob.CNSession(G, e, 2048, 2).build(2, True, with_stats=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
