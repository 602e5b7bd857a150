"""One GNN SpMM layer at citation2 shape (for an ncu capture of k_spmm)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = synth.make_graph("citation2", device="cuda:0")
G = ob.Graph(g.rowptr, g.col, g.n)
x = g.features(F, device="cuda:0")
for _ in range(3):
    y = ob.pure_conv(x, G, "sum")
torch.cuda.synchronize()
