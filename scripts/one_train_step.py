"""One optimiser step of the citation2 predictor loop (for an ncu launch list: --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth
from ocn_b200.dist import sharded_train_step

DEV = "cuda:0"
g = synth.make_graph("citation2", device=DEV)
G = ob.Graph(g.rowptr, g.col, g.n)
torch.manual_seed(0)
pred = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0, weighted=True).to(DEV).train()
h = g.features(32, device=DEV).requires_grad_(True)
pos = g.query_edges(16384, "pos", device=DEV)
neg = torch.stack((pos[0], synth.hash_randint(16384, g.n, 5, 9, DEV)))
subs = [pos[:, k:k + 2048] for k in range(0, 16384, 2048)] + [neg[:, k:k + 2048] for k in range(0, 16384, 2048)]
signs = [1.0] * 8 + [-1.0] * 8
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    pred.zero_grad(set_to_none=True)
    h.grad = None
    sharded_train_step(pred, h, G, subs, signs, 16384, 0, 1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
