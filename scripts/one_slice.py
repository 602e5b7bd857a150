import os, sys, torch
sys.path.insert(0, '/root/repo')
import ocn_b200 as ob
from ocn_b200 import synth
dev="cuda:0"
g = synth.make_graph("citation2", device=dev); G = ob.Graph(g.rowptr, g.col, g.n); x = g.features(32, device=dev)
T=65536; s=int(sys.argv[1])
e = g.query_edges(26*T, "stream", device=dev)[:, s*T:(s+1)*T].contiguous()
ip3=torch.zeros(3,device=dev)
for rep in range(2):
    sess = ob.CNSession(G, e, 2048, 3, 0); sess.build(3, True); sess.stats(5,0.0,ip3,0); sess.aggregate(x,5,0.0,ip3); sess.release(); torch.cuda.synchronize()
