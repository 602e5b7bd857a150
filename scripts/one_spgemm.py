import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "collab"
g = synth.make_graph(name, device="cuda:0"); G = ob.Graph(g.rowptr, g.col, g.n)
ob.spgemm_a2(G, int(sys.argv[2]) if len(sys.argv) > 2 else 0, True)
