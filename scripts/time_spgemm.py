"""A^2 SpGEMM timings (ms, CUDA events, median of 3) on the graph shapes the reference materialises A^2 for."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth, _lib
from bench import timeit
for name in ("collab", "ddi", "pubmed", "cora"):
    g = synth.make_graph(name, device="cuda:0"); G = ob.Graph(g.rowptr, g.col, g.n)
    for fold in ((0, 1024) if name == "ddi" else (0,)):
        ms = timeit(lambda: ob.spgemm_a2(G, fold, True), reps=3, warm=1)
        print(name, "fold", fold, "ms", round(ms, 3), flush=True)
