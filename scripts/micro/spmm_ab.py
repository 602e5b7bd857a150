"""A/B of SpMM kernel variants in ONE process on ONE GPU (boxes differ by several percent): every variant is a
full libocn_b200 build loaded side by side; only ocn_spmm_csr is called."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ocn_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
P = ctypes.c_void_p
g = synth.make_graph("citation2", device="cuda:0")
libs = {}
for name in sys.argv[1:]:
    L = ctypes.CDLL(os.path.join(HERE, f"libocn_{name}.so"))
    L.ocn_spmm_csr.restype = ctypes.c_int
    L.ocn_spmm_csr.argtypes = [P, P, P, ctypes.c_int64, P, ctypes.c_int64, ctypes.c_int, P, P]
    libs[name] = L
st = torch.cuda.current_stream().cuda_stream
for F in (32, 128, 256):
    x = g.features(F, device="cuda:0")
    out = torch.empty(g.n, F, device="cuda:0")
    ref = None
    for rnd in range(3):
        for name, L in libs.items():
            def run():
                rc = L.ocn_spmm_csr(g.rowptr.data_ptr(), g.col.data_ptr(), None, g.n, x.data_ptr(), F, 0, out.data_ptr(), st)
                assert rc == 0
            run(); run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(ref, out))
            print(f"F={F} round {rnd} {name:6s} median {ts[2]:.3f} ms  min {ts[0]:.3f}  identical_to_first={same}", flush=True)
