"""One-GPU A/B of the GNN SpMM kernels at citation2 shape: register gather (k_spmm) against the cp.async.bulk gather
(k_spmm_tma), same process, same inputs; results compared first.   python scripts/ab_spmm.py [graph] [scale]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import _lib, synth

dev = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "citation2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
g = synth.make_graph(name, scale=scale, device=dev)
G = ob.Graph(g.rowptr, g.col, g.n)
Gv = ob.Graph(g.rowptr, g.col, g.n, value=torch.rand(G.nnz, device=dev) + 0.5)
print(f"{name} x{scale}: n={g.n} nnz={G.nnz}", flush=True)


def run(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return out, min(ts), sorted(ts)[len(ts) // 2]


for F in (32, 64, 128, 256):
    x = g.features(F, device=dev)
    for label, fn in (("sum", lambda: ob.pure_conv(x, G, "sum")), ("mean", lambda: ob.pure_conv(x, G, "mean")),
                      ("max", lambda: ob.pure_conv(x, G, "max")), ("gcn", lambda: ob.pure_conv(x, G, "gcn")), ("gcn3", lambda: ob.pure_conv3_gcn(x, G)),
                      ("sum valued", lambda: ob.pure_conv(x, Gv, "sum")), ("gcn valued", lambda: ob.pure_conv(x, Gv, "gcn"))):
        if F > 32 and label not in ("sum", "gcn", "gcn3"):
            continue
        _lib.set_option("spmm_tma", 2)
        ref, t0, m0 = run(fn)
        _lib.set_option("spmm_tma", 1)
        got, t1, m1 = run(fn)
        _lib.set_option("spmm_tma", 3)
        got3, t3, m3 = run(fn)
        _lib.set_option("spmm_tma", 4)
        got4, t4, m4 = run(fn)
        _lib.set_option("spmm_tma", 0)
        err = max((got - ref).abs().max().item(), (got3 - ref).abs().max().item(), (got4 - ref).abs().max().item()) / (1 + ref.abs().max().item())
        alg = (8 * (g.n + 1) + 4 * G.nnz + 4 * F * G.nnz + 4 * F * g.n) / 1e9
        print(f"F={F:3d} {label:10s} register {t0:7.3f} ms (median {m0:7.3f}, {alg / t0 * 1e3:6.0f} GB/s)   "
              f"bulk {t1:7.3f} ms (median {m1:7.3f}, {alg / t1 * 1e3:6.0f} GB/s)   lane {t3:7.3f} ms ({alg / t3 * 1e3:6.0f} GB/s)   cp.async {t4:7.3f} ms (median {m4:7.3f}, {alg / t4 * 1e3:6.0f} GB/s)   rel.err {err:.2e}", flush=True)
    del x
