"""One-GPU A/B of the fused predictor head at in = hidden = 32: tcgen05 kernel (head_tc.cu) against the CUDA-core kernel
(head.cu) and the torch modules; errors are measured against the modules evaluated in float64."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import _lib

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for cls, ln in (("cn6", False), ("cn6", True), ("cn5", False)):
    torch.manual_seed(0)
    pred = ob.predictor_dict[cls](32, 32, 1, 3, 0.0, ln=ln).to(dev).eval()
    xs = [torch.randn(B, 32, device=dev) * s for s in (1.0, 3.0, 0.5, 2.0)]
    x3 = xs[2] if cls == "cn6" else None
    p64 = copy.deepcopy(pred).double()
    p64.fuse_head = False
    with torch.no_grad():
        ref = p64._head(xs[0].double(), xs[1].double(), None if x3 is None else x3.double(), xs[3].double())

    def timed(fn, reps=20):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record(); torch.cuda.synchronize()
        return out, a.elapsed_time(b) / reps

    with torch.no_grad():
        res = {}
        for name, opt, fuse in (("tcgen05", 1, True), ("tcgen05 A-in-TMEM x4", 3, True), ("cuda cores", 2, True), ("torch modules", 2, False)):
            _lib.set_option("head_tc", opt)
            pred.fuse_head = fuse
            out, ms = timed(lambda: pred._head(xs[0], xs[1], x3, xs[3]))
            err = (out.double() - ref).abs().max().item() / (1 + ref.abs().max().item())
            res[name] = (ms, err)
        _lib.set_option("head_tc", 0)
    print(f"{cls} ln={ln} B={B}: " + "   ".join(f"{k} {v[0] * 1e3:7.1f} us (rel.err {v[1]:.1e})" for k, v in res.items()), flush=True)
