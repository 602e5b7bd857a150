"""One-GPU A/B of the predictor head at hidden 64 / 128 / 256: ocn_linear_tc (tcgen05, one launch per layer) against the
fused CUDA-core kernel (where it serves the width) and the torch modules (cuBLAS fp32, and with TF32 allowed); errors against
the modules evaluated in float64."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import head

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for cls, F, ln in (("cn5", 64, True), ("cn6", 64, False), ("cn5", 128, True), ("cn5", 256, True), ("cn7", 256, False)):
    torch.manual_seed(0)
    pred = ob.predictor_dict[cls](F, F, 1, 3, 0.0, ln=ln).to(dev).eval()
    xs = [torch.randn(B, F, device=dev) * s for s in (1.0, 3.0, 0.5, 2.0)]
    x3 = xs[2] if cls == "cn6" else None
    p64 = copy.deepcopy(pred).double()
    p64.fuse_head = False
    with torch.no_grad():
        ref = p64._head(xs[0].double(), xs[1].double(), None if x3 is None else x3.double(), xs[3].double())

    def timed(fn, reps=10):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record(); torch.cuda.synchronize()
        return out, a.elapsed_time(b) / reps

    res = {}
    with torch.no_grad():
        assert head.wide_supported(pred, F)
        out, ms = timed(lambda: head.fused_head_wide(pred, xs[0], xs[1], x3, xs[3]))
        res["tcgen05 per layer"] = (ms, out)
        if head.supported(pred, F) > 0:
            out, ms = timed(lambda: head.fused_head(pred, xs[0], xs[1], x3, xs[3]))
            res["cuda cores fused"] = (ms, out)
        pred.fuse_head = False
        out, ms = timed(lambda: pred._head(xs[0], xs[1], x3, xs[3]))
        res["torch fp32"] = (ms, out)
        torch.backends.cuda.matmul.allow_tf32 = True
        out, ms = timed(lambda: pred._head(xs[0], xs[1], x3, xs[3]))
        res["torch tf32"] = (ms, out)
        torch.backends.cuda.matmul.allow_tf32 = False
        pred.fuse_head = True
    scale = 1 + ref.abs().max().item()
    print(f"{cls} F={F} ln={ln} B={B}: " + "   ".join(f"{k} {v[0] * 1e3:8.1f} us (rel.err {(v[1].double() - ref).abs().max().item() / scale:.1e})"
                                                         for k, v in res.items()), flush=True)
