"""Time single ocn_linear_tc launches (65 536 rows; Linear + ReLU, Linear + LayerNorm + ReLU) at N = K = 256 / 128 / 64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import head

dev = "cuda:0"
for F in (256, 128, 64):
    torch.manual_seed(0)
    pred = ob.CNLinkPredictorOringin(F, F, 1, 3, 0.0, ln=True).to(dev).eval()
    x = torch.randn(65536, F, device=dev)
    for label, ln in (("Linear + ReLU", None), ("Linear + LN + ReLU", pred.xcn1lin[4])):
        with torch.no_grad():
            fn = lambda: head.linear_tc(pred, x, pred.xcn1lin[3], ln, True)
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                fn()
            b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        print(f"F={F} {label:20s}: {us:7.1f} us  ({65536 * F * F * 2 / us / 1e6:6.1f} TFLOP/s of fp32-equivalent work)", flush=True)
