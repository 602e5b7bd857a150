"""Fused inference-mode predictor head: ``lin(a0*xcn1lin(xcn1) + a1*xcn2lin(xcn2) [+ a2*xcn3lin(xcn3)]
+ beta*xijlin(xij))`` (model.py:2192-2235, 2429-2437) as one kernel of libocn_b200 for hidden widths 32 / 64.
Training, autograd and wider heads keep the torch modules (cuBLAS GEMMs)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .cn import _stream


def _linears(seq: nn.Sequential) -> List[nn.Module]:
    return [m for m in seq if isinstance(m, (nn.Linear, nn.LayerNorm))]


def _pack_seq(seq: nn.Sequential, last_plain: bool = False) -> List[Tensor]:
    """Flat pieces of one Sequential in module order: Linear -> W^T, b; LayerNorm -> weight, bias.  The very last
    Linear of ``lin`` keeps its [out, h] layout (``last_plain``)."""
    mods = _linears(seq)
    out = []
    for k, m in enumerate(mods):
        if isinstance(m, nn.Linear):
            w = m.weight if (last_plain and k == len(mods) - 1) else m.weight.t()
            out += [w.reshape(-1), m.bias.reshape(-1)]
        else:
            out += [m.weight.reshape(-1), m.bias.reshape(-1)]
    return out


def supported(pred, in_ch: int) -> int:
    """Number of floats of the packed parameter buffer, or -1 when this head is not served by the fused kernel."""
    lin0 = pred.lin[0]
    hid, out_ch = lin0.in_features, pred.lin[-1].out_features
    if any(isinstance(m, nn.Dropout) and m.p > 0 and pred.training for m in pred.modules()):
        return -1
    return _lib.lib().ocn_cn_head_params(int(in_ch), int(hid), int(out_ch), _flags(pred), 3 if hasattr(pred, "xcn3lin") and pred.order >= 3 else 2)


def _flags(pred) -> int:
    ln = any(isinstance(m, nn.LayerNorm) for m in pred.lin)
    tailact = not isinstance(pred.xijlin[-1], nn.Linear)
    two = sum(isinstance(m, nn.Linear) for m in pred.lin) == 3
    return (1 if ln else 0) | (2 if tailact else 0) | (4 if two else 0)


def packed_params(pred) -> Tensor:
    """The head's parameters in the order csrc/head.cu documents; cached until a parameter changes."""
    ps = [p for n, p in pred.named_parameters() if n.split(".")[0] in ("xcn1lin", "xcn2lin", "xcn3lin", "xijlin", "lin")]
    key = tuple((p.data_ptr(), p._version) for p in ps)
    cache = getattr(pred, "_ocn_head_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    pieces = _pack_seq(pred.xcn1lin) + _pack_seq(pred.xcn2lin)
    if pred.order >= 3 and hasattr(pred, "xcn3lin"):
        pieces += _pack_seq(pred.xcn3lin)
    pieces += _pack_seq(pred.xijlin) + _pack_seq(pred.lin, last_plain=True)
    with torch.no_grad():
        flat = torch.cat([t.detach().float().reshape(-1) for t in pieces]).contiguous()
    pred._ocn_head_cache = (key, flat)
    return flat


def fused_head(pred, xcn1: Tensor, xcn2: Tensor, xcn3: Optional[Tensor], xij: Tensor) -> Tensor:
    L = _lib.lib()
    B, in_ch = xcn1.shape
    hid, out_ch = pred.lin[0].in_features, pred.lin[-1].out_features
    params = packed_params(pred)
    with torch.no_grad():
        alpha = torch.sigmoid(pred.alpha).cumprod(-1)
        mix = torch.cat((alpha[:3].float(), pred.beta.detach().float().reshape(1))).contiguous()
    out = torch.empty(B, out_ch, dtype=torch.float32, device=xcn1.device)
    with torch.cuda.device(xcn1.device):
        _lib.check(L.ocn_cn_head(_lib.ptr(xcn1.contiguous()), _lib.ptr(xcn2.contiguous()),
                                 _lib.ptr(None if xcn3 is None else xcn3.contiguous()), _lib.ptr(xij.contiguous()), B,
                                 in_ch, hid, out_ch, _flags(pred), _lib.ptr(params), params.numel(), _lib.ptr(mix),
                                 _lib.ptr(out), _stream(xcn1.device)), "ocn_cn_head")
    return out
