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


_HEAD_MODULES = ("xcn1lin", "xcn2lin", "xcn3lin", "xijlin", "lin")


def _static(pred, in_ch: int):
    """Per-module facts that do not change after construction (flags, widths, parameter list, served or not)."""
    st = pred.__dict__.get("_ocn_head_static")
    if st is None or st["in_ch"] != in_ch:
        three = pred.order >= 3 and hasattr(pred, "xcn3lin")
        hid, out_ch = pred.lin[0].in_features, pred.lin[-1].out_features
        flags = _flags(pred)
        names = [m for m in _HEAD_MODULES if m != "xcn3lin" or three]
        ps = [p for m in names for p in getattr(pred, m).parameters()]
        n = _lib.lib().ocn_cn_head_params(int(in_ch), int(hid), int(out_ch), flags, 3 if three else 2)
        st = {"in_ch": in_ch, "hid": hid, "out_ch": out_ch, "flags": flags, "three": three, "params": ps, "n": int(n)}
        pred.__dict__["_ocn_head_static"] = st
    return st


def supported(pred, in_ch: int) -> int:
    """Number of floats of the packed parameter buffer, or -1 when this head is not served by the fused kernel
    (widths other than 32 / 64, more than 200 KB of parameters)."""
    return _static(pred, in_ch)["n"]


def _flags(pred) -> int:
    ln = any(isinstance(m, nn.LayerNorm) for m in pred.lin)
    tailact = not isinstance(pred.xijlin[-1], nn.Linear)
    two = sum(isinstance(m, nn.Linear) for m in pred.lin) == 3
    return (1 if ln else 0) | (2 if tailact else 0) | (4 if two else 0)


def packed_params(pred, in_ch: int) -> Tensor:
    """The head's parameters in the order csrc/head.cu documents; repacked only when a parameter changed
    (in-place update -> version counter, or a new tensor -> data pointer)."""
    st = _static(pred, in_ch)
    key = tuple(p._version for p in st["params"]) + (st["params"][0].data_ptr(), pred.alpha._version, pred.beta._version)
    cache = pred.__dict__.get("_ocn_head_cache")
    if cache is not None and cache[0] == key:
        return cache[1], cache[2]
    pieces = _pack_seq(pred.xcn1lin) + _pack_seq(pred.xcn2lin)
    if st["three"]:
        pieces += _pack_seq(pred.xcn3lin)
    pieces += _pack_seq(pred.xijlin) + _pack_seq(pred.lin, last_plain=True)
    with torch.no_grad():
        flat = torch.cat([t.detach().float().reshape(-1) for t in pieces]).contiguous()
        alpha = torch.sigmoid(pred.alpha).cumprod(-1)
        mix = torch.cat((alpha[:3].float(), pred.beta.detach().float().reshape(1))).contiguous()
    if flat.is_cuda:  # rare (a parameter changed): finished before any other stream may read the cached buffers
        torch.cuda.current_stream(flat.device).synchronize()
    pred.__dict__["_ocn_head_cache"] = (key, flat, mix)
    return flat, mix


def fused_head(pred, xcn1: Tensor, xcn2: Tensor, xcn3: Optional[Tensor], xij: Tensor) -> Tensor:
    L = _lib.lib()
    B, in_ch = xcn1.shape
    st = _static(pred, in_ch)
    hid, out_ch = st["hid"], st["out_ch"]
    params, mix = packed_params(pred, in_ch)
    out = torch.empty(B, out_ch, dtype=torch.float32, device=xcn1.device)
    with torch.cuda.device(xcn1.device):
        _lib.check(L.ocn_cn_head(_lib.ptr(xcn1.contiguous()), _lib.ptr(xcn2.contiguous()),
                                 _lib.ptr(None if xcn3 is None else xcn3.contiguous()), _lib.ptr(xij.contiguous()), B,
                                 in_ch, hid, out_ch, st["flags"], _lib.ptr(params), params.numel(), _lib.ptr(mix),
                                 _lib.ptr(out), _stream(xcn1.device)), "ocn_cn_head")
    return out
