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


# ---- wider heads: one tensor-core launch per Linear with its element-wise tail (csrc/linear_tc.cu) ------------------------

def _layers(seq: nn.Sequential):
    """[(Linear, LayerNorm or None, relu)] in module order (Dropout is the identity in eval mode, Identity is skipped)."""
    out = []
    for m in seq:
        if isinstance(m, nn.Linear):
            out.append([m, None, False])
        elif isinstance(m, nn.LayerNorm):
            out[-1][1] = m
        elif isinstance(m, nn.ReLU):
            out[-1][2] = True
        elif not isinstance(m, (nn.Dropout, nn.Identity)):
            raise TypeError(f"unexpected module in a predictor head: {type(m).__name__}")
    return out


def _wide_static(pred, in_ch: int):
    st = pred.__dict__.get("_ocn_wide_static")
    if st is None or st["in_ch"] != in_ch:
        three = pred.order >= 3 and hasattr(pred, "xcn3lin")
        names = [m for m in _HEAD_MODULES if m != "xcn3lin" or three]
        plan = {m: _layers(getattr(pred, m)) for m in names}
        L = _lib.lib()
        ok = True
        for m, layers in plan.items():
            hidden = layers[:-1] if m == "lin" else layers
            for lin, _, _ in hidden:
                ok = ok and L.ocn_linear_tc_prep_floats(lin.out_features, lin.in_features) > 0
        ok = ok and len(plan["lin"]) >= 2 and plan["lin"][-1][0].out_features <= 8
        st = {"in_ch": in_ch, "plan": plan, "ok": bool(ok), "names": names}
        pred.__dict__["_ocn_wide_static"] = st
    return st


def wide_supported(pred, in_ch: int) -> bool:
    """Every Linear of the head but the last has 32 / 64 / 128 / 256 outputs and a multiple of 32 inputs."""
    return _wide_static(pred, in_ch)["ok"]


def _prepped(pred, lin: nn.Linear) -> Tensor:
    """The weight of one Linear split and laid out for ocn_linear_tc; redone when the parameter changed."""
    cache = pred.__dict__.setdefault("_ocn_wide_prepped", {})
    key = (lin.weight.data_ptr(), lin.weight._version)
    hit = cache.get(id(lin))
    if hit is not None and hit[0] == key:
        return hit[1]
    L = _lib.lib()
    n, k = lin.out_features, lin.in_features
    w = lin.weight.detach().float().contiguous()
    buf = torch.empty(L.ocn_linear_tc_prep_floats(n, k), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(L.ocn_linear_tc_prep(_lib.ptr(w), n, k, _lib.ptr(buf), _stream(w.device)), "ocn_linear_tc_prep")
        torch.cuda.current_stream(w.device).synchronize()    # rare (a parameter changed): other streams may read the cache
    cache[id(lin)] = (key, buf)
    return buf


def linear_tc(pred, x: Tensor, lin: nn.Linear, ln: Optional[nn.LayerNorm] = None, relu: bool = False, want_out: bool = True,
              z: Optional[Tensor] = None, z_scale: float = 1.0, z_accumulate: bool = False,
              final: Optional[nn.Linear] = None):
    """One layer on the tensor cores: returns (out or None, out_final or None); ``z`` is updated in place."""
    L = _lib.lib()
    x = x.contiguous()
    B, k = x.shape
    n = lin.out_features
    dev = x.device
    out = torch.empty(B, n, dtype=torch.float32, device=dev) if want_out else None
    fin = torch.empty(B, final.out_features, dtype=torch.float32, device=dev) if final is not None else None
    f32 = lambda t: None if t is None else t.detach().float().contiguous()
    bias, g, be = f32(lin.bias), f32(None if ln is None else ln.weight), f32(None if ln is None else ln.bias)
    wo, bo = f32(None if final is None else final.weight), f32(None if final is None else final.bias)
    with torch.cuda.device(dev):
        _lib.check(L.ocn_linear_tc(_lib.ptr(x), B, k, n, _lib.ptr(_prepped(pred, lin)), _lib.ptr(bias), _lib.ptr(g), _lib.ptr(be),
                                   int(relu), _lib.ptr(out), _lib.ptr(z), float(z_scale), int(z_accumulate), _lib.ptr(wo),
                                   _lib.ptr(bo), 0 if final is None else final.out_features, _lib.ptr(fin), _stream(dev)),
                   "ocn_linear_tc")
    return out, fin


def fused_head_wide(pred, xcn1: Tensor, xcn2: Tensor, xcn3: Optional[Tensor], xij: Tensor) -> Tensor:
    """The head of ``fused_head`` for hidden widths up to 256: 11 - 13 ``ocn_linear_tc`` launches (bias / LayerNorm / ReLU,
    the branch mix ``z += a_k * branch`` and the last Linear ride in the epilogues)."""
    st = _wide_static(pred, xcn1.shape[1])
    plan = st["plan"]
    with torch.no_grad():
        alpha = torch.sigmoid(pred.alpha).cumprod(-1)
        mix = torch.cat((alpha[:3].float(), pred.beta.detach().float().reshape(1))).tolist()
    B = xcn1.shape[0]
    hid = plan["lin"][0][0].in_features
    z = torch.empty(B, hid, dtype=torch.float32, device=xcn1.device)
    first = True
    for name, x, w in (("xcn1lin", xcn1, mix[0]), ("xcn2lin", xcn2, mix[1]), ("xcn3lin", xcn3, mix[2]), ("xijlin", xij, mix[3])):
        if name not in plan:
            continue
        layers = plan[name]
        t = x.float()
        for k, (lin, ln, relu) in enumerate(layers):
            if k == len(layers) - 1:
                linear_tc(pred, t, lin, ln, relu, want_out=False, z=z, z_scale=w, z_accumulate=not first)
                first = False
            else:
                t, _ = linear_tc(pred, t, lin, ln, relu)
    layers = plan["lin"]
    t = z
    for k, (lin, ln, relu) in enumerate(layers[:-1]):
        if k == len(layers) - 2:
            _, out = linear_tc(pred, t, lin, ln, relu, want_out=False, final=layers[-1][0])
            return out
        t, _ = linear_tc(pred, t, lin, ln, relu)
    raise AssertionError("unreachable")
