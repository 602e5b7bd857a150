"""Host-side mirror of the reference predictors on top of the fused CUDA path.

Same constructor arguments, parameter / buffer names (so ``state_dict`` interchanges with the
reference's ``gmodel/*.pt`` checkpoints) and forward signatures as

* ``CNLinkPredictorOringin``  = ``cn5`` = OCN      (model.py:2171-2443)
* ``CNLinkPredictor3hopCNs``  = ``cn6``            (model.py:2447-2954; the order-3 template)
* ``CNLinkPredictorbaselearn``= ``cn7`` = OCNP     (model.py:2958-3229)

The dense MLP heads stay torch modules; everything between the target links and
the three ``[B, F]`` aggregates runs in libocn_b200.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn
from torch import Tensor

from .cn import CNSession, SparseRows
from .graph import Graph


class DropAdj(nn.Module):
    """model.py:211-229.  cn5 / cn6 / cn7 build one but never apply it; the GNN applies its own to the
    adjacency of every layer in a training forward (model.py:312) -- ``forward`` does that on a ``Graph``."""

    def __init__(self, dp: float = 0.0, doscale=True):
        super().__init__()
        self.dp = dp
        self.register_buffer("ratio", torch.tensor(1 / (1 - dp)))
        self.doscale = doscale

    def forward(self, adj: Graph) -> Graph:
        from .sparse_ops import drop_adj
        return drop_adj(adj, self.dp, self.training, self.doscale)


def _mlp3(i, h, dropout, ln, last=True):
    lnfn = (lambda d: nn.LayerNorm(d)) if ln else (lambda d: nn.Identity())
    return nn.Sequential(nn.Linear(i, h), nn.Dropout(dropout, inplace=True), nn.ReLU(inplace=True),
                         nn.Linear(h, h), lnfn(h), nn.Dropout(dropout, inplace=True), nn.ReLU(inplace=True),
                         nn.Linear(h, h) if last else nn.Identity())


class _CNAggregateFn(torch.autograd.Function):
    """xcn_k = C-hat_k @ x and x_i * x_j; the weights do not depend on x or on any parameter."""

    @staticmethod
    def forward(ctx, x, sess: CNSession, variant: int, fill: float, ip3: Optional[Tensor]):
        # ip3 None: every batch of the session reads its own coefficients from its scalar slots (CNSession.set_batch_ip)
        xcn1, xcn2, xcn3, xij = sess.aggregate(x, variant, fill, ip3, want_xij=True)
        ctx.sess, ctx.variant, ctx.fill = sess, variant, fill
        ctx.per_batch = ip3 is None
        ctx.save_for_backward(x, ip3 if ip3 is not None else x.new_empty(0))
        ctx.has = (xcn2 is not None, xcn3 is not None)
        z = lambda t: t if t is not None else x.new_zeros(0)
        return xcn1, z(xcn2), z(xcn3), xij

    @staticmethod
    def backward(ctx, g1, g2, g3, gij):
        x, ip3 = ctx.saved_tensors
        if ctx.per_batch:
            ip3 = None
        gx = torch.zeros_like(x)
        ctx.sess.aggregate_bwd(x, ctx.variant, ctx.fill, ip3, g1, g2 if ctx.has[0] else None,
                               g3 if ctx.has[1] else None, gij, gx)
        if not ctx.sess._released:
            ctx.sess.release()
        return gx, None, None, None, None


class _OCNBase(nn.Module):
    variant = 5
    order = 2

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout, edrop=0.0, ln=False,
                 cndeg=-1, use_xlin=False, tailact=False, twolayerlin=False, beta=1.0, weighted: bool = False):
        super().__init__()
        lnfn = (lambda d: nn.LayerNorm(d)) if ln else (lambda d: nn.Identity())
        self.register_parameter("beta", nn.Parameter(beta * torch.ones((1))))
        self.dropadj = DropAdj(edrop)
        self.xlin = _mlp3(hidden_channels, hidden_channels, dropout, ln, last=False) if use_xlin else (lambda x: 0)
        self.xcnlin = _mlp3(in_channels, hidden_channels, dropout, ln, last=not tailact)
        self.xcn1lin = _mlp3(in_channels, hidden_channels, dropout, ln)
        self.xcn2lin = _mlp3(in_channels, hidden_channels, dropout, ln)
        self._extra_heads(in_channels, hidden_channels, dropout, ln)
        self.xijlin = nn.Sequential(nn.Linear(in_channels, hidden_channels), lnfn(hidden_channels),
                                    nn.Dropout(dropout, inplace=True), nn.ReLU(inplace=True),
                                    nn.Linear(hidden_channels, hidden_channels) if not tailact else nn.Identity())
        self.lin = nn.Sequential(nn.Linear(hidden_channels, hidden_channels), lnfn(hidden_channels),
                                 nn.Dropout(dropout, inplace=True), nn.ReLU(inplace=True),
                                 nn.Linear(hidden_channels, hidden_channels) if twolayerlin else nn.Identity(),
                                 lnfn(hidden_channels) if twolayerlin else nn.Identity(),
                                 nn.Dropout(dropout, inplace=True) if twolayerlin else nn.Identity(),
                                 nn.ReLU(inplace=True) if twolayerlin else nn.Identity(),
                                 nn.Linear(hidden_channels, out_channels))
        self.cndeg = cndeg
        self.register_parameter("alpha", nn.Parameter(torch.ones((3))))
        self.register_buffer("innerprod", torch.tensor([0.0]))
        self.n = 0
        # values of CN_k for k >= 2: pygho walk counts (citation2/ppa drivers) or 0/1 (the _large drivers), SURVEY Q11
        self.weighted = weighted
        self.spd = False  # SPD.py's get_cn1_cn2: CN2 counts only for nodes at distance exactly 2 from the destination

    def _extra_heads(self, i, h, dropout, ln):
        self.xcn4lin = _mlp3(i, h, dropout, ln)

    # -- the fused CN stage -------------------------------------------------------------------
    def _running_mean_update(self, s: Tensor):
        """``innerprod1`` (model.py:2241-2250): n += 1; ip = ip*(1-1/n) + s/n, on device, no sync."""
        self.n += 1
        beta = self.n ** -1
        self.innerprod *= (1 - beta)
        self.innerprod += beta * s

    def batch_inner_product(self, sess: CNSession, fill: float = 0.0) -> Tensor:
        """``s`` of ``innerprod1`` (model.py:2244) for a one-batch session: sum(C2 * C1-hat).  It depends on the
        CN sets and the column statistics only, not on the running mean -- which is what lets the sub-batches
        of one optimiser step run on different ranks (``ocn_b200.dist.sharded_train_step``)."""
        ip3 = self.innerprod.detach().float().repeat(3).contiguous()
        return sess.stats(5, fill, ip3, 0)[0, 1].detach().clone()

    def _ip3(self) -> Tensor:
        """The inner-product buffer broadcast to the three coefficient slots of the kernels; rebuilt only when the
        buffer changed (three tiny launches per call otherwise, in an evaluation loop that never updates it)."""
        key = (self.innerprod.data_ptr(), self.innerprod._version)
        c = self.__dict__.get("_ip3_cache")
        if c is None or c[0] != key:
            c = (key, self.innerprod.detach().float().reshape(-1)[:1].repeat(3).contiguous())
            if c[1].is_cuda:  # rare: finished before any other stream may read the cached tensor
                torch.cuda.current_stream(c[1].device).synchronize()
            self.__dict__["_ip3_cache"] = c
        return c[1]

    def cn_stage(self, x: Tensor, adj: Graph, tar_ei: Tensor, fill: float = 0.0, sess: Optional[CNSession] = None,
                 ip: Optional[Tensor] = None, per_batch_ip: bool = False):
        """``ip``: use this inner-product coefficient (shape [1]) instead of reading / updating the module's
        running mean -- the caller has already folded this batch's ``s`` into it.  ``per_batch_ip``: the session
        holds several batches and the caller has stored every batch's coefficient with ``sess.set_batch_ip`` (after
        ``sess.stats(..., stage 0)``): nothing is read from or written to the module's buffer."""
        if per_batch_ip:
            if self.variant == 5 and self.order >= 3:
                raise ValueError("per-batch coefficients are supported for order 2 (SURVEY Q9: the order-3 template chains "
                                 "three buffer updates per batch)")
            xcn1, xcn2, xcn3, xij = _CNAggregateFn.apply(x, sess, self.variant, float(fill), None)
            if not (torch.is_grad_enabled() and x.requires_grad):
                sess.release()
            return xcn1, xcn2, (xcn3 if xcn3.numel() else None), xij, sess
        if sess is None:
            sess = CNSession(adj, tar_ei, None, self.order)
            sess.build(self.order, self.weighted, spd=self.spd)
        ip3 = self._ip3() if ip is None else ip.detach().float().reshape(-1)[:1].repeat(3).contiguous()
        if self.variant == 5 and ip is not None:
            if self.order >= 3:
                raise ValueError("an explicit coefficient is supported for order 2 (cn5); the order-3 template chains "
                                 "three updates per batch through the buffer (SURVEY Q9)")
            sess.stats(5, fill, ip3, 0)
        elif self.variant == 5:
            if self.training and sess.nb != 1:
                raise ValueError("training updates the inner-product running mean once per link batch; "
                                 "pass one batch per call (the reference does, NeighborOverlapCitation2.py:162-179)")
            bs = sess.stats(5, fill, ip3, 0)
            if self.training:
                self._running_mean_update(bs[0, 1].detach())
                ip3 = self.innerprod.detach().float().repeat(3).contiguous()
                if self.order >= 3:
                    bs = sess.stats(5, fill, ip3, 1)
                    self._running_mean_update(bs[0, 2].detach())
                    # the second order-3 inner product is taken against C2-hat built with ip3[0]; its own
                    # coefficient and the first one both read the buffer after this last update
                    self._running_mean_update(bs[0, 3].detach())
                    final = self.innerprod.detach().float()
                    ip3 = torch.cat((ip3[:1], final, final)).contiguous()
        xcn1, xcn2, xcn3, xij = _CNAggregateFn.apply(x, sess, self.variant, float(fill), ip3)
        if not (torch.is_grad_enabled() and x.requires_grad):
            sess.release()  # nothing will run backward through this session: hand the statistics buffer back
        return xcn1, xcn2, (xcn3 if xcn3.numel() else None), xij, sess

    fuse_head = True  # inference: one fused kernel for hidden widths 32 / 64 (csrc/head.cu, head_tc.cu) ...
    fuse_wide_head = True  # ... one tensor-core launch per layer up to hidden 256 (csrc/linear_tc.cu); else the torch modules

    def _head(self, xcn1, xcn2, xcn3, xij):
        if self.fuse_head and not self.training and not torch.is_grad_enabled() and xcn1.is_cuda:
            from . import head
            if head.supported(self, xcn1.shape[1]) > 0:
                return head.fused_head(self, xcn1, xcn2, xcn3, xij)
            if self.fuse_wide_head and xcn1.shape[0] >= 4096 and head.wide_supported(self, xcn1.shape[1]):  # (smaller: launch-bound)
                return head.fused_head_wide(self, xcn1, xcn2, xcn3, xij)
        xij = self.xijlin(xij)
        xcn1 = self.xcn1lin(xcn1)
        xcn2 = self.xcn2lin(xcn2)
        alpha = torch.sigmoid(self.alpha).cumprod(-1)
        z = alpha[0] * xcn1 + alpha[1] * xcn2
        if xcn3 is not None:
            z = z + alpha[2] * self.xcn3lin(xcn3)
        return self.lin(z + self.beta * xij)


class CNLinkPredictorOringin(_OCNBase):
    """cn5 / OCN.  ``cn1``/``cn2`` may be None (the sets are built from ``adj`` and ``tar_ei``
    inside the fused path) or a prepared ``CNSession``."""
    variant, order = 5, 2

    def multidomainforward(self, x, adj, cn1, cn2, tar_ei, filled1: bool = False, cndropprobs: Iterable[float] = []):
        if isinstance(cn1, SparseRows) and isinstance(cn2, SparseRows):
            # the reference's own data flow: explicit adjoverlap outputs (incl. the folded adj2byblock matrix)
            from .explicit import cn5_explicit
            xcn1, xcn2, xij = cn5_explicit(self, x, cn1, cn2, tar_ei)
            return self._head(xcn1, xcn2, None, xij)
        sess = cn1 if isinstance(cn1, CNSession) else None
        xcn1, xcn2, _, xij, _ = self.cn_stage(x, adj, tar_ei, 0.0, sess)
        return self._head(xcn1, xcn2, None, xij)

    def forward(self, x, adj, cn1, cn2, tar_ei, filled1: bool = False):
        return self.multidomainforward(x, adj, cn1, cn2, tar_ei, filled1, [])


class CNLinkPredictor3hopCNs(_OCNBase):
    """cn6: order-3 OCN (north_star "depth 3")."""
    variant, order = 5, 3

    def _extra_heads(self, i, h, dropout, ln):
        self.xcn3lin = _mlp3(i, h, dropout, ln)

    def multidomainforward(self, x, adj, cn1, cn2, cn3, tar_ei, args=None, cndropprobs: Iterable[float] = []):
        sess = cn1 if isinstance(cn1, CNSession) else None
        xcn1, xcn2, xcn3, xij, _ = self.cn_stage(x, adj, tar_ei, 0.0, sess)
        return self._head(xcn1, xcn2, xcn3, xij)

    def forward(self, x, adj, cn1, cn2, cn3, tar_ei, args=None):
        return self.multidomainforward(x, adj, cn1, cn2, cn3, tar_ei, args)


class CNLinkPredictorbaselearn(_OCNBase):
    """cn7 / OCNP: singletons weigh ``args.sum``, CN2 enters raw."""
    variant, order = 7, 2

    def multidomainforward(self, x, adj, cn1, cn2, tar_ei, args, filled1: bool = False,
                           cndropprobs: Iterable[float] = []):
        fill = float(getattr(args, "sum", args if isinstance(args, (int, float)) else 0.0))
        if isinstance(cn1, SparseRows) and isinstance(cn2, SparseRows):
            from .explicit import cn7_explicit
            xcn1, xcn2, xij = cn7_explicit(self, x, cn1, cn2, tar_ei, fill)
            return self._head(xcn1, xcn2, None, xij)
        sess = cn1 if isinstance(cn1, CNSession) else None
        xcn1, xcn2, _, xij, _ = self.cn_stage(x, adj, tar_ei, fill, sess)
        return self._head(xcn1, xcn2, None, xij)

    def forward(self, x, adj, cn1, cn2, tar_ei, filled1=False):
        # the reference's forward passes `args` in the filled1 slot (model.py:3228-3229)
        return self.multidomainforward(x, adj, cn1, cn2, tar_ei, filled1, [])


predictor_dict = {
    "cn5": CNLinkPredictorOringin,
    "cn6": CNLinkPredictor3hopCNs,
    "cn7": CNLinkPredictorbaselearn,
}
