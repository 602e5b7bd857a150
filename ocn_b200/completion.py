"""The NCNC-style completion predictor ``cn2`` = ``IncompleteCN1Predictor`` (model.py:843-1146) on libocn_b200.

Same constructor arguments, parameter / buffer names and forward signature as the reference class (its base class
``CNLinkPredictor``, model.py:524-592, included: ``xijlin`` starts with a ``Linear(64, hidden)`` and is applied TWICE
at depth >= 0, model.py:576, 902, 1126, so the class only runs with ``in_channels = hidden_channels = 64``).

Data flow of one call at depth 1 (the default):

1. ``adjoverlap(adj, adj, tar_ei, calresadj=True)``: CN1 and the two residual sets ``N(i) \\ N(j)``, ``N(j) \\ N(i)``
   (``ocn_rows_intersect_*`` / ``ocn_rows_difference_*``), the residuals thinned by ``sparsesample_reweight``;
2. every residual entry ``(b, k)`` is a link ``(j_b, k)`` resp. ``(i_b, k)`` scored by the depth-0 pass of the same
   module (CN1 aggregate through ``ocn_rows_intersect_*`` + ``ocn_spmm_csr``, heads) -- ``sum_b |N(i_b) (+) N(j_b)|``
   links, the bulk of the work;
3. the scores, squashed by ``clampprob``, become the values of the residual matrices, which then go through the cn5
   normalisation / running inner product / orthogonalisation (model.py:960-1123) and three ``spmm_add``.

cn3 / cn4 (``IncompleteCN1Predictorhighorder`` / ``...SaveMemory``, below) recombine the same operators (cn3: plus A^2 and
its residual sets); all three also run unmodified on the import shim (``ocn_b200.shim``; traces in
tests/golden/ref_trace_cn2_*.pt, _cn3_eval.pt, _cn4_eval.pt).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch
import torch.nn as nn
from torch import Tensor

from .cn import SparseRows, adjoverlap
from .explicit import _hadamard_sum, _normalise_cn1, _orthogonalise
from .graph import Graph
from .predictor import DropAdj, _mlp3
from .sample import sparsesample_reweight
from .sparse_ops import spmm_add


class CNLinkPredictor(nn.Module):
    """Parameter layout of the reference's ``cn1`` base class (model.py:524-592); only what cn2 inherits is used."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout, edrop=0.0, ln=False, cndeg=-1,
                 use_xlin=False, tailact=False, twolayerlin=False, beta=1.0):
        super().__init__()
        lnfn = (lambda d: nn.LayerNorm(d)) if ln else (lambda d: nn.Identity())
        self.register_parameter("beta", nn.Parameter(beta * torch.ones((1))))
        self.dropadj = DropAdj(edrop)
        self.xlin = _mlp3(hidden_channels, hidden_channels, dropout, ln, last=False) if use_xlin else (lambda x: 0)
        self.xcnlin = _mlp3(in_channels, hidden_channels, dropout, ln, last=not tailact)
        self.xcn1lin = _mlp3(in_channels, hidden_channels, dropout, ln)
        self.xcn2lin = _mlp3(in_channels, hidden_channels, dropout, ln)
        self.xcn4lin = _mlp3(in_channels, hidden_channels, dropout, ln)
        self.xijlin = nn.Sequential(nn.Linear(64, hidden_channels), lnfn(hidden_channels),
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

    def _running_mean_update(self, s: Tensor):
        """``innerprod1`` (model.py:594-603)."""
        self.n += 1
        beta = self.n ** -1
        self.innerprod *= (1 - beta)
        self.innerprod += beta * s


class IncompleteCN1Predictor(CNLinkPredictor):
    residual_fill = 0.0   # weight of a residual column that sums to exactly 1 (model.py:963-966)
    xij_passes = 2        # xijlin is applied to the pair term this many times (model.py:902 and :1126)

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout, edrop=0.0, ln=False, cndeg=-1,
                 use_xlin=False, tailact=False, twolayerlin=False, beta=1.0, alpha=1.0, scale=5, offset=3, trainresdeg=8,
                 testresdeg=128, pt=0.5, learnablept=False, depth=1, splitsize=-1):
        super().__init__(in_channels, hidden_channels, out_channels, num_layers, dropout, edrop, ln, cndeg, use_xlin,
                         tailact, twolayerlin, beta)
        if learnablept:
            # model.py:947-949: clampprob(prob[nnz], pt[potcn[0]] of shape [nnz, 1]) broadcasts to [nnz, nnz] and that
            # matrix is then stored as the VALUES of the residual matrix -- the branch cannot run on real sizes
            raise NotImplementedError("learnablept=True: the reference's own branch broadcasts the residual weights to "
                                      "[nnz, nnz] (model.py:947-949); no README command enables it")
        self.learnablept = learnablept
        self.depth = depth
        self.splitsize = splitsize
        self.lins = nn.Sequential()
        self.register_buffer("alpha2", torch.tensor([alpha]))
        self.register_buffer("pt", torch.tensor([pt]))
        self.register_buffer("scale", torch.tensor([scale]))
        self.register_buffer("offset", torch.tensor([offset]))
        self.trainresdeg = trainresdeg
        self.testresdeg = testresdeg
        self.ptlin = nn.Sequential(nn.Linear(hidden_channels, hidden_channels), nn.ReLU(inplace=True),
                                   nn.Linear(hidden_channels, 1), nn.Sigmoid())
        self.rand_fn: Optional[Callable] = None     # source of the samplers' random numbers (default: torch.rand)

    def clampprob(self, prob, pt):
        p0 = torch.sigmoid_(self.scale * (prob - self.offset))
        return self.alpha2 * pt * p0 / (pt * p0 + 1 - p0)

    def setalpha(self, alpha: float):
        self.alpha2.fill_(alpha)

    def _sample(self, rows: SparseRows, deg: int) -> SparseRows:
        return sparsesample_reweight(rows, deg, self.rand_fn) if deg > 0 else rows

    def _score_residual(self, x, adj, ends: Tensor, res: SparseRows, filled1, depth) -> Tensor:
        ei = torch.stack((ends[res.row()], res.col))
        if self.splitsize < 0 or ei.shape[1] == 0:
            return self.forward(x, adj, ei, filled1, depth).flatten() if ei.shape[1] else x.new_zeros(0)
        out = torch.empty(ei.shape[1], dtype=torch.float32, device=x.device)
        for s in range(0, ei.shape[1], self.splitsize):
            out[s:s + self.splitsize] = self.forward(x, adj, ei[:, s:s + self.splitsize], filled1, depth).flatten()
        return out

    def multidomainforward(self, x, adj: Graph, tar_ei, filled1: bool = False, cndropprobs: Iterable[float] = [],
                           depth: Optional[int] = None):
        assert len(cndropprobs) == 0
        if depth is None:
            depth = self.depth
        adj = self.dropadj(adj)
        xij = self.xijlin(x[tar_ei[0]] * x[tar_ei[1]])
        x = x + self.xlin(x)
        resdeg = self.trainresdeg if self.training else self.testresdeg
        if depth > 0.5:
            cn, cnres1, cnres2 = adjoverlap(adj, adj, tar_ei, filled1, calresadj=True)
            cn = self._sample(cn, self.cndeg)
            cnres1, cnres2 = self._sample(cnres1, resdeg), self._sample(cnres2, resdeg)
        else:
            cn = self._sample(adjoverlap(adj, adj, tar_ei, filled1), self.cndeg)
        xcn = spmm_add(cn, x)
        if depth > 0.5:
            with torch.no_grad():
                probcn1 = self._score_residual(x, adj, tar_ei[1], cnres1, filled1, depth - 1)
                probcn2 = self._score_residual(x, adj, tar_ei[0], cnres2, filled1, depth - 1)
            probcn1 = self.clampprob(probcn1, self.pt)
            probcn2 = self.clampprob(probcn2, self.pt)
            cnres1 = SparseRows(cnres1.rowptr, cnres1.col, probcn1 * cnres1.value, cnres1.shape)
            cnres2 = SparseRows(cnres2.rowptr, cnres2.col, probcn2 * cnres2.value, cnres2.shape)
            # model.py:960-1123: the cn5 combination with the weighted residuals in the roles of cn1 / cn2
            ncn1 = _normalise_cn1(cnres1, self.residual_fill)
            if self.training:
                self._running_mean_update(_hadamard_sum(cnres2, ncn1).detach())
            ip = self.innerprod.detach().float()
            if cnres1.col.numel() + cnres2.col.numel() > 0:
                scale = ncn1.value.detach().abs().max() if ncn1.value.numel() else torch.zeros((), device=x.device)
            else:
                scale = torch.ones((), device=x.device)
            coeff = torch.where(scale > 0, ip / torch.where(scale > 0, scale, torch.ones_like(scale)), ip)
            ncn2 = _orthogonalise(cnres2, ncn1, coeff)
            xcn = xcn + spmm_add(ncn2, x) + spmm_add(ncn1, x)
        for _ in range(self.xij_passes - 1):
            xij = self.xijlin(xij)
        return self.lin(self.xcnlin(xcn) * self.beta + xij)

    def forward(self, x, adj, tar_ei, filled1: bool = False, depth: Optional[int] = None):
        if depth is None:
            depth = self.depth
        return self.multidomainforward(x, adj, tar_ei, filled1, [], depth)


class IncompleteCN1PredictorSaveMemory(IncompleteCN1Predictor):
    """cn4 (model.py:1532-1886): cn2's forward with ``del`` statements and ``torch.cuda.empty_cache()`` between its steps and
    TWO arithmetic differences: a residual column whose weights sum to exactly 1 keeps weight 1 instead of 0
    (``inv_col_sum[inv_col_sum == 1] = 1``, model.py:1700-1703), and the pair term goes through ``xijlin`` a third time
    (``xij = self.xijlin(xij)`` at :1861 and ``self.xijlin(xij)`` again inside the last line, :1863)."""
    residual_fill = 1.0
    xij_passes = 3


class IncompleteCN1Predictorhighorder(IncompleteCN1Predictor):
    """cn3 (model.py:1150-1530): CN1 = A[i] cap A[j] and CN2 = A[i] cap A^2[j] through the cn5 combination (singleton weight
    1), plus -- at depth > 0 -- the four residual sets of A and A^2 scored by the depth - 1 pass and added unnormalised.
    The reference rebuilds ``spadj @ spadj`` in EVERY call, the recursive ones included (:1211-1212); here A^2 is built once
    per top-level call (``ocn_spgemm_a2_*``) and handed down.  Same signature: ``forward(x, adj, cn1, cn2, tar_ei, ...)``
    (the ``cn1`` / ``cn2`` arguments are unused by the reference too)."""

    def _weighted(self, x, adj, adj2, ends: Tensor, res: SparseRows, filled1, depth) -> SparseRows:
        ei = torch.stack((ends[res.row()], res.col))
        with torch.no_grad():
            if ei.shape[1] == 0:
                prob = x.new_zeros(0)
            elif self.splitsize < 0:
                prob = self.multidomainforward(x, adj, None, None, ei, filled1, [], depth, _adj2=adj2).flatten()
            else:
                prob = torch.empty(ei.shape[1], dtype=torch.float32, device=x.device)
                for s in range(0, ei.shape[1], self.splitsize):
                    prob[s:s + self.splitsize] = self.multidomainforward(x, adj, None, None, ei[:, s:s + self.splitsize], filled1,
                                                                         [], depth, _adj2=adj2).flatten()
        return SparseRows(res.rowptr, res.col, self.clampprob(prob, self.pt) * res.value, res.shape)

    def multidomainforward(self, x, adj: Graph, cn1, cn2, tar_ei, filled1: bool = False, cndropprobs: Iterable[float] = [],
                           depth: Optional[int] = None, _adj2: Optional[Graph] = None):
        from .sparse_ops import spgemm_a2
        assert len(cndropprobs) == 0
        if depth is None:
            depth = self.depth
        adj = self.dropadj(adj)
        xij = self.xijlin(x[tar_ei[0]] * x[tar_ei[1]])
        x = x + self.xlin(x)
        adj2 = _adj2 if _adj2 is not None else spgemm_a2(adj, 0, False)
        resdeg = self.trainresdeg if self.training else self.testresdeg
        if depth > 0.5:
            cn, cnres1, cnres2 = adjoverlap(adj, adj, tar_ei, filled1, calresadj=True)
            cn = self._sample(cn, self.cndeg)
            cnres1, cnres2 = self._sample(cnres1, resdeg), self._sample(cnres2, resdeg)
            cn22, cn2res1, cn2res2 = adjoverlap(adj, adj2, tar_ei, filled1, calresadj=True)
            cn22 = self._sample(cn22, self.cndeg)
            cn2res1, cn2res2 = self._sample(cn2res1, resdeg), self._sample(cn2res2, resdeg)
        else:
            cn = self._sample(adjoverlap(adj, adj, tar_ei, filled1), self.cndeg)
            cn22, d1, d2 = adjoverlap(adj, adj2, tar_ei, filled1, calresadj=True)
            cn22 = self._sample(cn22, self.cndeg)
            self._sample(d1, resdeg), self._sample(d2, resdeg)      # built, sampled and dropped, as the reference does (:1240-1244)
        # model.py:1245-1409: the cn5 combination of (cn, cn22), a column summing to exactly 1 keeps weight 1
        ncn1 = _normalise_cn1(cn, 1.0)
        if self.training:
            self._running_mean_update(_hadamard_sum(cn22, ncn1).detach())
        ip = self.innerprod.detach().float()
        if cn.col.numel() + cn22.col.numel() > 0:
            scale = ncn1.value.detach().abs().max() if ncn1.value.numel() else torch.zeros((), device=x.device)
        else:
            scale = torch.ones((), device=x.device)
        coeff = torch.where(scale > 0, ip / torch.where(scale > 0, scale, torch.ones_like(scale)), ip)
        ncn2 = _orthogonalise(cn22, ncn1, coeff)
        xcn_a, xcn_b = spmm_add(ncn1, x), spmm_add(ncn2, x)
        if depth > 0.5:
            w1 = self._weighted(x, adj, adj2, tar_ei[1], cnres1, filled1, depth - 1)
            w2 = self._weighted(x, adj, adj2, tar_ei[0], cnres2, filled1, depth - 1)
            xcn_a = xcn_a + spmm_add(w2, x) + spmm_add(w1, x)
            v1 = self._weighted(x, adj, adj2, tar_ei[1], cn2res1, filled1, depth - 1)
            v2 = self._weighted(x, adj, adj2, tar_ei[0], cn2res2, filled1, depth - 1)
            xcn_b = xcn_b + spmm_add(v2, x) + spmm_add(v1, x)
        xij = self.xijlin(xij)
        return self.lin(self.xcnlin(xcn_a) * self.beta + self.xcnlin(xcn_b) * self.beta + xij)

    def forward(self, x, adj, cn1, cn2, tar_ei, filled1: bool = False, depth: Optional[int] = None):
        if depth is None:
            depth = self.depth
        return self.multidomainforward(x, adj, cn1, cn2, tar_ei, filled1, [], depth)
