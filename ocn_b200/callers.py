"""The callers either side of the hot path: the scoring / training loops of the ppa, citation2 and
SPD drivers restated on the fused sessions.  Only what touches the path is here -- argument parsing, logging, data
loading and checkpoints stay with the reference's drivers, which call these functions where they now loop in Python.

* ``get_cn1_cn2_spd``        SPD.py:65-126    ``get_cn1_cn2`` with ``compute_adj2_with_shortest_paths``
* ``score_links``            NeighborOverlap_large_ppa.py:199-234, NeighborOverlapCitation2.py:241-254: the
                             ``torch.cat([predictor(h, adj, *get_cn1_cn2(adj, e), e, args) for perm in PermIterator(...)])``
                             loops; many link batches per session, CN sets built ONCE per batch (the ppa loop calls
                             ``get_cn1_cn2`` twice per batch)
* ``test_ppa``               NeighborOverlap_large_ppa.py:176-259 (Hits@20/50/100 on the device)
* ``test_citation2_split``   NeighborOverlapCitation2.py:236-259 (MRR of one split on the device)
* ``train_step``             NeighborOverlap_large_ppa.py:69-141 / NeighborOverlapCitation2.py:131-209: one optimiser
                             step -- ``--maskinput`` adjacency, positive then negative sub-batches, gradient of the
                             detached embedding handed back to the GNN
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import metrics
from .cn import CNSession, SparseRows, get_cn
from .graph import Graph


def get_cn1_cn2_spd(adj: Graph, tedge: Tensor) -> Tuple[SparseRows, SparseRows]:
    """SPD.py:98-126.  ``compute_adj2_with_shortest_paths`` (:65-95) multiplies the 2-walk counts ``Ej . A`` by a
    mask that removes the entries of ``A`` -- a node adjacent to the destination is at distance 1, not 2 -- and keeps
    the zeroed entries in the matrix; CN2 = ``Ei (.) Ej2`` therefore has the pattern of the plain CN2 with value 0 on
    the common neighbours (k in N(i), k in N(j)) and the walk count elsewhere.

    The reference's loop compares the row index of ``Ej . A`` (a position in the batch) with row indices of ``A``
    (node ids), and its last line builds a torch_sparse ``SparseTensor`` with pygho's argument list (SPD.py:93:
    TypeError) -- the function does not run as written.  What is built here is the masking the comments and the name
    describe: by the destination NODE of every link."""
    cn1, cn2 = get_cn(adj, tedge, 2, True)
    n = adj.n_cols
    k1, k2 = cn1.row() * n + cn1.col, cn2.row() * n + cn2.col
    val = cn2.value.clone()
    if k1.numel() and k2.numel():
        idx = torch.searchsorted(k1, k2).clamp_(max=k1.numel() - 1)
        val[k1[idx] == k2] = 0
    return cn1, SparseRows(cn2.rowptr, cn2.col, val, cn2.shape)


def _forward(predictor, h: Tensor, adj: Graph, sess: CNSession, e: Tensor, args) -> Tensor:
    if getattr(predictor, "order", 2) >= 3:
        return predictor.multidomainforward(h, adj, sess, None, None, e, args)
    if getattr(predictor, "variant", 5) == 7:
        return predictor.multidomainforward(h, adj, sess, None, e, args)
    return predictor.multidomainforward(h, adj, sess, None, e)


@torch.no_grad()
def score_links(predictor, h: Tensor, adj: Graph, edges: Tensor, batch_size: int, args=None,
                batches_per_session: int = 32) -> Tensor:
    """Scores of ``edges`` ([2, E]) in order, every ``batch_size`` consecutive links normalised as one batch (the
    reference's ``PermIterator(device, E, batch_size, False)`` order, utils.py:8-36).  Returns ``[E]`` on the device:
    the drivers' per-batch ``.cpu()`` becomes one copy at the end, or none when the metric is computed by
    ``ocn_b200.metrics``."""
    if predictor.training:
        raise RuntimeError("score_links is the evaluation loop: call predictor.eval() first (training goes through train_step)")
    E = int(edges.shape[1])
    out = torch.empty(E, dtype=torch.float32, device=h.device)
    step = batch_size * max(1, int(batches_per_session))
    order, weighted = predictor.order, predictor.weighted
    for s in range(0, E, step):
        e = edges[:, s:s + step].contiguous()
        sess = CNSession(adj, e, batch_size, order).build(order, weighted, spd=getattr(predictor, "spd", False))
        out[s:s + e.shape[1]] = _forward(predictor, h, adj, sess, e, args).reshape(-1)
    return out


@torch.no_grad()
def test_ppa(predictor, h: Tensor, adj: Graph, split_edge: Dict[str, Dict[str, Tensor]], batch_size: int, args=None,
             h_test: Optional[Tensor] = None, adj_test: Optional[Graph] = None, ks: Sequence[int] = (20, 50, 100)):
    """``test()`` of NeighborOverlap_large_ppa.py:176-259: validation and test links scored in batches of
    ``batch_size``, Hits@K on the device.  As in the reference, "train" hits are the validation hits (:240-243).
    ``h_test`` / ``adj_test``: the embedding and graph that include the validation edges (``use_valedges_as_input``)."""
    dev = h.device
    edge = lambda split, key: split_edge[split][key].to(dev).t().contiguous()
    pos_valid = score_links(predictor, h, adj, edge("valid", "edge"), batch_size, args)
    neg_valid = score_links(predictor, h, adj, edge("valid", "edge_neg"), batch_size, args)
    ht, at = (h_test if h_test is not None else h), (adj_test if adj_test is not None else adj)
    pos_test = score_links(predictor, ht, at, edge("test", "edge"), batch_size, args)
    neg_test = score_links(predictor, ht, at, edge("test", "edge_neg"), batch_size, args)
    return metrics.evaluate_hits(pos_valid, pos_valid, neg_valid, pos_test, neg_test, ks)


@torch.no_grad()
def test_citation2_split(predictor, h: Tensor, adj: Graph, source: Tensor, target: Tensor, target_neg: Tensor,
                         batch_size: int, args=None) -> Tensor:
    """``test_split`` of NeighborOverlapCitation2.py:236-259: every source against its target and its ``K`` negative
    targets (``target_neg`` [S, K]); returns the mean reciprocal rank as a 0-dim device tensor."""
    dev = h.device
    source, target, target_neg = source.to(dev), target.to(dev), target_neg.to(dev)
    pos = score_links(predictor, h, adj, torch.stack((source, target)), batch_size, args)
    K = target_neg.shape[1]
    neg_e = torch.stack((source.view(-1, 1).repeat(1, K).view(-1), target_neg.reshape(-1)))
    neg = score_links(predictor, h, adj, neg_e, batch_size, args).view(-1, K)
    return metrics.mrr_list(pos, neg).mean()


def train_step(model: Callable[[Tensor, Graph], Tensor], predictor, x: Tensor, adj: Graph, pos_edge: Tensor,
               neg_edge: Tensor, linkbatchsize: int, maskinput: bool = True, args=None, backprop_gnn: bool = True):
    """One optimiser step's forward / backward (the caller owns ``optimizer.zero_grad()`` / ``step()``):

    * ``--maskinput``: the batch's own positive links leave the adjacency (``Graph.masked``: multiplicity
      decrements + one compaction instead of re-sorting the edge list, NeighborOverlap_large_ppa.py:71-80);
    * ``h0 = model(x, adj)``; the predictor sees ``h = h0.detach().requires_grad_()`` (:94-95);
    * the positive links then the negative links in sub-batches of ``linkbatchsize``, each sub-batch's loss
      ``-(1 / totallen) * logsigmoid(+-out).sum()`` back-propagated on its own (:99-131) -- the sub-batches run as
      one session of several link batches (``dist.sharded_train_step``: same running inner product per batch);
    * ``h0.backward(h.grad)`` (:137; the citation2 driver leaves this out, SURVEY Q7: ``backprop_gnn=False``).

    Returns the step loss (0-dim device tensor)."""
    from .dist import sharded_train_step
    if maskinput:
        adj = adj.masked(pos_edge)
    h0 = model(x, adj)
    h = h0.detach().requires_grad_(True)
    fill = float(getattr(args, "sum", 0.0) or 0.0)
    loss = torch.zeros((), device=x.device)
    for edge, sign in ((pos_edge, 1.0), (neg_edge, -1.0)):
        total = int(edge.shape[1])
        subs = [edge[:, s:s + linkbatchsize].contiguous() for s in range(0, total, linkbatchsize)]
        # equal-width sub-batches go through one session; a ragged tail gets its own
        full = [u for u in subs if u.shape[1] == linkbatchsize]
        for group in (full, [u for u in subs if u.shape[1] != linkbatchsize]):
            if group:
                loss = loss + sharded_train_step(predictor, h, adj, group, [sign] * len(group), total, 0, 1, fill, args)
    if backprop_gnn and h0.requires_grad:
        h0.backward(h.grad)
    return loss
