"""Multi-GPU plumbing: the graph and features are replicated on every rank, the
reference's own batch sequence (PermIterator order, utils.py:8-36) is dealt round-robin -- batch t goes
to rank t mod world -- and every rank runs the full fused path on its batches (``deal_batches``); for long streams whose pieces differ in
cost, ``predicted_walk_cost`` + ``deal_by_cost`` deal equal numbers of slices with balanced predicted work instead.
A batch is never split (its links are coupled through the column statistics), so scores are identical for any
world size and any dealing.
The only collective is the gather of fp32 scores."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def deal_batches(num_links: int, batch_size: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) link ranges of the batches owned by ``rank``."""
    nb = (num_links + batch_size - 1) // batch_size
    return [(b * batch_size, min(num_links, (b + 1) * batch_size)) for b in range(rank, nb, world)]


def predicted_walk_cost(rowptr: torch.Tensor, col: torch.Tensor, src: torch.Tensor, slice_links: int,
                        batch_size: int) -> torch.Tensor:
    """int64 [ceil(T / slice_links)]: for every slice of ``slice_links`` consecutive links of the stream, the number
    of index entries its order-3 build writes when the slice is scored as one session -- sum over the runs (maximal
    pieces of one source inside the slice, as ocn_cn_plan cuts them; a run may cross a link-batch boundary, so
    ``batch_size`` no longer enters) of sum_{k in N(src)} d(k).  The time of a slice follows this count
    (profiles/r01_step_times_v22.txt: 0.56 ms + 0.82 ms per million entries at 65 536 links), which makes it the
    weight for dealing slices to ranks.  Pure integer torch ops on whatever device the graph is on; exact, so every
    rank computes the same numbers without talking to the others."""
    T = src.numel()
    deg = rowptr[1:] - rowptr[:-1]
    t = torch.arange(T, device=src.device)
    first = (t % slice_links == 0)
    first[1:] |= src[1:] != src[:-1]
    heads = src[first]
    # F[i] = sum of the degrees of i's neighbours, for the run heads only (rows gathered, not the whole matrix)
    uniq, inv = torch.unique(heads, return_inverse=True)
    cnt = deg[uniq]
    owner = torch.repeat_interleave(torch.arange(uniq.numel(), device=src.device), cnt)
    pos = torch.arange(int(cnt.sum()), device=src.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt) \
        + torch.repeat_interleave(rowptr[uniq], cnt)
    F = torch.zeros(uniq.numel(), dtype=torch.int64, device=src.device).index_add_(0, owner, deg[col[pos].long()])
    num_slices = (T + slice_links - 1) // slice_links
    slice_of_head = torch.div(t[first], slice_links, rounding_mode="floor")
    return torch.zeros(num_slices, dtype=torch.int64, device=src.device).index_add_(0, slice_of_head, F[inv])


def deal_by_cost(costs: Sequence[int], world: int, per_rank: int) -> List[List[int]]:
    """Slices -> ranks, ``per_rank`` slices each, total predicted cost balanced: longest-processing-time greedy
    (heaviest slice first, to the least loaded rank that still has room; ties by lower index, so the result is the
    same on every rank).  Every rank returns its slices in ascending stream order.  A slice is a whole number of
    link batches, so scores do not depend on the dealing."""
    if len(costs) != world * per_rank:
        raise ValueError(f"{len(costs)} slices cannot be dealt as {per_rank} to each of {world} ranks")
    load, owned = [0] * world, [[] for _ in range(world)]
    for sl in sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i)):
        r = min((r for r in range(world) if len(owned[r]) < per_rank), key=lambda r: (load[r], r))
        owned[r].append(sl)
        load[r] += int(costs[sl])
    return [sorted(o) for o in owned]


def gather_scores(local_scores: Sequence[torch.Tensor], owned: Sequence[Tuple[int, int]], num_links: int) -> torch.Tensor:
    """All ranks contribute the scores of their batches; every rank receives the full [num_links] vector
    in link order (fp32, 4 B per link over NCCL/NVLink, or gloo on CPU).

    One ``all_gather`` of the ranks' [start, end) lists (a few integers) and one ``all_gather_into_tensor`` of the
    scores, padded to the longest rank: every link travels once (an all-reduce of a zero-padded full-length vector, the
    first version, moved ``world`` times as much)."""
    dev = local_scores[0].device if len(local_scores) else torch.device("cpu")
    world = _world()
    mine = torch.cat([sc.reshape(-1).float() for sc in local_scores]) if len(local_scores) else torch.zeros(0, device=dev)
    full = torch.empty(num_links, dtype=torch.float32, device=dev)
    if world <= 1:
        o = 0
        for (s, e) in owned:
            full[s:e] = mine[o:o + (e - s)]
            o += e - s
        return full
    ranges = [None] * world
    dist.all_gather_object(ranges, [(int(s), int(e)) for (s, e) in owned])
    longest = max(sum(e - s for (s, e) in r) for r in ranges)
    host_hop = mine.is_cuda and dist.get_backend() == "gloo"     # (CPU tests, or two ranks sharing one GPU)
    send = torch.zeros(longest, dtype=torch.float32, device="cpu" if host_hop else dev)
    send[:mine.numel()] = mine
    recv = torch.empty(world * longest, dtype=torch.float32, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.to(dev).view(world, longest)
    for r, rr in enumerate(ranges):
        o = 0
        for (s, e) in rr:
            full[s:e] = recv[r, o:o + (e - s)]
            o += e - s
    return full


# ---- training step ---------------------------------------------------------------
# NeighborOverlapCitation2.py:131-209: one optimiser step = the positive sub-batches then the negative
# sub-batches of one permutation batch (16 384 links in sub-batches of 2048), every sub-batch calling
# backward() on its share of the loss, gradients accumulating until optimizer.step().  The sub-batches are
# independent given (A, h, weights) EXCEPT for cn5's running mean of inner products, which every training
# forward updates (model.py:2241-2250): sub-batch u uses ip_u = mean(s_1 .. s_u).  s_u depends only on the
# CN sets and column sums of sub-batch u, so the ranks compute their s_u first, exchange the scalars,
# replay the running mean in sequence order and only then run the weighted aggregation -- one scalar
# exchange plus the gradient all-reduce per step, results independent of the world size.

def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _all_reduce_sum(t: torch.Tensor) -> None:
    """In-place sum over the ranks: NCCL on device memory; the gloo backend (CPU tests, or two ranks sharing one
    GPU in the parity test) goes through a host copy."""
    if _world() <= 1:
        return
    if t.is_cuda and dist.get_backend() == "gloo":
        c = t.cpu()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        t.copy_(c)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def exchange_batch_scalars(local: torch.Tensor, owned_ids: Sequence[int], num_batches: int) -> torch.Tensor:
    """``local[k]`` is the scalar of sub-batch ``owned_ids[k]``; returns all ``num_batches`` scalars in
    sequence order on every rank (disjoint supports: the sum is a gather)."""
    full = torch.zeros(num_batches, dtype=torch.float32, device=local.device)
    if len(owned_ids):
        full[torch.as_tensor(list(owned_ids), device=local.device, dtype=torch.long)] = local.float().reshape(-1)
    _all_reduce_sum(full)
    return full


def replay_running_mean(s_all: torch.Tensor, ip0: torch.Tensor, n0: int) -> Tuple[torch.Tensor, int]:
    """``innerprod1`` in training mode (model.py:2245-2248) replayed over the sub-batches in sequence:
    n += 1; ip = ip * (1 - 1/n) + s / n, in fp32 exactly as the module does it.  Returns the coefficient every
    sub-batch sees (``[num_batches]``; the last one is the buffer's value after the step) and the final n."""
    ip = ip0.detach().float().reshape(-1)[:1].clone()
    out = torch.empty(s_all.numel(), dtype=torch.float32, device=s_all.device)
    n = int(n0)
    for u in range(s_all.numel()):
        n += 1
        beta = n ** -1
        ip *= (1 - beta)
        ip += beta * s_all[u]
        out[u] = ip[0]
    return out, n


def allreduce_gradients(tensors: Iterable[torch.Tensor]) -> None:
    """Sum the gradients of ``tensors`` (parameters, or leaves such as the detached ``h`` of
    NeighborOverlapCitation2.py:155) over the ranks in ONE flat bucket.  The reference scales every sub-batch
    loss by 1/len(step) before backward (:177, :200), so the plain sum is the step's gradient; tensors without
    a gradient on this rank contribute zeros (and receive the sum)."""
    ts = [t for t in tensors if t.requires_grad]
    if _world() <= 1 or not ts:
        return
    for t in ts:
        if t.grad is None:
            t.grad = torch.zeros_like(t)
    # large gradients (the [N, F] gradient of the detached embedding: 375 MB at citation2 shape, F = 32) are reduced in
    # place, one collective each; the predictor's many small tensors share one flat bucket
    small = []
    for t in ts:
        if t.grad.numel() >= (1 << 18) and t.grad.is_contiguous() and t.grad.dtype == torch.float32:
            _all_reduce_sum(t.grad)
        else:
            small.append(t)
    if not small:
        return
    flat = torch.cat([t.grad.reshape(-1).float() for t in small])
    _all_reduce_sum(flat)
    o = 0
    for t in small:
        k = t.numel()
        t.grad.copy_(flat[o:o + k].view_as(t.grad))
        o += k


def sharded_train_step(pred, h: torch.Tensor, graph, sub_batches: Sequence[torch.Tensor], signs: Sequence[float],
                       total_len: int, rank: int, world: int, fill: float = 0.0, args=None,
                       fuse_sub_batches: bool = True):
    """The predictor part of one optimiser step of the citation2 / ppa drivers with its sub-batches dealt to
    ranks (sub-batch u -> rank u mod world).  ``sub_batches[u]`` is ``[2, b]`` links, ``signs[u]`` is +1 for a
    positive and -1 for a negative sub-batch (loss = -(1/total_len) * sum(logsigmoid(sign * out)),
    NeighborOverlapCitation2.py:177,200).  ``h`` is the detached node embedding with requires_grad.  After the
    call every rank holds the summed gradients of the predictor parameters and of ``h``, and the predictor's
    inner-product buffer / counter in the state the sequential loop leaves them in.  Returns the step loss.

    ``fuse_sub_batches``: the owned sub-batches (equal width) run as ONE session of several link batches instead of
    one session each -- same column statistics per batch, same coefficients, same gradients up to fp32 summation
    order (and up to the order in which dropout draws its masks, when the heads use dropout)."""
    import torch.nn.functional as F
    from .cn import CNSession
    U = len(sub_batches)
    mine = list(range(rank, U, world))
    ocn5 = getattr(pred, "variant", 5) == 5
    dev = h.device
    widths = {int(sub_batches[u].shape[1]) for u in mine}
    fused = fuse_sub_batches and len(mine) > 1 and len(widths) == 1 and not (ocn5 and pred.order >= 3)
    loss = torch.zeros((), device=dev)
    if fused:
        # The owned sub-batches as ONE session of len(mine) link batches: one plan, one build, one aggregation forward
        # and backward for all of them (16 x fewer launches, kernels 16 x larger); every batch still has its own column
        # statistics, and its own inner-product coefficient through the batches' scalar slots.
        b = widths.pop()
        e_all = torch.cat([sub_batches[u] for u in mine], dim=1)
        sess_all = CNSession(graph, e_all, b, pred.order).build(pred.order, pred.weighted)
        ips, n_end = None, pred.n
        if ocn5:
            ip3 = pred.innerprod.detach().float().reshape(-1)[:1].repeat(3).contiguous()
            s_local = sess_all.stats(5, fill, ip3, 0)[:, 1].detach().clone()        # s of every owned sub-batch
            s_all = exchange_batch_scalars(s_local, mine, U)
            ips, n_end = replay_running_mean(s_all, pred.innerprod, pred.n)
            sess_all.set_batch_ip(ips[torch.as_tensor(mine, device=dev)])
        xcn1, xcn2, xcn3, xij, _ = pred.cn_stage(h, graph, e_all, fill, sess_all, per_batch_ip=ocn5)
        out = pred._head(xcn1, xcn2, xcn3, xij).reshape(len(mine), b)
        sg = torch.as_tensor([signs[u] for u in mine], dtype=out.dtype, device=dev).unsqueeze(1)
        l = -(1.0 / total_len) * F.logsigmoid(sg * out).sum()
        l.backward()
        loss += l.detach()
    else:
        sess, s_local = {}, []
        for u in mine:  # phase 1: CN sets + column statistics of the owned sub-batches, their inner products
            sess[u] = CNSession(graph, sub_batches[u], None, pred.order).build(pred.order, pred.weighted)
            if ocn5:
                s_local.append(pred.batch_inner_product(sess[u], fill))
        if ocn5:  # phase 2: one scalar exchange, running mean replayed in sequence order
            s_all = exchange_batch_scalars(torch.stack(s_local) if s_local else torch.zeros(0, device=dev), mine, U)
            ips, n_end = replay_running_mean(s_all, pred.innerprod, pred.n)
        for u in mine:  # phase 3: weighted aggregation, heads, loss, backward
            e = sub_batches[u]
            if ocn5:
                xcn1, xcn2, xcn3, xij, _ = pred.cn_stage(h, graph, e, fill, sess[u], ip=ips[u:u + 1])
            else:
                xcn1, xcn2, xcn3, xij, _ = pred.cn_stage(h, graph, e, fill, sess[u])
            out = pred._head(xcn1, xcn2, xcn3, xij)
            l = -(1.0 / total_len) * F.logsigmoid(signs[u] * out).sum()
            l.backward()
            loss += l.detach()
    if ocn5:
        with torch.no_grad():
            pred.innerprod.copy_(ips[-1:].to(pred.innerprod.dtype))
        pred.n = n_end
    allreduce_gradients(list(pred.parameters()) + [h])  # phase 4
    _all_reduce_sum(loss)
    return loss
