"""Multi-GPU plumbing: the graph and features are replicated on every rank, the
reference's own batch sequence (PermIterator order, utils.py:8-36) is dealt round-robin -- batch t goes
to rank t mod world -- and every rank runs the full fused path on its batches.  A batch is never split
(its links are coupled through the column statistics), so scores are identical for any world size.
The only collective is the gather of fp32 scores."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def deal_batches(num_links: int, batch_size: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) link ranges of the batches owned by ``rank``."""
    nb = (num_links + batch_size - 1) // batch_size
    return [(b * batch_size, min(num_links, (b + 1) * batch_size)) for b in range(rank, nb, world)]


def gather_scores(local_scores: Sequence[torch.Tensor], owned: Sequence[Tuple[int, int]], num_links: int) -> torch.Tensor:
    """All ranks contribute the scores of their batches; every rank receives the full [num_links] vector
    in link order (fp32, 4 B per link over NCCL/NVLink, or gloo on CPU)."""
    dev = local_scores[0].device if len(local_scores) else torch.device("cpu")
    full = torch.zeros(num_links, dtype=torch.float32, device=dev)
    for sc, (s, e) in zip(local_scores, owned):
        full[s:e] = sc.reshape(-1).float()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)  # disjoint supports: a sum is a gather in link order
    return full
