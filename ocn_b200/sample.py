"""``sparsesample_reweight`` (utils.py:109-143) on the GPU: rows with more than ``deg`` entries are replaced by ``deg``
draws WITH replacement (each worth ``rowcount / deg``; repeated draws of one column add up), the other rows keep
their entries with value 1.  The completion predictors cn2-cn4 call it on the residual sets of every batch
(``ressampledeg`` = 8 in training / 128 in evaluation, model.py:868-869, 910).

The random numbers are torch's (``torch.rand((rows, deg), device=...)``, as in the reference), so a seeded run draws
what the reference would draw on the same device; ``rand_fn(shape, device)`` replaces the source (the parity tests
feed the reference's recorded draws).  Torch CUDA ops around one sort: this is bookkeeping over O(rows * deg) entries,
not a hot kernel, and nothing runs on the CPU.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .cn import SparseRows
from .graph import _require_cuda


def sparsesample_reweight(adj: SparseRows, deg: int, rand_fn: Optional[Callable] = None) -> SparseRows:
    _require_cuda(adj.col)
    dev = adj.col.device
    M, N = adj.shape
    rowptr, col = adj.rowptr, adj.col
    rowcount = rowptr[1:] - rowptr[:-1]
    mask = rowcount > deg
    rc = rowcount[mask]
    # (the draw happens even for zero rows, as in the reference: a replayed sequence of draws stays aligned)
    rand = rand_fn((rc.size(0), deg), dev) if rand_fn is not None else torch.rand((rc.size(0), deg), device=dev)
    if rc.numel() == 0:                       # nothing to sample: values become 1 (ones_like(nosamplerow), utils.py:139)
        return SparseRows(rowptr, col, torch.ones(col.numel(), dtype=torch.float32, device=dev), adj.shape)
    rand = rand.to(device=dev, dtype=torch.float32).mul(rc.to(torch.float32).reshape(-1, 1)).to(torch.long)
    rand.add_(rowptr[:-1][mask].reshape(-1, 1))
    samplecol = col[rand].flatten()
    rows = torch.arange(M, device=dev)
    samplerow = rows[mask].reshape(-1, 1).expand(-1, deg).flatten()
    samplevalue = (rc * (1 / deg)).to(torch.float32).reshape(-1, 1).expand(-1, deg).flatten()
    keep = torch.repeat_interleave(~mask, rowcount, output_size=int(col.numel()))       # entries of the unsampled rows
    row_all = torch.repeat_interleave(rows, rowcount, output_size=int(col.numel()))
    r = torch.cat((samplerow, row_all[keep]))
    c = torch.cat((samplecol, col[keep]))
    v = torch.cat((samplevalue, torch.ones(int(keep.sum()), dtype=torch.float32, device=dev)))
    key, inv = torch.unique(r * N + c, return_inverse=True)                              # .coalesce(): sorted, duplicates summed
    val = torch.zeros(key.numel(), dtype=torch.float32, device=dev).index_add_(0, inv, v)
    orow = torch.div(key, N, rounding_mode="floor")
    out_rowptr = torch.zeros(M + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(orow, minlength=M), 0, out=out_rowptr[1:])
    return SparseRows(out_rowptr, key - orow * N, val, adj.shape)
