"""Import shims: the reference's own ``model.py`` / ``utils.py`` / driver text on libocn_b200.

The reference (qingpingmo/OCN) has no plugin interface; its boundary is the API of three third-party packages
(SURVEY.md 8b).  ``install()`` registers CUDA stand-ins for exactly the subset it touches:

    import ocn_b200.shim as shim
    shim.install()                 # before `import model` / `import utils`
    import model, utils            # the reference's files, unmodified
    shim.accelerate(utils, model)  # optional: utils.adjoverlap -> ocn_rows_intersect_* / ocn_rows_difference_*

* ``torch_sparse``            ``SparseTensor`` (CSR on the device), ``matmul.spmm_add/mean/max``, ``masked_select_nnz``
* ``pygho``                   ``SparseTensor`` with lazy row selections / products, ``backend.Spspmm.spsphadamard`` /
                              ``spspmm`` (``get_cn1_cn2``'s text runs the fused CN kernels), ``backend.Spmm.spmm``
* ``torch_geometric.nn``      ``GCNConv``

Every compute method runs on CUDA tensors through ``include/ocn_b200.h`` (plus torch CUDA ops for bookkeeping) and
raises ``OcnError`` on CPU tensors.  What the shim cannot reach is torch's own sparse arithmetic on plain tensors
(``spadj @ spadj`` on torch COO matrices, NeighborOverlap_large.py:74): that line is replaced by
``ocn_b200.shim.a2(adj)``.
"""
from __future__ import annotations

import importlib
import sys
from typing import Optional

import torch
from torch import Tensor

_NAMES = {
    "torch_sparse": "ocn_b200.shim.torch_sparse",
    "torch_sparse.tensor": "ocn_b200.shim.torch_sparse.tensor",
    "torch_sparse.matmul": "ocn_b200.shim.torch_sparse.matmul",
    "pygho": "ocn_b200.shim.pygho",
    "pygho.backend": "ocn_b200.shim.pygho.backend",
    "pygho.backend.Spspmm": "ocn_b200.shim.pygho.backend.Spspmm",
    "pygho.backend.Spmm": "ocn_b200.shim.pygho.backend.Spmm",
    "torch_geometric": "ocn_b200.shim.torch_geometric",
    "torch_geometric.nn": "ocn_b200.shim.torch_geometric.nn",
}
_saved = {}


def install(force: bool = False) -> None:
    """Make ``import torch_sparse`` / ``pygho`` / ``torch_geometric`` resolve to the CUDA stand-ins.  A real package that
    is already imported is left alone unless ``force`` (the stand-ins cover the reference's subset only)."""
    for public, private in _NAMES.items():
        if public in sys.modules and not force and not getattr(sys.modules[public], "__name__", "").startswith("ocn_b200."):
            continue
        _saved.setdefault(public, sys.modules.get(public))
        sys.modules[public] = importlib.import_module(private)


def uninstall() -> None:
    for public, old in _saved.items():
        if old is None:
            sys.modules.pop(public, None)
        else:
            sys.modules[public] = old
    _saved.clear()


def adjoverlap(adj1, adj2, tarei: Tensor, filled1: bool = False, calresadj: bool = False, cnsampledeg: int = -1,
               ressampledeg: int = -1, _sample=None):
    """``utils.adjoverlap`` (utils.py:248-285) on shim ``SparseTensor``s: per-link row intersection (and, with
    ``calresadj``, the two residual sets) straight on the CSR of the two matrices -- no row gather, no packed keys, no
    searchsorted.  The random samplers stay the reference's own ``sparsesample_reweight`` (passed in by ``accelerate``)."""
    from .. import cn as _cn
    from .torch_sparse.tensor import SparseTensor

    def wrap(rows):
        return SparseTensor._from_csr(rows.rowptr, rows.col, rows.value, (rows.shape[0], rows.shape[1]))

    if adj1.sizes()[1] != adj2.sizes()[1]:
        raise AssertionError("adj1.sizes() == adj2.sizes()")     # utils.py:165 (after the row gather: [B, N] both)
    out = _cn.adjoverlap(adj1.graph(), adj2.graph(), tarei.to(adj1.device()), calresadj=calresadj)
    if not calresadj:
        ov = wrap(out)
        return _sample(ov, cnsampledeg) if cnsampledeg > 0 else ov
    ov, r1, r2 = (wrap(o) for o in out)
    if cnsampledeg > 0:
        ov = _sample(ov, cnsampledeg)
    if ressampledeg > 0:
        r1, r2 = _sample(r1, ressampledeg), _sample(r2, ressampledeg)
    return ov, r1, r2


def accelerate(utils_module, *other_modules) -> None:
    """Point ``utils.adjoverlap`` -- and the copies ``from utils import adjoverlap`` left in ``model`` / a driver --
    at the fused version above."""
    sample = getattr(utils_module, "sparsesample_reweight", None)

    def fused(adj1, adj2, tarei, filled1=False, calresadj=False, cnsampledeg=-1, ressampledeg=-1):
        return adjoverlap(adj1, adj2, tarei, filled1, calresadj, cnsampledeg, ressampledeg, _sample=sample)

    fused.__doc__ = adjoverlap.__doc__
    fused.__wrapped__ = getattr(utils_module, "adjoverlap", None)
    for m in (utils_module,) + tuple(other_modules):
        if hasattr(m, "adjoverlap"):
            m.adjoverlap = fused


def a2(adj, fold: int = 0, has_value: bool = False):
    """``SparseTensor.from_torch_sparse_coo_tensor(spadj @ spadj, False)`` (NeighborOverlap_large.py:74) /
    ``sparse_tensor_multiply(spadj, block)`` (:71, ``fold=block``) in one call on ``ocn_spgemm_a2_*``."""
    from ..sparse_ops import spgemm_a2
    from .torch_sparse.tensor import SparseTensor
    g = spgemm_a2(adj.graph(), fold=fold, with_value=has_value or fold > 0)
    return SparseTensor._from_csr(g.rowptr, g.col, g.value, (g.n, g.n_cols))
