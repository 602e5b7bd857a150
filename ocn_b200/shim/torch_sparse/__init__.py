"""Stand-in for ``torch_sparse`` 0.6.18 (the subset qingpingmo/OCN touches) on libocn_b200's CUDA kernels.
Activated by ``ocn_b200.shim.install()``; see ocn_b200/shim/__init__.py."""
from .tensor import SparseTensor, masked_select_nnz  # noqa: F401
from . import matmul as _matmul_module  # noqa: F401
from .matmul import matmul, spmm, spmm_add, spmm_max, spmm_mean, spmm_sum  # noqa: F401

__version__ = "0.6.18+ocn_b200"
