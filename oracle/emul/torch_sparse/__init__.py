"""Stand-in for torch_sparse 0.6.18 (subset). TEST INFRASTRUCTURE ONLY -- see ../README.md."""
from .tensor import SparseTensor, masked_select_nnz  # noqa: F401
from . import matmul  # noqa: F401
from .matmul import spmm_add, spmm_max, spmm_mean  # noqa: F401
