"""``GCNConv`` as ``convdict`` configures it (model.py:58-71): ``gcn`` (symmetric normalisation with self loops),
``sage`` / ``gin`` / ``max`` (plain mean / sum / max aggregation of ``x W``), always followed by ``+ bias``.  The
neighbour aggregation runs in ``ocn_gcn_spmm`` / ``ocn_spmm_csr``; parameter names (``lin.weight``, ``bias``) and
initialisation (glorot weight, zero bias) follow PyG 2.6.1 so that a ``state_dict`` moves across."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from ....graph import Graph
from ....sparse_ops import gcnconv_propagate


class _Lin(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x: Tensor) -> Tensor:
        return x @ self.weight.t()


class GCNConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops=None, normalize: bool = True, bias: bool = True, aggr: str = "add", **kwargs):
        super().__init__()
        if improved:
            raise NotImplementedError("convdict never builds GCNConv(improved=True)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize = normalize
        self.add_self_loops = normalize if add_self_loops is None else add_self_loops
        self.aggr = {"add": "sum"}.get(aggr, aggr)
        self.cached = cached
        self.lin = _Lin(in_channels, out_channels)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, adj_t, edge_weight=None) -> Tensor:
        g = adj_t if isinstance(adj_t, Graph) else adj_t.graph()
        out = gcnconv_propagate(self.lin(x), g, self.normalize, self.add_self_loops, self.aggr)
        return out if self.bias is None else out + self.bias
