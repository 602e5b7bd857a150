"""Stand-in for the one torch_geometric symbol on the path: ``torch_geometric.nn.GCNConv`` (model.py:1, 58-71)."""
from . import nn  # noqa: F401

__version__ = "2.6.1+ocn_b200"
