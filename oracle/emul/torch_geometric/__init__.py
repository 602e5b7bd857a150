"""Stand-in so that `from torch_geometric.nn import GCNConv` in the reference's model.py resolves."""
