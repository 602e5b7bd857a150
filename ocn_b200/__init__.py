"""ocn_b200 -- B200-native common-neighbour hot path of OCN.

Importing the package does not load the CUDA library; the first op does, and raises if
``libocn_b200.so`` is missing (there is no CPU fallback).
"""
from . import synth  # noqa: F401
from .graph import Graph  # noqa: F401
from .cn import (CNSession, SparseRows, adjoverlap, cn_aggregate_eval, get_cn, get_cn1_cn2,  # noqa: F401
                 reserve_stream_pool)
from .sparse_ops import (drop_adj, gcn_norm, gcnconv_propagate, pure_conv, pure_conv3_gcn, sparse_tensor_multiply,  # noqa: F401
                         spgemm_a2, spmm, spmm_add, spmm_max, spmm_mean)
from . import ops  # noqa: F401  (registers torch.ops.ocn.*)
from . import dist, metrics  # noqa: F401
from .predictor import (CNLinkPredictor3hopCNs, CNLinkPredictorbaselearn, CNLinkPredictorOringin,  # noqa: F401
                        predictor_dict)
from .completion import (IncompleteCN1Predictor, IncompleteCN1Predictorhighorder,  # noqa: F401
                         IncompleteCN1PredictorSaveMemory)
from .sample import sparsesample_reweight  # noqa: F401

predictor_dict["cn2"] = IncompleteCN1Predictor
predictor_dict["cn3"] = IncompleteCN1Predictorhighorder
predictor_dict["cn4"] = IncompleteCN1PredictorSaveMemory

__all__ = [
    "Graph", "CNSession", "SparseRows", "adjoverlap", "cn_aggregate_eval", "get_cn", "get_cn1_cn2",
    "drop_adj", "gcn_norm", "gcnconv_propagate", "pure_conv", "pure_conv3_gcn", "sparse_tensor_multiply", "spgemm_a2",
    "spmm", "spmm_add", "spmm_max", "spmm_mean", "CNLinkPredictorOringin", "CNLinkPredictor3hopCNs",
    "CNLinkPredictorbaselearn", "IncompleteCN1Predictor", "IncompleteCN1PredictorSaveMemory", "IncompleteCN1Predictorhighorder", "sparsesample_reweight", "predictor_dict", "synth", "reserve_stream_pool", "metrics", "dist",
]
