"""Device-side Hits@K / MRR against the oracle's restatement of the ogb 1.3.6 evaluator."""
import pytest
import torch

from ocn_b200 import metrics
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,K", [(1, 1), (7, 1000), (513, 37), (86, 1000)])
def test_mrr_matches_evaluator(B, K):
    g = torch.Generator().manual_seed(B * 1000 + K)
    pos = torch.randn(B, generator=g)
    neg = torch.randn(B, K, generator=g)
    neg[:, ::5] = pos.unsqueeze(1)[:, :1].expand(-1, neg[:, ::5].shape[1])  # ties: optimistic != pessimistic rank
    got = metrics.mrr_list(pos.to(DEV), neg.to(DEV)).cpu()
    assert torch.equal(got, R.mrr(pos, neg))


@pytest.mark.parametrize("P,M,K", [(100, 5, 20), (1000, 5000, 1), (1000, 5000, 50), (3, 100000, 100), (0, 10, 3)])
def test_hits_matches_evaluator(P, M, K):
    g = torch.Generator().manual_seed(P + M + K)
    pos = torch.round(torch.randn(P, generator=g) * 8) / 8  # coarse grid: ties with the threshold
    neg = torch.round(torch.randn(M, generator=g) * 8) / 8
    got = float(metrics.hits_at_k(pos.to(DEV), neg.to(DEV), K))
    assert abs(got - R.hits_at_k(pos, neg, K)) < 1e-7


def test_result_dictionary():
    g = torch.Generator().manual_seed(0)
    t = [torch.randn(n, generator=g) for n in (300, 200, 4000, 250, 4000)]
    res = metrics.evaluate_hits(*[v.to(DEV) for v in t], ks=(20, 50, 100))
    for K in (20, 50, 100):
        want = (R.hits_at_k(t[0], t[2], K), R.hits_at_k(t[1], t[2], K), R.hits_at_k(t[3], t[4], K))
        assert all(abs(a - b) < 1e-7 for a, b in zip(res[f"Hits@{K}"], want))
