"""Seeded synthetic graphs of each BASELINE.json config's shape.

The generators are device independent: every random number comes from a counter-based
integer hash (splitmix64 on the element index) evaluated with wrapping int64 torch ops, so
the same (shape, seed) yields the same graph on the CPU (oracle side) and on a B200.
The only floating-point step, the Chung-Lu inverse CDF, is always evaluated on the host in
float64 and looked up with an exact ``searchsorted``.

Shapes follow SURVEY.md §8(d) / BASELINE.md §4 (the reference's datasets are loaded by
``ogbdataset.py:29-71`` and need the network; only their node/edge shape is reproduced).
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import torch

_M1 = -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687  # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685  # 0x94D049BB133111EB


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """Logical shift right on int64 (torch's >> is arithmetic)."""
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser, bit-identical on every device."""
    z = x + _M1
    z = (z ^ _lsr(z, 30)) * _M2
    z = (z ^ _lsr(z, 27)) * _M3
    return z ^ _lsr(z, 31)


def hash_u01(n: int, seed: int, stream: int, device, start: int = 0) -> torch.Tensor:
    """n uniform doubles in [0,1), element t = f(seed, stream, start + t)."""
    idx = torch.arange(start, start + n, dtype=torch.int64, device=device)
    z = mix64(idx + mix64(torch.tensor(seed * 1000003 + stream, dtype=torch.int64, device=device)))
    return _lsr(z, 11).to(torch.float64) * (2.0 ** -53)


def hash_randint(n: int, high: int, seed: int, stream: int, device, start: int = 0) -> torch.Tensor:
    """n integers uniform in [0, high): elements start .. start + n - 1 of the sequence."""
    u = hash_u01(n, seed, stream, device, start)
    return (u * high).to(torch.int64).clamp_(max=high - 1)


def hash_normal(shape, seed: int, stream: int, device) -> torch.Tensor:
    """fp32 N(0,1) via Box-Muller on hashed uniforms (device independent up to libm ulps)."""
    n = 1
    for s in shape:
        n *= int(s)
    u1 = hash_u01(n, seed, stream, device).clamp_(min=2.0 ** -53)
    u2 = hash_u01(n, seed, stream + 7919, device)
    z = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)
    return z.to(torch.float32).reshape(shape)


def symmetric_csr(src: torch.Tensor, dst: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Undirected, deduplicated, self-loop-free CSR (rowptr int64[n+1], col int32[nnz] sorted).

    Mirrors ``SparseTensor.from_edge_index(...).to_symmetric()`` + coalesce as the reference
    builds ``data.adj_t`` (ogbdataset.py:44-45, NeighborOverlap_large.py:59-63).
    """
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = torch.cat((src * n + dst, dst * n + src))
    key = torch.unique(key)  # sorted
    row = torch.div(key, n, rounding_mode="floor")
    col = (key - row * n).to(torch.int32)
    counts = torch.bincount(row, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr, col


def chung_lu_edges(n: int, m: int, gamma: float, max_deg_frac: float, seed: int, device,
                   permute: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """m directed endpoint pairs with P(u) ∝ w_u, w_v = (v + i0)^(-1/(gamma-1)).

    ``max_deg_frac`` caps the expected share of endpoints on the heaviest node (sets i0).
    """
    alpha = 1.0 / (gamma - 1.0)
    # choose i0 so that w_0 / sum(w) ~= max_deg_frac (host, float64, bisection on a fixed grid)
    ar = torch.arange(n, dtype=torch.float64)
    lo, hi = 1e-3, float(n)
    for _ in range(60):
        mid = (lo * hi) ** 0.5
        w = (ar + mid) ** (-alpha)
        if (w[0] / w.sum()).item() > max_deg_frac:
            lo = mid
        else:
            hi = mid
    w = (ar + hi) ** (-alpha)
    cdf = torch.cumsum(w, 0)
    cdf = (cdf / cdf[-1]).to(device)
    u = hash_u01(m, seed, 1, device)
    v = hash_u01(m, seed, 2, device)
    src = torch.searchsorted(cdf, u).clamp_(max=n - 1)
    dst = torch.searchsorted(cdf, v).clamp_(max=n - 1)
    if permute:
        order = torch.argsort(mix64(torch.arange(n, dtype=torch.int64, device=device) + seed * 7 + 13), stable=True)
        src, dst = order[src], order[dst]
    return src, dst


def erdos_renyi_edges(n: int, p: float, seed: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every unordered pair u<v kept with probability p."""
    iu = torch.triu_indices(n, n, offset=1, device=device)
    keep = hash_u01(iu.shape[1], seed, 3, device) < p
    return iu[0][keep], iu[1][keep]


@dataclasses.dataclass
class SynthGraph:
    name: str
    n: int
    rowptr: torch.Tensor  # int64 [n+1]
    col: torch.Tensor     # int32 [nnz]
    f_in: int
    hidden: int
    batch: int
    predictor: str        # "cn5" | "cn7"
    conv: str
    mplayers: int
    order: int = 2
    weighted: bool = False
    seed: int = 0
    raw_src: Optional[torch.Tensor] = None  # the directed/raw positive edges the graph was built from
    raw_dst: Optional[torch.Tensor] = None

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def features(self, width: Optional[int] = None, device=None) -> torch.Tensor:
        device = device if device is not None else self.col.device
        return hash_normal((self.n, width or self.hidden), self.seed, 101, device)

    def stream_links(self, start: int, num: int, device=None) -> torch.Tensor:
        """Links start .. start + num - 1 of the 'stream' kind of ``query_edges`` without generating the ones before
        them (element t of the stream depends on t only)."""
        device = device if device is not None else self.col.device
        rp = self.rowptr.to(device)
        deg = rp[1:] - rp[:-1]
        cand = torch.nonzero(deg > 0).flatten()
        s0, s1 = start // 1000, (start + num + 999) // 1000
        srcs = cand[hash_randint(s1 - s0, cand.numel(), self.seed, 205, device, start=s0)]
        src = srcs.repeat_interleave(1000)[start - s0 * 1000:start - s0 * 1000 + num]
        dst = hash_randint(num, self.n, self.seed, 204, device, start=start)
        return torch.stack((src, dst))

    def query_edges(self, num: int, kind: str = "mixed", device=None) -> torch.Tensor:
        """[2,num] int64 target links: 'pos' = edges of the graph, 'neg' = uniform pairs,
        'mixed' = first half pos, second half neg, 'stream' = citation2 test stream shape
        (each positive source repeated against uniform random destinations, 1+1000 per source;
        NeighborOverlapCitation2.py:241-254)."""
        device = device if device is not None else self.col.device
        rp, col = self.rowptr.to(device), self.col.to(device)
        if kind == "pos":
            e = hash_randint(num, self.nnz, self.seed, 201, device)
            src = torch.searchsorted(rp, e, right=True) - 1
            return torch.stack((src, col[e].to(torch.int64)))
        if kind == "neg":
            return torch.stack((hash_randint(num, self.n, self.seed, 202, device),
                                hash_randint(num, self.n, self.seed, 203, device)))
        if kind == "mixed":
            h = num // 2
            return torch.cat((self.query_edges(h, "pos", device), self.query_edges(num - h, "neg", device)), dim=1)
        if kind == "stream":
            # negatives of the citation2 test split: every source repeated against 1000 uniform
            # destinations (source.view(-1,1).repeat(1,1000), NeighborOverlapCitation2.py:248-252);
            # sources are uniform over nodes that have at least one neighbour.
            deg = rp[1:] - rp[:-1]
            cand = torch.nonzero(deg > 0).flatten()
            nsrc = (num + 999) // 1000
            srcs = cand[hash_randint(nsrc, cand.numel(), self.seed, 205, device)]
            src = srcs.repeat_interleave(1000)[:num]
            dst = hash_randint(num, self.n, self.seed, 204, device)
            return torch.stack((src, dst))
        raise ValueError(kind)


# name -> (n, directed edge draws, generator, params..., f_in, hidden, batch, predictor, conv, mplayers, order, weighted)
CONFIGS = {
    # BASELINE.json configs[0]: README.md:24-30 (Cora, cn5, puregcn, hiddim 256, mplayers 1)
    "cora": dict(n=2708, m=5278, gen="cl", gamma=2.5, maxfrac=170 / 10556, f_in=1433, hidden=256, batch=1152,
                 predictor="cn5", conv="puregcn", mplayers=1, order=2, weighted=False, seed=0),
    # configs[1]: Pubmed cn7
    "pubmed": dict(n=19717, m=44324, gen="cl", gamma=2.5, maxfrac=171 / 88648, f_in=500, hidden=256, batch=2048,
                   predictor="cn7", conv="puregcn", mplayers=1, order=2, weighted=False, seed=1),
    # configs[2]: collab cn5 gin, valedges as input
    "collab": dict(n=235868, m=1285465 + 60084, gen="cl", gamma=2.6, maxfrac=671 / 2.4e6, f_in=128, hidden=256,
                   batch=65536, predictor="cn5", conv="gin", mplayers=1, order=2, weighted=False, seed=2),
    # configs[3]: ddi cn7 dense
    "ddi": dict(n=4267, m=0, gen="er", p=0.1466, f_in=0, hidden=64, batch=32768,
                predictor="cn7", conv="puregcn", mplayers=3, order=2, weighted=False, seed=3),
    # configs[4]: citation2 cn5, README.md:53 (hiddim 32, mplayers 5, testbs 2048)
    "citation2": dict(n=2927963, m=30561187, gen="cl", gamma=2.3, maxfrac=10000 / 6.1e7, f_in=128, hidden=32,
                      batch=2048, predictor="cn5", conv="gcn", mplayers=5, order=3, weighted=True, seed=4),
}


def make_graph(name: str, device="cpu", scale: float = 1.0) -> SynthGraph:
    """Build the named config's graph; ``scale`` < 1 shrinks node and edge counts together
    (used by CPU tests that want the same degree law at a size the oracle finishes quickly)."""
    c = CONFIGS[name]
    n = max(16, int(round(c["n"] * scale)))
    seed = c["seed"]
    if c["gen"] == "er":
        src, dst = erdos_renyi_edges(n, c["p"], seed, device)
    else:
        m = max(8, int(round(c["m"] * scale)))
        src, dst = chung_lu_edges(n, m, c["gamma"], min(0.2, c["maxfrac"] / max(scale, 1e-9) ** 0.5), seed, device)
    rowptr, col = symmetric_csr(src, dst, n)
    return SynthGraph(name=name, n=n, rowptr=rowptr, col=col, f_in=c["f_in"], hidden=c["hidden"], batch=c["batch"],
                      predictor=c["predictor"], conv=c["conv"], mplayers=c["mplayers"], order=c["order"],
                      weighted=c["weighted"], seed=seed, raw_src=src, raw_dst=dst)


def tiny_graph(n: int, m: int, seed: int, device="cpu") -> SynthGraph:
    """Small uniform random graph for property tests."""
    src = hash_randint(m, n, seed, 11, device)
    dst = hash_randint(m, n, seed, 12, device)
    rowptr, col = symmetric_csr(src, dst, n)
    return SynthGraph(name=f"tiny{n}", n=n, rowptr=rowptr, col=col, f_in=8, hidden=8, batch=16, predictor="cn5",
                      conv="puregcn", mplayers=1, seed=seed, raw_src=src, raw_dst=dst)
