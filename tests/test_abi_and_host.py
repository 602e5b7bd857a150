"""CPU-side checks: the C-ABI library loads and exports every symbol include/ocn_b200.h declares,
argument validation works without a GPU, the product never falls back to the CPU, host-side logic."""
import ctypes
import os

import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import _lib, synth
from ocn_b200.cn import waves


def test_library_exports_every_header_symbol():
    L = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ocn_b200.h but not exported"
        assert s in _lib._SIGS, f"{s} has no ctypes signature"
    assert L.ocn_abi_version() == 3


def test_argument_errors_do_not_need_a_gpu():
    L = _lib.lib()
    assert L.ocn_cn_plan_bytes(-1) == 0 and L.ocn_cn_plan_bytes(1000) > 1000 * 8
    assert L.ocn_cn_record_bytes() == 8 and L.ocn_cn_colstat_bytes(10) == 320
    rc = L.ocn_cn_build(None, None, 0, None, None, 0, 0, 2, 1, None, None, None, 0, None, 0, None, None, 0, None, None)
    assert rc == -1 and b"null pointer" in L.ocn_last_error()
    rc = L.ocn_spmm_csr(None, None, None, 4, None, 32, 0, None, None)
    assert rc == -1
    rc = L.ocn_spgemm_a2_symbolic(None, None, 4, 8, 0, None, None, None)
    assert rc == -1
    # the entry points either side of the path (graph build / mask, fused head, metrics)
    assert L.ocn_graph_build_bytes(1000, 1) > 2 * 1000 * (8 + 8 + 8 + 4) and L.ocn_graph_mask_bytes(100, 10) > 0
    assert L.ocn_graph_build_count(None, None, None, 5, 10, 1, None, 0, None, None, None) == -1
    assert L.ocn_graph_mask_count(None, None, None, 10, 0, None, None, 0, 1, None, None, 0, None, None, None) == -1
    assert L.ocn_cn_head_params(32, 32, 1, 0, 2) == 2 * (32 * 32 * 3 + 32 * 3) + (32 * 32 * 2 + 64) + (32 * 32 + 32 + 32 + 1)
    assert L.ocn_cn_head_params(256, 256, 1, 0, 2) == -1 and L.ocn_cn_head_params(64, 64, 1, 4, 3) == -1  # beyond 200 KB
    assert L.ocn_cn_head(None, None, None, None, 5, 32, 32, 1, 0, None, 0, None, None, None) == -1
    assert L.ocn_hits_bytes(1000) >= 4000 and L.ocn_mrr(None, None, 3, 4, None, None) == -1
    assert L.ocn_rows_difference_count(None, None, 4, None, None, 4, None, None, 3, None, None) == -1


def test_no_cpu_fallback():
    g = synth.tiny_graph(30, 80, 1)
    G = ob.Graph(g.rowptr, g.col, g.n)  # CPU tensors
    e = g.query_edges(8, "neg")
    with pytest.raises(_lib.OcnError):
        ob.adjoverlap(G, G, e)
    with pytest.raises(_lib.OcnError):
        ob.CNSession(G, e)
    with pytest.raises(_lib.OcnError):
        ob.spmm_add(G, torch.zeros(30, 4))
    with pytest.raises(_lib.OcnError):
        ob.pure_conv(torch.zeros(30, 4), G, "gcn")
    with pytest.raises(_lib.OcnError):
        G.masked(e)
    with pytest.raises(_lib.OcnError):
        ob.metrics.mrr_list(torch.zeros(4), torch.zeros(4, 10))
    with pytest.raises(_lib.OcnError):
        ob.metrics.hits_at_k(torch.zeros(4), torch.zeros(40), 3)
    with pytest.raises(_lib.OcnError):
        ob.adjoverlap(G, G, e, calresadj=True)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "ocn_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_graph_and_synth_host_logic():
    g = synth.make_graph("cora")
    assert g.n == 2708 and abs(g.nnz - 10556) < 400
    keep = g.raw_src != g.raw_dst  # the generator drops self loops, from_edge_index (like the reference) keeps them
    G = ob.Graph.from_edge_index(torch.stack((g.raw_src[keep], g.raw_dst[keep])), g.n)
    assert torch.equal(G.rowptr, g.rowptr) and torch.equal(G.col, g.col)
    deg = G.degree()
    assert int(deg.sum()) == g.nnz and G.sizes() == (2708, 2708)
    # symmetric, sorted, deduplicated
    r = G.row()
    key = r * g.n + G.col.long()
    assert bool((key[1:] > key[:-1]).all())
    rev = torch.sort(G.col.long() * g.n + r).values
    assert torch.equal(rev, key)
    # generators are seed-deterministic and device independent by construction (integer hash)
    g2 = synth.make_graph("cora")
    assert torch.equal(g2.col, g.col)
    e = g.query_edges(3000, "stream")
    assert e.shape == (2, 3000) and bool((e[0, :1000] == e[0, 0]).all()) and bool((deg[e[0]] > 0).all())


def test_waves_cover_the_stream():
    w = waves(10 * 2048 + 5, 2048, 1000, budget_bytes=3 * 32 * 1000)
    assert w[0] == (0, 3 * 2048) and w[-1][1] == 10 * 2048 + 5
    assert all(a[1] == b[0] for a, b in zip(w, w[1:]))


def test_predictor_state_dict_keys_match_reference_layout():
    p = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0)
    keys = set(p.state_dict().keys())
    for k in ("beta", "alpha", "innerprod", "dropadj.ratio", "xcn1lin.0.weight", "xcn2lin.7.bias", "xcn4lin.3.weight",
              "xijlin.0.weight", "xijlin.4.bias", "lin.0.weight", "lin.8.weight", "xcnlin.7.weight"):
        assert k in keys, k
    p6 = ob.CNLinkPredictor3hopCNs(32, 32, 1, 3, 0.0)
    assert "xcn3lin.7.weight" in p6.state_dict() and "xcn4lin.0.weight" not in p6.state_dict()


def test_head_parameter_packing_layout():
    """Host logic of the fused head (ocn_b200/head.py): the flat parameter buffer read back in the order csrc/head.cu
    documents reproduces the module's own forward (CPU, no kernel involved)."""
    import torch.nn.functional as F
    from ocn_b200 import head
    from ocn_b200.predictor import CNLinkPredictor3hopCNs

    for ln, tailact, two in ((False, False, False), (True, False, True), (True, True, False)):
        torch.manual_seed(3)
        I, H, O = 32, 64, 2
        pred = CNLinkPredictor3hopCNs(I, H, O, 3, 0.0, ln=ln, tailact=tailact, twolayerlin=two).eval()
        st = {"three": True}
        pieces = head._pack_seq(pred.xcn1lin) + head._pack_seq(pred.xcn2lin) + head._pack_seq(pred.xcn3lin) \
            + head._pack_seq(pred.xijlin) + head._pack_seq(pred.lin, last_plain=True)
        flat = torch.cat([p.detach().reshape(-1) for p in pieces])
        assert head._flags(pred) == (1 if ln else 0) | (2 if tailact else 0) | (4 if two else 0)
        off = 0

        def take(n):
            nonlocal off
            v = flat[off:off + n]
            off += n
            return v

        def lin_t(x, i, o):  # weights are stored transposed: [in, out]
            w, b = take(i * o).view(i, o), take(o)
            return x @ w + b

        def norm(x):
            g, b = take(H), take(H)
            return F.layer_norm(x, (H,), g, b, 1e-5)

        xs = [torch.randn(7, I) for _ in range(4)]
        with torch.no_grad():
            alpha = torch.sigmoid(pred.alpha).cumprod(-1)
            mix = [float(alpha[0]), float(alpha[1]), float(alpha[2]), float(pred.beta)]
            z = 0
            for br in range(3):
                a = torch.relu(lin_t(xs[br], I, H))
                a = lin_t(a, H, H)
                if ln:
                    a = norm(a)
                a = lin_t(torch.relu(a), H, H)
                z = z + mix[br] * a
            a = lin_t(xs[3], I, H)
            if ln:
                a = norm(a)
            a = torch.relu(a)
            if not tailact:
                a = lin_t(a, H, H)
            z = z + mix[3] * a
            a = lin_t(z, H, H)
            if ln:
                a = norm(a)
            a = torch.relu(a)
            if two:
                a = lin_t(a, H, H)
                if ln:
                    a = norm(a)
                a = torch.relu(a)
            wo, bo = take(O * H).view(O, H), take(O)  # the last Linear keeps its [out, in] layout
            out = a @ wo.t() + bo
            assert off == flat.numel()
            pred.fuse_head = False
            ref = pred._head(xs[0], xs[1], xs[2], xs[3])
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_header_is_plain_c():
    """include/ocn_b200.h is what a cgo / ctypes / JNI maintainer binds: it must compile as C99 on its own."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write('#include "include/ocn_b200.h"\nint main(void) { return ocn_abi_version() == OCN_ABI_VERSION ? 0 : 1; }\n')
        path = f.name
    try:
        res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", root, path], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    finally:
        os.unlink(path)


def test_cost_dealing_of_stream_slices():
    """ocn_b200.dist.predicted_walk_cost / deal_by_cost (host logic of the multi-GPU bench and evaluation loop):
    the predicted cost is the exact index-entry count of a slice, the dealing is a partition with equal counts that
    is identical on every rank and never worse balanced than contiguous chunks."""
    from ocn_b200 import dist as obdist
    g = synth.tiny_graph(300, 2400, 7)
    rp, col = g.rowptr, g.col
    deg = (rp[1:] - rp[:-1]).numpy()
    slice_links, batch, world, per_rank = 64, 16, 4, 3
    T = slice_links * world * per_rank
    srcs = torch.randint(0, g.n, (T // 10 + 1,), generator=torch.Generator().manual_seed(3))
    src = srcs.repeat_interleave(10)[:T]                      # runs of 10 links, cut again at the slice boundaries
    cost = obdist.predicted_walk_cost(rp, col, src, slice_links, batch).tolist()
    # brute force: runs = maximal pieces of one source inside one slice (= one session); entries = sum over runs, over k in N(src), of d(k)
    ref = [0] * (T // slice_links)
    cn, rpn, s = col.numpy(), rp.numpy(), src.numpy()
    for t in range(T):
        if t % slice_links == 0 or s[t] != s[t - 1]:
            ref[t // slice_links] += int(sum(deg[k] for k in cn[rpn[s[t]]:rpn[s[t] + 1]]))
    assert cost == ref
    owned = obdist.deal_by_cost(cost, world, per_rank)
    assert sorted(i for o in owned for i in o) == list(range(world * per_rank))
    assert all(len(o) == per_rank and o == sorted(o) for o in owned)
    assert owned == obdist.deal_by_cost(list(cost), world, per_rank)   # deterministic: same on every rank
    load = lambda ids: sum(cost[i] for i in ids)
    contiguous = max(load(range(r * per_rank, (r + 1) * per_rank)) for r in range(world))
    assert max(load(o) for o in owned) <= contiguous
    # one very heavy slice: it ends up with the lightest companions
    skew = [100, 1, 2, 3, 4, 5, 6, 7]
    o = obdist.deal_by_cost(skew, 2, 4)
    heavy = o[0] if 0 in o[0] else o[1]
    assert sorted(heavy) == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        obdist.deal_by_cost([1, 2, 3], 2, 2)


def test_bench_algorithmic_bytes_follow_the_survey_formula():
    """bench.algorithmic_bytes: the SURVEY.md 8(d) per-link figure (index terms, the x[i] / x[j] gathers and the
    outputs; the gathers over the CN sets themselves are left out, which only lowers the figure) and the part
    of its order-3 index term that runs through rows of >= hub_degree columns (numerator of roofline.achieved),
    against a per-link python count."""
    import bench
    g = synth.tiny_graph(80, 400, 2)
    G = ob.Graph(g.rowptr, g.col, g.n)
    e = g.query_edges(64, "mixed")
    rp, col = g.rowptr.tolist(), g.col.tolist()
    d = lambda v: rp[v + 1] - rp[v]
    N = lambda v: col[rp[v]:rp[v + 1]]
    feat, hub_d = 8, 6
    for order in (1, 2, 3):
        _, _, survey, hub = bench.algorithmic_bytes(G, e, order, feat, 16, hub_d=hub_d)
        want, want_hub = 0, 0
        for i, j in e.t().tolist():
            want += 32 + 4 * (d(i) + d(j)) + 4 * feat * (order + 1 + 2)
            if order >= 2:
                want += 8 * d(j) + 4 * sum(d(m) for m in N(j))
            if order >= 3:
                want += 8 * d(i) + 4 * sum(d(k) for k in N(i))
                want_hub += sum(8 + 4 * d(m) for m in N(j) if d(m) >= hub_d)
        assert survey == want
        assert hub == want_hub
