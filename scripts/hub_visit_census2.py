"""Second CPU census of `k_cn_hub_count`'s entry visits: what exact run signatures and per-run segments buy.

Same slice reconstruction as `scripts/hub_visit_census.py` (numpy / scipy only, no device, no product code).
For every column occurrence (m, l) of a shared row the script knows

    len(l)   entries of l's list,   k(l) = distinct runs in it,   hits = |runs(l) & active runs of m|,
    useful   entries of l in active runs of m,

and prices candidate walks in entry visits:

    today       whole list if the 64-bit folded signature meets the active set
    exact       whole list if the exact signature meets it (<= 128 runs)
    segments    only the entries of the active runs (useful), the list's run segments found by rank
    mixed(L)    lists of up to L entries whole (exact filter), longer lists by segments

Output: visit totals per scheme, split by list length and by active runs of the row.

    python scripts/hub_visit_census2.py [--slice 0]
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocn_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slice", type=int, default=0)
    ap.add_argument("--links", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--hub", type=int, default=0)
    a = ap.parse_args()
    t0 = time.time()
    g = synth.make_graph("citation2", device="cpu")
    n = g.n
    rowptr = g.rowptr.numpy().astype(np.int64)
    col = g.col.numpy().astype(np.int32)
    deg = np.diff(rowptr)
    T = a.links
    e = g.query_edges((a.slice + 1) * T, "stream", device="cpu").numpy()[:, a.slice * T:(a.slice + 1) * T]
    src, dst = e[0], e[1]
    hub = a.hub if a.hub > 0 else max(32, -(-n // T))
    A = sp.csr_matrix((np.ones(col.size, np.float32), col, rowptr), shape=(n, n))
    t = np.arange(T)
    first = (t == 0) | (np.r_[-1, src[:-1]] != src)   # runs cross batch boundaries (round 2; round 1 cut them: t % batch == 0)
    run_of_link = np.cumsum(first) - 1
    run_src = src[first]
    R = run_src.size
    assert R <= 128
    pos_run = np.repeat(np.arange(R), deg[run_src])
    pos_k = np.concatenate([col[rowptr[s]:rowptr[s + 1]] for s in run_src])
    P = pos_k.size
    K = sp.csr_matrix((np.ones(P, np.float32), (pos_run, pos_k)), shape=(R, n))
    E = (K @ A).tocsc()   # E[r, l] = entries of l in run r
    # per key l: list length, number of distinct runs, exact signature (2 x uint64), folded 64-bit signature
    llen = np.asarray(E.sum(0)).ravel().astype(np.int64)
    ecoo = E.tocoo()
    sig = np.zeros((n, 2), np.uint64)
    np.bitwise_or.at(sig[:, 0], ecoo.col[ecoo.row < 64], (np.uint64(1) << ecoo.row[ecoo.row < 64].astype(np.uint64)))
    hi = ecoo.row >= 64
    np.bitwise_or.at(sig[:, 1], ecoo.col[hi], (np.uint64(1) << (ecoo.row[hi] - 64).astype(np.uint64)))
    fold = sig[:, 0] | sig[:, 1]
    kruns = np.bitwise_count(sig[:, 0]).astype(np.int64) + np.bitwise_count(sig[:, 1])
    print(f"slice {a.slice}: runs {R} positions {P} entries {int(llen.sum())} keys {int((llen > 0).sum())} "
          f"simple keys (one entry per run) {int(((llen > 0) & (llen == kruns)).sum())}  ({time.time() - t0:.1f} s)")
    # active runs per shared row
    D = sp.csr_matrix((np.ones(T, np.float32), (run_of_link, dst)), shape=(R, n))
    big = deg >= hub
    L = (D @ A).tocsc()
    hub_ids = np.nonzero(big)[0]
    L = L[:, hub_ids]
    touched = np.diff(L.indptr) > 0
    M = hub_ids[touched]
    L = L[:, touched].T.tocsr()   # [|M| x R]
    lc = L.tocoo()
    act = np.zeros((M.size, 2), np.uint64)
    lo = lc.col < 64
    np.bitwise_or.at(act[:, 0], lc.row[lo], np.uint64(1) << lc.col[lo].astype(np.uint64))
    np.bitwise_or.at(act[:, 1], lc.row[~lo], np.uint64(1) << (lc.col[~lo] - 64).astype(np.uint64))
    afold = act[:, 0] | act[:, 1]
    nact = np.bitwise_count(act[:, 0]).astype(np.int64) + np.bitwise_count(act[:, 1])
    npairs = np.asarray(L.sum(1)).ravel()
    print(f"shared rows {M.size} pairs {int(npairs.sum())} columns {int(deg[M].sum())}")

    Ecsc = E  # columns = l
    len_edges = [1, 2, 3, 5, 9, 17, 33, 65, 129, 1 << 30]
    act_edges = [1, 2, 3, 5, 9, 17, 33, 1 << 30]
    nl, na = len(len_edges) - 1, len(act_edges) - 1
    tab = {k: np.zeros((nl, na)) for k in ("cols", "keycols", "today", "exact", "hits", "useful", "hitcols")}
    Er = E.tocsr()
    seg_edges = np.array([1, 2, 3, 5, 9, 17, 33, 65, 1 << 30])
    seg_hist = np.zeros(seg_edges.size - 1)      # hit segments by length
    seg_ent = np.zeros(seg_edges.size - 1)       # their entries
    for c0 in range(0, M.size, 8192):
        rows = A[M[c0:c0 + 8192]].tocoo()
        l = rows.col
        mi = rows.row + c0
        ll = llen[l]
        h0 = sig[l, 0] & act[mi, 0]
        h1 = sig[l, 1] & act[mi, 1]
        hits = np.bitwise_count(h0).astype(np.int64) + np.bitwise_count(h1)
        tod = (fold[l] & afold[mi]) != 0
        # useful entries: sum over active runs of E[r, l]  -> via sparse product of the block
        Ab = sp.csr_matrix((np.ones(l.size, np.float32), (np.arange(l.size), l)), shape=(l.size, n))
        # E^T rows for l: [occ x R]; multiply elementwise with active mask of the row
        El = Ecsc.T.tocsr()[l]            # [occ x R]
        Lb = (L[mi] > 0).astype(np.float32)
        hitseg = El.multiply(Lb).tocsr()
        useful = np.asarray(hitseg.sum(1)).ravel()
        sl = hitseg.data[hitseg.data > 0]
        bi = np.searchsorted(seg_edges, sl, side="right") - 1
        np.add.at(seg_hist, bi, 1)
        np.add.at(seg_ent, bi, sl)
        li = np.searchsorted(len_edges, np.maximum(ll, 1), side="right") - 1
        ai = np.searchsorted(act_edges, nact[mi], side="right") - 1
        key = ll > 0
        for name, val in (("cols", np.ones(l.size)), ("keycols", key.astype(float)), ("today", ll * tod), ("exact", ll * (hits > 0)),
                          ("hits", hits.astype(float)), ("useful", useful), ("hitcols", (hits > 0).astype(float))):
            np.add.at(tab[name], (li, ai), val)
    tot = {k: v.sum() / 1e6 for k, v in tab.items()}
    print("totals (M): " + "  ".join(f"{k} {v:.2f}" for k, v in tot.items()))
    print("\nby list length (all rows):  cols(M) keycols today exact hitcols hits(run segs) useful")
    for i in range(nl):
        print(f"  len [{len_edges[i]:4d},{len_edges[i + 1] if len_edges[i + 1] < 1 << 30 else 'inf':>4})"
              + "".join(f" {tab[k][i].sum() / 1e6:9.2f}" for k in ("cols", "keycols", "today", "exact", "hitcols", "hits", "useful")))
    print("\nby active runs of the row:  cols(M) keycols today exact hitcols hits useful")
    for j in range(na):
        print(f"  act [{act_edges[j]:3d},{act_edges[j + 1] if act_edges[j + 1] < 1 << 30 else 'inf':>4})"
              + "".join(f" {tab[k][:, j].sum() / 1e6:9.2f}" for k in ("cols", "keycols", "today", "exact", "hitcols", "hits", "useful")))
    print("\nhit segments by length:   segments(M)  entries(M)")
    for i in range(seg_hist.size):
        print(f"  len [{seg_edges[i]:3d},{seg_edges[i + 1] if seg_edges[i + 1] < 1 << 30 else 'inf':>4}) {seg_hist[i] / 1e6:10.2f} {seg_ent[i] / 1e6:10.2f}")
    print("\nmixed(L, Amax): lists <= L entries walked whole under the exact filter; longer lists by run segments when the row has"
          " <= Amax active runs, whole otherwise\n    L  Amax   visits(M)  segment look-ups(M)")
    for Lth in (4, 8, 16, 32):
        for amax_i, amax in ((na, 1 << 30), (5, 16), (4, 8)):
            vis = seg = 0.0
            for i in range(nl):
                for j in range(na):
                    short = len_edges[i + 1] - 1 <= Lth
                    few = act_edges[j + 1] - 1 <= amax if amax < 1 << 30 else True
                    if short or not few:
                        vis += tab["exact"][i, j]
                    else:
                        vis += tab["useful"][i, j]
                        seg += tab["hits"][i, j]
            print(f"  {Lth:3d} {amax if amax < 1 << 30 else 'inf':>5} {vis / 1e6:10.1f} {seg / 1e6:12.1f}")
    print(f"\n({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
