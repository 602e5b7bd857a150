"""CPU census of the work `k_cn_hub_count` (ocn_b200/csrc/cn_hub.cu) does on one slice of the bench stream.

No GPU and no product code path is involved: the slice is rebuilt with numpy / scipy from the same seeded
synthetic graph and link stream as `bench.py`, and the entry visits of the indexed order-3 build are counted

  * as the kernel walks them today: every shared row N(m), d(m) >= hub_degree, next to some destination is
    streamed once and the WHOLE entry list of each of its columns l is counted (optionally only lists that hold
    an entry of a run active for m -- what the 64-bit run signature approximates);
  * as they are needed: only the entries whose run r has a link next to m ("useful");
  * under a candidate layout: the runs cut into G groups, one entry list per (column, group), a row walked once
    per group that is active for it (`--groups`).

Output: totals, the split by row degree and by number of active runs, and the G-group what-if table.  Used to
decide the next change of the kernel before any device time is spent.

    python scripts/hub_visit_census.py [--slice 0] [--links 65536] [--batch 2048] [--hub 0]
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocn_b200 import synth  # noqa: E402  (host-side generator only; works without a GPU)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="citation2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--slice", type=int, default=0)
    ap.add_argument("--links", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--hub", type=int, default=0, help="hub_degree; 0 = the plan's rule max(32, ceil(n / links))")
    ap.add_argument("--groups", default="1,2,4,8,16,32,0", help="run-group counts to evaluate (0 = one group per run)")
    a = ap.parse_args()

    t0 = time.time()
    g = synth.make_graph(a.graph, device="cpu", scale=a.scale)
    n = g.n
    rowptr = g.rowptr.numpy().astype(np.int64)
    col = g.col.numpy().astype(np.int32)
    deg = np.diff(rowptr)
    T = a.links
    e = g.query_edges((a.slice + 1) * T, "stream", device="cpu").numpy()[:, a.slice * T:(a.slice + 1) * T]
    src, dst = e[0], e[1]
    hub = a.hub if a.hub > 0 else max(32, -(-n // T))
    print(f"graph {a.graph} n={n} nnz={col.size}  slice {a.slice}: {T} links, batch {a.batch}, hub_degree {hub} "
          f"({time.time() - t0:.1f} s)")

    A = sp.csr_matrix((np.ones(col.size, np.float32), col, rowptr), shape=(n, n))

    # runs: maximal pieces of one source inside one link batch (k_plan_edges)
    t = np.arange(T)
    first = (t == 0) | (np.r_[-1, src[:-1]] != src)   # runs cross batch boundaries (round 2; round 1 cut them: t % batch == 0)
    run_of_link = np.cumsum(first) - 1
    run_src = src[first]
    R = run_src.size
    # positions (r, p) -> k = N(src_r)[p];  entries (l, position) for l in N(k)
    pos_run = np.repeat(np.arange(R), deg[run_src])
    pos_k = np.concatenate([col[rowptr[s]:rowptr[s + 1]] for s in run_src]) if R else np.zeros(0, np.int32)
    P = pos_k.size
    # E[r, l] = entries of column l that belong to run r  (= sum_p [l in N(k_{r,p})])
    K = sp.csr_matrix((np.ones(P, np.float32), (pos_run, pos_k)), shape=(R, n))
    E = (K @ A).tocsr()
    entries = int(E.sum())
    # pairs (t, m): m in N(dst_t), d(m) >= hub;  Act[m, r] = some link of run r is next to m
    D = sp.csr_matrix((np.ones(T, np.float32), (run_of_link, dst)), shape=(R, n))   # links per (run, destination)
    big = deg >= hub
    L = (D @ A).tocsc()                       # L[r, m] = links of run r next to m
    L = L[:, np.nonzero(big)[0]]              # hub rows only
    hub_ids = np.nonzero(big)[0]
    touched = np.diff(L.indptr) > 0
    M = hub_ids[touched]
    L = L[:, touched].T.tocsr()               # [|M| x R]
    pairs = int(L.sum())
    print(f"runs {R}  positions {P}  entries {entries}  pairs {pairs}  shared rows {M.size}  "
          f"columns of the shared rows {int(deg[M].sum())}")
    # the other side of hub_degree: rows below it are walked by every link next to them (k_cn_link)
    small_cols = A @ np.where(big, 0, deg).astype(np.float64)          # per node j: columns of its rows below hub_degree
    print(f"per-link walk (k_cn_link): {small_cols[dst].sum() / 1e6:.1f} M column look-ups through rows below hub_degree, "
          f"{(A @ (~big).astype(np.float64))[dst].sum() / 1e6:.2f} M such (link, row) pairs")

    # W[m, r] = entries of run r met while streaming N(m)  = sum_{l in N(m)} E[r, l]
    W = np.asarray((A[M] @ E.T).todense(), dtype=np.float64)      # [|M| x R]
    Act = np.asarray((L > 0).todense())
    n_act = Act.sum(1)
    total = W.sum()
    useful = (W * Act).sum()
    # lists that hold at least one entry of an active run (exact version of the signature filter)
    Eb = (E > 0).astype(np.float32).T.tocsr()                      # [n x R]
    tot_l = np.asarray(E.sum(0)).ravel()                           # whole list length per column
    filt = 0.0
    for lo in range(0, M.size, 4096):
        rows = A[M[lo:lo + 4096]].tocoo()
        hit = np.asarray(Eb[rows.col].multiply(sp.csr_matrix(Act[lo:lo + 4096].astype(np.float32))[rows.row]).sum(1)).ravel() > 0
        filt += tot_l[rows.col][hit].sum()
    print(f"entry visits: whole lists {total / 1e6:.1f} M   lists with an active run {filt / 1e6:.1f} M   "
          f"useful {useful / 1e6:.1f} M  ({useful / total:.3f} of all)")

    print("\nby row degree d(m):   rows   columns    visits(M)  useful(M)  mean active runs")
    edges_d = [hub, 64, 128, 256, 512, 1024, 4096, 1 << 30]
    dm = deg[M]
    for lo, hi in zip(edges_d[:-1], edges_d[1:]):
        s = (dm >= lo) & (dm < hi)
        if s.any():
            print(f"  [{lo:5d}, {hi if hi < 1 << 30 else 'inf':>5}) {int(s.sum()):7d} {int(dm[s].sum()):9d} {W[s].sum() / 1e6:10.1f} "
                  f"{(W[s] * Act[s]).sum() / 1e6:10.1f} {n_act[s].mean():10.1f}")
    print("\nby active runs of the row:   rows    visits(M)  useful(M)")
    for lo, hi in [(1, 2), (2, 3), (3, 5), (5, 9), (9, 17), (17, 33), (33, 1 << 30)]:
        s = (n_act >= lo) & (n_act < hi)
        if s.any():
            print(f"  [{lo:3d}, {hi if hi < 1 << 30 else 'inf':>3}) {int(s.sum()):10d} {W[s].sum() / 1e6:10.1f} {(W[s] * Act[s]).sum() / 1e6:10.1f}")

    print("\nwhat-if: runs cut into G contiguous groups, one list per (column, group), a row walked per active group")
    print("   G   entry visits(M)   column look-ups(M)   (today: G = 1)")
    for G in [int(x) for x in a.groups.split(",")]:
        Gn = R if G == 0 else min(G, R)
        grp = (np.arange(R) * Gn) // R
        S = np.zeros((R, Gn))
        S[np.arange(R), grp] = 1
        Wg = W @ S                                  # [|M| x Gn]
        Ag = (Act @ S) > 0
        print(f" {Gn:4d} {(Wg * Ag).sum() / 1e6:14.1f} {(Ag.sum(1) * dm).sum() / 1e6:18.1f}")
    # routing by the REALISED number of links next to a row instead of by its degree alone: rows with few links
    # gain nothing from being streamed once, and pay for the entries of every run
    npairs = np.asarray(L.sum(1)).ravel()
    lookups_link = (npairs * dm)                    # column look-ups if each (link, m) pair walks N(m) itself
    print("\nwhat-if: rows with at most `theta` links next to them leave the shared pass and are walked per link "
          "(k_cn_link looks (l, run of t) up)")
    print(" theta    rows   shared pass: look-ups(M) visits(M)   per-link instead: look-ups(M) entries met(M)   "
          "left in shared pass: look-ups(M) visits(M)")
    for theta in (0, 1, 2, 3, 4, 6, 8):
        s = npairs <= theta
        print(f" {theta:5d} {int(s.sum()):7d} {dm[s].sum() / 1e6:22.1f} {W[s].sum() / 1e6:9.1f} "
              f"{lookups_link[s].sum() / 1e6:28.1f} {(W[s] * Act[s]).sum() / 1e6:14.1f} "
              f"{dm[~s].sum() / 1e6:30.1f} {W[~s].sum() / 1e6:9.1f}")
    print(f"\n({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
