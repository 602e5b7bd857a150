"""TEST INFRASTRUCTURE ONLY -- independent dense brute force for tiny graphs (N <= ~500).

Shares no code with oracle/ref_ops.py: it works on a dense 0/1 adjacency and python sets, and is
used to catch a wrongly "recalled" third-party semantic in the restatement.
"""
from __future__ import annotations

import numpy as np


def dense_adj(rowptr, col, n):
    a = np.zeros((n, n), dtype=np.int64)
    rowptr = np.asarray(rowptr)
    col = np.asarray(col)
    for r in range(n):
        a[r, col[rowptr[r]:rowptr[r + 1]]] = 1
    return a


def cn_sets(a, edges, order):
    """[(sorted node list, walk counts)] per edge for CN_order = A[i] * (A^order)[j]."""
    ak = np.linalg.matrix_power(a, order)
    out = []
    for i, j in zip(edges[0], edges[1]):
        v = a[i] * ak[j]
        idx = np.nonzero(v)[0]
        out.append((idx, v[idx]))
    return out


def cn1_python_sets(rowptr, col, edges):
    rowptr = np.asarray(rowptr)
    col = np.asarray(col)
    res = []
    for i, j in zip(edges[0], edges[1]):
        si = set(col[rowptr[i]:rowptr[i + 1]].tolist())
        sj = set(col[rowptr[j]:rowptr[j + 1]].tolist())
        res.append(sorted(si & sj))
    return res


def cn5_dense(a, edges, x, ip, order=2, weighted=True):
    """Dense restatement of the cn5/cn6 combination given a fixed inner-product value ``ip``
    (eval mode): returns [xcn1, xcn2(, xcn3)] as float64 arrays."""
    B, n = len(edges[0]), a.shape[0]
    C = []
    for k in range(1, order + 1):
        ak = np.linalg.matrix_power(a, k)
        m = np.stack([a[i] * ak[j] for i, j in zip(edges[0], edges[1])]).astype(np.float64)
        if not weighted:
            m = (m > 0).astype(np.float64)
        C.append(m)
    c1 = C[0].sum(0)
    w1 = np.where((c1 == 0) | (c1 == 1), 0.0, 1.0 / np.where(c1 == 0, 1, c1))
    H = [C[0] * w1[None, :]]
    pat = C[0] > 0
    scale = np.abs(H[0]).max() if pat.any() else 0.0
    outs = [H[0] @ x]
    for k in range(1, order):
        pat = pat | (C[k] > 0)
        sc = scale if pat.any() else 1.0
        coef = ip / sc if sc > 0 else ip
        v = C[k] - sum(coef * h for h in H)
        v = np.where(pat, v, 0.0)
        cs = v.sum(0)
        cs = np.where(cs == 0, 1.0, cs)
        H.append(v / cs[None, :])
        outs.append(H[-1] @ x)
    return outs
