#!/usr/bin/env python
"""bench.py -- higher-order CN aggregation links/s on a citation2-shape synthetic graph.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU op sequence (oracle port)

One *session* = one pass of the hot path (CN sets of orders 1..3 -> batch normalisation /
orthogonalisation -> CN-indicator SpMM + pair term) over ``--batches`` consecutive link batches of
2048 links in the shape of the citation2 evaluation stream (NeighborOverlapCitation2.py:241-254:
every source against 1000 uniform destinations).  One *step* = ``--sessions`` consecutive sessions
(default 48: 3.1 M links, ~50 ms, so that 20 steps time about a second; round 1 timed 25 ms in all).
Under torchrun every rank holds a replica of the graph and features and scores its own sessions
(weak scaling, no data-path collective); in the end-to-end arm every rank reads its scores back per
session and the ranks gather all fp32 scores once at the end (NCCL).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte accounting.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

# every kernel is loaded at start-up: with CUDA's default lazy loading the first launch of a kernel variant that the
# warm-up steps did not reach (e.g. the CTA-wide counter variant a hub source needs) stalls a timed step for milliseconds
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "higher-order CN aggregation edges/sec (citation2 shape)"
UNIT = "edges/s"


# The ONE JSON line goes to the real stdout; everything else a library may print there (e.g. "NCCL version ..."
# from the communicator init when NCCL_DEBUG is set) is diverted to stderr.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ocn_b200", choices=["ocn_b200", "reference"])
    ap.add_argument("--graph", default="citation2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--batches", type=int, default=32, help="link batches per session (one plan / build / aggregate pass)")
    ap.add_argument("--sessions", type=int, default=48, help="sessions per step")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` measurements (SURVEY 8(d))")
    ap.add_argument("--feat", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=48, help="links of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hub", type=int, default=0, help="hub_degree of the indexed order-3 path: 0 auto, -1 off (per-run tables)")
    ap.add_argument("--no-plan-stream", action="store_true", help="plan on the main stream (no overlap with the previous step)")
    ap.add_argument("--streams", type=int, default=2,
                    help="end-to-end leg: consecutive steps alternate between this many CUDA streams, so that the copies, "
                         "the small head / gather kernels and the D2H read of one step overlap the walks of the next "
                         "(measured 41.0 -> 43.2 M links/s); the device-resident leg always uses one stream")
    ap.add_argument("--slice-offset", type=int, default=0,
                    help="skip this many sessions' worth of links of the stream (one GPU re-enacts the sessions another rank gets)")
    ap.add_argument("--deal", choices=("cost", "contiguous"), default="cost",
                    help="several ranks: 'cost' deals the timed slices of the stream so that every rank gets the same "
                         "number of slices and a balanced predicted cost (ocn_b200.dist.predicted_walk_cost / "
                         "deal_by_cost); 'contiguous' gives rank r the r-th run of consecutive slices")
    ap.add_argument("--device-streams", type=int, default=1,
                    help="device-resident leg: consecutive steps alternate between this many of the --streams streams")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library tuning option (ocn_b200._lib.OPTIONS), e.g. --opt hub_exact=1 for an A/B run")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed device steps (ncu --profile-from-start off)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 2 ms period, 5 ms with several ranks; nvidia-smi
    as a fallback).  A daemon thread polls while the steps run."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, period: float = 0.002):
        self.index, self.sm, self.bits, self.stop_flag, self.thread = index, [], 0, False, None
        self.period = period
        self.max_mhz, self.h = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _sample(self):
        if self.h is not None:
            self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            parts = [p.strip() for p in out.strip().split(",")]
            if len(parts) >= 2 and parts[0].isdigit():
                self.sm.append(int(parts[0]))
                self.max_mhz = int(parts[1])

    def _run(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        self.sm, self.bits, self.stop_flag = [], 0, False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(self.sm)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(G, e, order, feat, batch, hub_d=0):
    """Bytes the fused path must move for these links with no cache credit (DESIGN.md "Byte accounting").
    Returns (build_stage_bytes, whole_step_bytes, survey_formula_bytes, shared_row_kernel_bytes); the last is the
    part of SURVEY.md 8(d)'s per-link index term that runs through rows of >= hub_d columns -- the walk
    k_cn_hub_count performs once per stream instead of once per link."""
    import ocn_b200 as ob  # noqa: F401
    deg = G.degree()
    colL = G.col.long()
    Fv = torch.zeros(G.n, dtype=torch.int64, device=G.device).index_add_(0, G.row(), deg[colL])
    i, j = e[0], e[1]
    di, dj, Fi, Fj = deg[i], deg[j], Fv[i], Fv[j]
    T = e.shape[1]
    # j side, once per (link, 32-position chunk of N(i)): rowptr pair + N(j) + rowptr pairs of N(j) + frontier
    chunks = torch.div(di + 31, 32, rounding_mode="floor").clamp(min=0)
    per_link_j = 16 + (4 * dj if order >= 2 else 0) + ((8 * dj + 4 * Fj) if order >= 3 else 0)
    j_bytes = int((per_link_j * chunks.clamp(min=1)).sum())
    # i side, once per work unit (run x chunk x 32-link sub-list): N(i) chunk, its rowptr pairs, its rows
    t = torch.arange(T, device=e.device)
    first = (t == 0) | (torch.cat((i[:1] - 1, i[:-1])) != i)   # runs cross batch boundaries (ocn_cn_plan)
    run_id = torch.cumsum(first.long(), 0) - 1
    run_len = torch.bincount(run_id)
    run_src = i[first]
    subs = torch.div(run_len + 31, 32, rounding_mode="floor")
    i_bytes = int((subs * (16 + 12 * deg[run_src] + 4 * Fv[run_src])).sum())
    rec_bytes = int(8 * di.sum())
    build = j_bytes + i_bytes + rec_bytes + 16 * T
    # aggregate + stats + release: records re-read 3x, col of N(i), colstat 32 B per non-empty record (upper bound: all),
    # feature gathers 4F per non-empty record (upper bound), outputs
    agg = int((3 * 8 + 4 + 32 + 4 * feat) * di.sum()) + T * (4 * feat * (order + 1) + 2 * 4 * feat + 32)
    survey = int((32 + 4 * (di + dj) + (8 * dj + 4 * Fj if order >= 2 else 0) + (8 * di + 4 * Fi if order >= 3 else 0)).sum()) \
        + T * 4 * feat * (order + 1 + 2)
    hub = 0
    if hub_d > 0 and order >= 3:
        big = deg >= hub_d
        Fh = torch.zeros(G.n, dtype=torch.int64, device=G.device).index_add_(0, G.row(), (8 + 4 * deg[colL]) * big[colL])
        hub = int(Fh[j].sum())
    return build, build + agg, survey, hub


def run_reference(a, rank, world):
    """CPU arm: the oracle's restatement of the reference op sequence (pygho-style get_cn ->
    cn6/cn5 combination -> spmm_add), all host threads, on a bounded sample per step."""
    if rank != 0:
        return
    from ocn_b200 import synth
    from oracle import ref_ops as R
    torch.set_num_threads(os.cpu_count() or 1)
    g = synth.make_graph(a.graph, device="cuda:0" if torch.cuda.is_available() else "cpu", scale=a.scale)
    rowptr, col = g.rowptr.cpu(), g.col.cpu()
    A = R.sp_from_csr(rowptr, col)
    x = g.features(a.feat, device="cpu")
    S = a.cpu_sample
    e_all = g.query_edges((a.steps + a.warmup) * S, "stream", device="cpu")
    st = R.InnerProdState()
    times = []
    for s in range(a.steps + a.warmup):
        e = e_all[:, s * S:(s + 1) * S]
        t0 = time.perf_counter()
        cns = R.get_cn(A, e, a.order)
        if a.order >= 3:
            R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, st, training=False)
        else:
            R.cn5_aggregate(cns[0], cns[1], x, e, st, training=False)
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            times.append(dt)
    total = sum(times)
    val = S * len(times) / total
    sample = f"{S} links per step of the same citation2-shape stream (order {a.order}, F={a.feat}), torch CPU ops"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, g),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(a, g):
    T = a.batch * a.batches
    return {"workload": f"{a.graph}-shape synthetic graph (N={g.n}, nnz={g.nnz}), cn5 order {a.order} (cn6 template) "
                        f"scoring, F={a.feat}, evaluation link stream (each source x 1000 uniform destinations)",
            "link_batch": a.batch, "batches_per_session": a.batches, "links_per_session": T,
            "sessions_per_step": a.sessions, "links_per_step_per_gpu": T * a.sessions,
            "weighted": True, "l2": "inputs larger than L2 (CSR col 244 MB + fresh links every session)",
            "parallelism": f"graph replicated, sessions of {a.batches} link batches sharded over {a.gpus} GPU(s)"}


def cpu_baseline(a, g, rowptr, col):
    from oracle import ref_ops as R
    torch.set_num_threads(os.cpu_count() or 1)
    A = R.sp_from_csr(rowptr, col)
    x = g.features(a.feat, device="cpu")
    S = a.cpu_sample
    e_all = g.query_edges(2 * S, "stream", device="cpu")
    best = None
    for s in range(2):
        e = e_all[:, s * S:(s + 1) * S]
        t0 = time.perf_counter()
        cns = R.get_cn(A, e, a.order)
        if a.order >= 3:
            R.cn6_aggregate(cns[0], cns[1], cns[2], x, e, R.InnerProdState(), training=False)
        else:
            R.cn5_aggregate(cns[0], cns[1], x, e, R.InnerProdState(), training=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": S / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{S} links of the same stream (order {a.order}, F={a.feat}); oracle/ref_ops.py get_cn + "
                      f"cn6_aggregate, best of 2; the GPU arm's steps are {a.batch * a.batches * a.sessions} links each"}


def kernel_source_sha():
    """sha1 over the sources of the dominant kernel: profiles/traffic.json carries the one its ncu capture was taken
    from, and `roofline.traffic` is only reported when the two agree."""
    import hashlib
    h = hashlib.sha1()
    for f in ("cn_hub.cu", "common.cuh"):
        with open(os.path.join(ROOT, "ocn_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def percentile(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))] if xs else None


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def secondary(a, ob, synth, G, g, x, sampler, peak, dev):
    """The other measurements SURVEY.md 8(d) lists, each CUDA-event timed (median of 5 after 2 warm-ups) with the SM
    clock sampled while it runs: GNN SpMM GB/s, A^2 SpGEMM products/s, order 2 (the reference-pinned K = 2), one
    optimiser step, and configs 1-4 through the fused path.  Kept under ~20 s."""
    out = {}

    def measure(name, fn, extra, reps=5, warm=2):
        sampler.start()
        ms = timeit(fn, reps, warm)
        clk = sampler.stop()
        rec = {"ms": ms, "sm_mhz": clk["sm_mhz"], "reasons": clk["reasons"]}
        rec.update(extra(ms))
        out[name] = rec

    # (1) GNN neighbour aggregation at citation2 shape (PureConv3 / GCNConv propagate, model.py:128-142)
    norm = ob.gcn_norm(G)
    for F in (32, 128):
        xf = x if F == a.feat else g.features(F, device=dev)
        byt = 8 * (g.n + 1) + 4 * g.nnz + 4 * F * g.nnz + 4 * F * g.n
        measure(f"gnn_spmm_gcn_citation2_F{F}", lambda: ob.pure_conv(xf, G, "gcn", norm),
                lambda ms: {"alg_GBs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / peak})
    xf = g.features(128, device=dev)
    byt = 8 * (g.n + 1) + 4 * g.nnz + 4 * 128 * g.nnz + 4 * 128 * g.n
    measure("gnn_spmm_sum_citation2_F128_bulk_gather", lambda: ob.pure_conv(xf, G, "sum"),
            lambda ms: {"alg_GBs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / peak,
                        "kernel": "k_spmm_tma (cp.async.bulk + mbarrier)"})
    del xf
    byt = 8 * (g.n + 1) + 4 * g.nnz + 4 * a.feat * g.nnz + 4 * a.feat * g.n
    measure(f"gnn_spmm_pureconv3_gcn_citation2_F{a.feat}", lambda: ob.pure_conv3_gcn(x, G, norm),
            lambda ms: {"alg_GBs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / peak})
    # (1b) the predictor head of one session (65 536 links, cn6, in = hidden = 32): tensor cores against CUDA cores / torch
    from ocn_b200 import _lib as _l
    torch.manual_seed(0)
    hp = ob.CNLinkPredictor3hopCNs(a.feat, a.feat, 1, 3, 0.0, weighted=True).to(dev).eval()
    Th = a.batch * a.batches
    hx = [torch.randn(Th, a.feat, device=dev) for _ in range(4)]
    if a.feat == 32:
        with torch.no_grad():
            for label, opt, fuse in (("tcgen05", 3, True), ("cuda_cores", 2, True), ("torch_modules", 2, False)):
                _l.set_option("head_tc", opt)
                hp.fuse_head = fuse
                measure(f"head_cn6_F32_{label}", lambda: hp._head(hx[0], hx[1], hx[2], hx[3]),
                        lambda ms: {"links": Th, "us": ms * 1e3}, reps=5, warm=2)
            _l.set_option("head_tc", 0)
    del hp, hx
    # (1c) the wider heads of configs 1-3 (hidden 256): one tensor-core launch per layer against the torch modules
    from ocn_b200 import head as _head
    torch.manual_seed(0)
    hp = ob.CNLinkPredictorOringin(256, 256, 1, 3, 0.0, ln=True).to(dev).eval()
    hx = [torch.randn(Th, 256, device=dev) for _ in range(3)]
    with torch.no_grad():
        measure("head_cn5_F256_tcgen05_per_layer", lambda: _head.fused_head_wide(hp, hx[0], hx[1], None, hx[2]),
                lambda ms: {"links": Th, "us": ms * 1e3}, reps=3, warm=1)
        hp.fuse_head = False
        measure("head_cn5_F256_torch_modules", lambda: hp._head(hx[0], hx[1], None, hx[2]),
                lambda ms: {"links": Th, "us": ms * 1e3}, reps=3, warm=1)
    del hp, hx
    # (2) order 2 on the same stream (get_cn1_cn2 as the reference's citation2 driver calls it)
    T = a.batch * a.batches
    e2 = g.stream_links(7 * T, 4 * T, device=dev)
    ip3 = torch.zeros(3, device=dev)

    def order2():
        for k in range(4):
            s = ob.CNSession(G, e2[:, k * T:(k + 1) * T], a.batch, 2).build(2, True)
            s.stats(5, 0.0, ip3, 0)
            s.aggregate(x, 5, 0.0, ip3)
            s.release()
    measure("cn5_order2_citation2_stream", order2, lambda ms: {"links": 4 * T, "Mlinks_per_s": 4 * T / ms / 1e3}, reps=3, warm=1)
    # (3) one optimiser step of the citation2 driver's predictor loop (NeighborOverlapCitation2.py:131-209)
    from ocn_b200.dist import sharded_train_step
    torch.manual_seed(0)
    pred = ob.CNLinkPredictorOringin(a.feat, a.feat, 1, 3, 0.0, weighted=True).to(dev).train()
    h = x.clone().requires_grad_(True)
    pos = g.query_edges(16384, "pos", device=dev)
    neg = torch.stack((pos[0], synth.hash_randint(16384, g.n, 5, 9, dev)))
    subs = [pos[:, k:k + 2048] for k in range(0, 16384, 2048)] + [neg[:, k:k + 2048] for k in range(0, 16384, 2048)]
    signs = [1.0] * 8 + [-1.0] * 8

    def train_step():
        pred.zero_grad(set_to_none=True)
        h.grad = None
        return sharded_train_step(pred, h, G, subs, signs, 16384, 0, 1)
    measure("training_step_cn5_order2_32768_links", train_step, lambda ms: {"Mlinks_per_s": 32768 / ms / 1e3}, reps=3, warm=1)
    del pred, h
    # (4) configs 1-4 of BASELINE.json: A^2 (true) and the fused order-2 path at each config's batch and width
    for name in ("cora", "pubmed", "collab", "ddi"):
        gg = synth.make_graph(name, device=dev)
        GG = ob.Graph(gg.rowptr, gg.col, gg.n)
        deg = GG.degree()
        prods = int(deg[gg.col.long()].sum())
        if name in ("collab", "ddi"):
            measure(f"spgemm_a2_{name}", lambda: ob.spgemm_a2(GG, 0, True),
                    lambda ms: {"products": prods, "Gproducts_per_s": prods / ms / 1e6}, reps=3, warm=1)
        e = gg.query_edges(gg.batch, "mixed", device=dev)
        xx = gg.features(gg.hidden, device=dev)
        variant = 5 if gg.predictor == "cn5" else 7

        def fused():
            s = ob.CNSession(GG, e, gg.batch, 2).build(2, False)
            if variant == 5:
                s.stats(5, 0.0, ip3, 0)
            r = s.aggregate(xx, variant, 1.0 if variant == 7 else 0.0, ip3)
            s.release()
            return r
        measure(f"fused_{gg.predictor}_order2_{name}", fused,
                lambda ms: {"B": gg.batch, "F": gg.hidden, "Mlinks_per_s": gg.batch / ms / 1e3}, reps=3, warm=1)
        del GG, gg
    return out


def main():
    a = parse()
    rank, world, local = dist_env()
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    import ocn_b200 as ob
    from ocn_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (ocn_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    for kv in a.opt:
        k, v = kv.split("=")
        ob._lib.set_option(k, int(v))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))

    g = synth.make_graph(a.graph, device=dev, scale=a.scale)
    G = ob.Graph(g.rowptr, g.col, g.n)
    x = g.features(a.feat, device=dev)
    T = a.batch * a.batches          # links per session
    S = a.sessions                   # sessions per step
    nsteps = a.steps + a.warmup
    nsess = nsteps * S               # sessions of this rank; session q of rank r covers stream sessions r * nsess + q
    base = a.slice_offset + rank * nsess
    ids = list(range(base, base + nsess))  # stream session of every local session
    dealing = "contiguous sessions per rank"
    if world > 1 and a.deal == "cost":
        # sessions differ in cost (the time follows the index entries a session writes; one in ~26 holds a hub source):
        # the timed sessions of all ranks are dealt by predicted cost, same count per rank, no session split.  The
        # prediction is exact integer arithmetic on the replicated graph, so the ranks agree without a collective.
        try:
            from ocn_b200 import dist as obdist
            src_all = g.stream_links(a.slice_offset * T, world * nsess * T, device=dev)[0]
            cost = obdist.predicted_walk_cost(g.rowptr, g.col, src_all, T, a.batch).tolist()
            del src_all
            pool = [r * nsess + a.warmup * S + k for r in range(world) for k in range(a.steps * S)]
            mine = obdist.deal_by_cost([cost[i] for i in pool], world, a.steps * S)[rank]
            ids = ids[:a.warmup * S] + [a.slice_offset + pool[i] for i in mine]
            dealing = "timed sessions dealt by predicted index entries (longest first, equal count per rank)"
        except Exception as ex:  # never lose a measurement to the dealing: fall back to the contiguous sessions
            dealing = f"contiguous sessions per rank (cost dealing failed: {type(ex).__name__}: {ex})"
    e_rank = torch.cat([g.stream_links(i * T, T, device=dev) for i in ids], dim=1).contiguous()
    e_host = e_rank.cpu().pin_memory()
    torch.manual_seed(0)
    cls = ob.CNLinkPredictor3hopCNs if a.order >= 3 else ob.CNLinkPredictorOringin
    pred = cls(a.feat, a.feat, 1, 3, 0.0, weighted=True).to(dev).eval()
    ip3 = torch.zeros(3, device=dev)

    # the plan of a session (and its size read-back) runs on its own stream: the host waits for the plan only,
    # while the main stream is still executing the previous session
    plan_stream = None if a.no_plan_stream else torch.cuda.Stream(device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, a.streams))]
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    # one cached block per stream (and one for the default stream of the roofline leg): the per-session buffers are
    # carved out of it, no cudaMalloc (1 - 40 ms when it happens) inside a timed region
    ob.reserve_stream_pool(2 << 30, dev)
    for st_ in streams + [st2 for st2 in (plan_stream, comm_stream) if st2 is not None]:
        with torch.cuda.stream(st_):
            ob.reserve_stream_pool(4 << 30 if st_ in streams else 1 << 30, dev)

    def fork():  # the work streams start after everything already queued on the default stream ...
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())

    def join():  # ... and the default stream (where the timing events are recorded) waits for all of them
        for st in streams + ([comm_stream] if comm_stream is not None else []):
            torch.cuda.current_stream().wait_stream(st)

    L = ob._lib.lib()
    sess_events = []  # (start, end) CUDA events of every timed session of the device-resident leg

    def session_device(q, timed):
        st = streams[q % max(1, min(a.device_streams, len(streams)))]
        with torch.cuda.stream(st):
            ev = None
            if timed:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            e = e_rank[:, q * T:(q + 1) * T]
            sess = ob.CNSession(G, e, a.batch, a.order, a.hub, plan_stream=plan_stream).build(a.order, True)
            sess.stats(5, 0.0, ip3, 0)
            out = sess.aggregate(x, 5, 0.0, ip3)
            sess.release()
            if timed:
                ev[1].record()
                sess_events.append(ev)
        return out

    out_host = torch.empty(a.steps * S * T, dtype=torch.float32).pin_memory()
    rank_scores = torch.empty(a.steps * S, T, dtype=torch.float32, device=dev) if world > 1 else None
    # (rank 0 alone reads the gathered scores: the other ranks do not pin a world-sized host buffer)
    gathered_host = torch.empty(world * a.steps * S * T, dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None

    def session_e2e(q, timed):
        with torch.cuda.stream(plan_stream if plan_stream is not None else torch.cuda.current_stream()):
            e = e_host[:, q * T:(q + 1) * T].to(dev, non_blocking=True)
        with torch.no_grad(), torch.cuda.stream(streams[q % len(streams)]):
            sess = ob.CNSession(G, e, a.batch, a.order, a.hub, plan_stream=plan_stream).build(a.order, True)
            if a.order >= 3:
                out = pred(x, G, sess, None, None, e)
            else:
                out = pred(x, G, sess, None, e)
            scores = out.squeeze(-1).contiguous()
            # device -> host read of the session's result: asynchronous copy into pinned memory (an evaluation loop
            # collects the scores of every batch and ranks them at the end); the timed region ends with a full sync.
            # With several ranks every rank reads back its own scores and the ranks meet once per STEP (48 sessions): the
            # step's scores of all ranks are gathered on the communication stream and read by rank 0 while the next step's
            # sessions run (north_star: "the final gather of scores").  A gather per session would couple the ranks; one
            # gather after the last step left rank 0's read of world x 126 MB over PCIe (26 - 77 ms at 8 GPUs, depending on
            # the box) at the end of the timed region with nothing to hide it behind.
            if timed:
                k = q - a.warmup * S
                out_host[k * T:(k + 1) * T].copy_(scores, non_blocking=True)
                if world > 1:
                    rank_scores[k].copy_(scores)
                    if (k + 1) % S == 0:
                        step_gather(k // S)
        return scores

    allsc_steps = [torch.empty(world * S * T, dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None

    def step_gather(step):
        """All ranks' scores of one timed step, gathered on the communication stream and read by rank 0 (layout of
        gathered_host: [step][rank][S * T]); the work streams go on with the next step meanwhile."""
        import torch.distributed as dist
        for st in streams:
            comm_stream.wait_stream(st)
        with torch.cuda.stream(comm_stream):
            allsc = allsc_steps[step & 1]  # (two buffers: the copy of step s - 2 has long left when s overwrites it -- same stream)
            dist.all_gather_into_tensor(allsc, rank_scores[step * S:(step + 1) * S].reshape(-1))
            if rank == 0:
                gathered_host[step * world * S * T:(step + 1) * world * S * T].copy_(allsc, non_blocking=True)

    def final_gather():
        """(The last step's gather was queued by its last session; join() makes the timing event wait for it.)"""
        return

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # NVML is initialised (and queried once) before any timed region: its first calls take driver locks for
    # milliseconds and would otherwise stall the launches of the first timed sessions
    sampler = ClockSampler(local, 0.002 if world == 1 else 0.005)  # N ranks share the host cores: poll less often
    try:
        sampler._sample()
    except Exception:
        pass

    def timed(fn, profile=False, finalize=None):
        for q in range(a.warmup * S):
            fn(q, False)
        barrier()
        if profile:
            torch.cuda.cudart().cudaProfilerStart()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()  # a generation-2 collection in the middle of the loop stalls the launches for 10+ ms
        launches0 = L.ocn_launch_count()
        t0 = time.perf_counter()
        ev0.record()
        fork()
        for q in range(a.warmup * S, nsess):
            fn(q, True)
        if finalize is not None:
            finalize()
        join()
        ev1.record()
        barrier()
        gc.enable()
        if profile:
            torch.cuda.cudart().cudaProfilerStop()
        wall = time.perf_counter() - t0
        launches = L.ocn_launch_count() - launches0
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, clocks, launches

    ms_dev, _, clocks, launches_dev = timed(session_device, a.profile_range)
    sess_ms = [e0.elapsed_time(e1) for e0, e1 in sess_events]
    ms_e2e_ev, wall_e2e, _, _ = timed(session_e2e, finalize=final_gather)
    ms_e2e = max(ms_e2e_ev, wall_e2e * 1e3)  # the D2H read ends after the last event: use the host clock too

    # roofline of the dominant kernel: CUDA events recorded by the library on the build's stream right
    # before / after k_cn_hub_count (indexed path) -- or around the whole build stage when the per-run
    # table kernel k_cn_build is in use (--hub -1) -- fresh links each launch, over a sample of the timed sessions
    build_ms, kern_ms, hub_ds, sample_q = [], [], [], []
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(); k1.record()
    torch.cuda.synchronize()
    stride = max(1, (nsess - a.warmup * S) // 24)
    for q in range(a.warmup * S, nsess, stride):
        e = e_rank[:, q * T:(q + 1) * T]
        sess = ob.CNSession(G, e, a.batch, a.order, a.hub)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        indexed = sess.hub_degree > 0
        if indexed:
            L.ocn_cn_hub_timing_events(k0.cuda_event, k1.cuda_event)
        ev0.record()
        sess.build(a.order, True)
        ev1.record()
        torch.cuda.synchronize()
        L.ocn_cn_hub_timing_events(None, None)
        sess.release()
        build_ms.append(ev0.elapsed_time(ev1))
        kern_ms.append(k0.elapsed_time(k1) if indexed else build_ms[-1])
        hub_ds.append(sess.hub_degree)
        sample_q.append(q)
    build_avg = sum(build_ms) / len(build_ms)
    kern_avg = sum(kern_ms) / len(kern_ms)
    indexed = hub_ds[-1] > 0
    bb = [algorithmic_bytes(G, e_rank[:, q * T:(q + 1) * T], a.order, a.feat, a.batch, hd) for q, hd in zip(sample_q, hub_ds)]
    build_bytes = sum(b[0] for b in bb) / len(bb)
    survey_bytes = sum(b[2] for b in bb) / len(bb)
    kern_bytes = (sum(b[3] for b in bb) / len(bb)) if indexed else build_bytes
    kern_name = "k_cn_hub_count" if indexed else "k_cn_build"
    peak, peak_src = peaks()
    achieved = kern_bytes / (kern_avg * 1e-3) / 1e9
    # DRAM traffic and issue-slot use of that kernel from the committed ncu capture -- only if it was taken from the
    # kernel source this library was built from
    traffic, ncu_facts, traffic_note = None, {}, "profiles/traffic.json missing"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("kernel_source_sha") == kernel_source_sha():
            traffic = tj.get(kern_name + "_dram_bytes_per_launch")
            ncu_facts = tj.get(kern_name + "_ncu", {})
            traffic_note = tj.get("source", "")
        else:
            traffic_note = (f"profiles/traffic.json was captured from kernel source {tj.get('kernel_source_sha')}, this "
                            f"tree is {kernel_source_sha()}: not reported")

    if rank == 0:
        ms_session = ms_dev / (a.steps * S)
        links = world * T * S * a.steps
        value = links / (ms_dev * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 walk counts / f32 features", "data": "synthetic", "config": workload_config(a, g),
            "clocks": clocks,
            "timed_region_ms": ms_dev,
            "session_ms": {"mean": ms_session, "p50": percentile(sess_ms, 0.5), "p95": percentile(sess_ms, 0.95),
                           "max": max(sess_ms) if sess_ms else None, "n": len(sess_ms),
                           "note": "per-session figures are event pairs on the work stream of rank 0 (plan excluded: it runs "
                                   "a session ahead on its own stream)"},
            "e2e": {"value": links / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * T * S * world,
                    "d2h_bytes_per_step": 4 * T * S * world, "ms_per_step": ms_e2e / a.steps,
                    "final_gather_bytes": (4 * T * S * a.steps * world * world) if world > 1 else 0,
                    "gather": "none (one rank)" if world == 1 else "all ranks' scores of a step all-gathered (NCCL) and read by rank 0 on a "
                              "communication stream while the next step runs; the last step's gather is inside the timed region",
                    "api": "CNLinkPredictor*.forward(h, adj, CNSession, ..., edges) -> scores.cpu()"},
            # launches of the library's own kernels inside the timed region of the device-resident leg, counted by the
            # library (ocn_launch_count; the CUB scans / radix sorts it calls are not in the figure), this rank
            "gpu_launches": int(launches_dev),
            "gpu_launches_per_session": launches_dev / (a.steps * S),
            "roofline": {"bound": "hbm" if not indexed else "issue",
                         "bound_note": ("the contract's roofline is the HBM one (frac = algorithmic bytes / time / measured copy "
                                        "bandwidth); ncu shows this kernel bound by instruction issue and L2 latency, not by "
                                        "DRAM: see dram_frac and ncu below") if indexed else "",
                         "kernel": kern_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "dram_gbs": (traffic / (kern_avg * 1e-3) / 1e9) if traffic else None,
                         "dram_frac": (traffic / (kern_avg * 1e-3) / 1e9 / peak) if traffic else None,
                         "ncu": ncu_facts, "traffic_source": traffic_note,
                         "algorithmic_bytes_per_launch": kern_bytes, "kernel_ms": kern_avg,
                         "kernel_share_of_session": kern_avg / ms_session,
                         "build_stage_ms": build_avg, "hub_degree": hub_ds[-1], "sessions_sampled": len(kern_ms),
                         "whole_session_survey_gbs": survey_bytes / (ms_session * 1e-3) / 1e9,
                         "whole_session_survey_frac": survey_bytes / (ms_session * 1e-3) / 1e9 / peak,
                         "note": "achieved = SURVEY 8(d) per-link index bytes of the rows this kernel covers "
                                 "(8 + 4 d(m) for every (link, m in N(dst)) with d(m) >= hub_degree) / its duration; the "
                                 "kernel streams each such row once per session, so the figure exceeds the DRAM traffic"},
        }
        if world > 1:
            line["dealing"] = dealing
        if world == 1 and not a.no_secondary:
            try:
                line["secondary"] = secondary(a, ob, synth, G, g, x, sampler, peak, dev)
            except Exception as ex:  # the headline line is never lost to a secondary measurement
                line["secondary"] = {"error": f"{type(ex).__name__}: {ex}"}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, g, g.rowptr.cpu(), g.col.cpu())
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
