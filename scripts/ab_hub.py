"""One-process A/B of library options on the bench workload (citation2 shape, 65 536-link slices of the evaluation
stream, order 3): per option set the dominant kernel alone (library timing events), the build stage, and the whole
device-resident step (plan on its own stream), averaged over the same slices.

    python scripts/ab_hub.py "" "hub_exact=1" "hub_seg_ctas=4" [--slices 10] [--first 3]
"""
import argparse
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob  # noqa: E402
from ocn_b200 import _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sets", nargs="*", default=[""])
    ap.add_argument("--slices", type=int, default=10)
    ap.add_argument("--first", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = "cuda:0"
    g = synth.make_graph("citation2", device=dev)
    G = ob.Graph(g.rowptr, g.col, g.n)
    x = g.features(32, device=dev)
    T = 65536
    e_all = g.query_edges((a.first + a.slices) * T, "stream", device=dev)
    ip3 = torch.zeros(3, device=dev)
    L = _lib.lib()
    plan_stream = torch.cuda.Stream(device=dev)
    ob.reserve_stream_pool(4 << 30, dev)
    with torch.cuda.stream(plan_stream):
        ob.reserve_stream_pool(1 << 30, dev)

    def step(s, ps=None):
        e = e_all[:, s * T:(s + 1) * T]
        sess = ob.CNSession(G, e, 2048, 3, 0, plan_stream=ps).build(3, True)
        sess.stats(5, 0.0, ip3, 0)
        out = sess.aggregate(x, 5, 0.0, ip3)
        sess.release()
        return out

    for spec in a.sets:
        _lib.reset_options()
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            _lib.set_option(k, int(v))
        for s in range(a.first):
            step(s, plan_stream)
        torch.cuda.synchronize()
        best = None
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(a.first, a.first + a.slices):
                step(s, plan_stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.slices
            best = ms if best is None else min(best, ms)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); k1.record()  # torch only reads events it has seen recorded
        torch.cuda.synchronize()
        kern, build = [], []
        for s in range(a.first, a.first + a.slices):
            e = e_all[:, s * T:(s + 1) * T]
            sess = ob.CNSession(G, e, 2048, 3, 0)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            L.ocn_cn_hub_timing_events(k0.cuda_event, k1.cuda_event)
            b0.record()
            sess.build(3, True)
            b1.record()
            torch.cuda.synchronize()
            L.ocn_cn_hub_timing_events(None, None)
            sess.release()
            kern.append(k0.elapsed_time(k1))
            build.append(b0.elapsed_time(b1))
        print(f"[{spec or 'default'}] step {best:.4f} ms ({T / best / 1e3:.1f} M links/s)  hub kernel alone {sum(kern) / len(kern):.4f} ms  "
              f"build stage (serialised) {sum(build) / len(build):.4f} ms", flush=True)
    _lib.reset_options()


if __name__ == "__main__":
    main()
