"""Per-step host / device times of the bench's device-resident leg (diagnostic)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth
dev = "cuda:0"
g = synth.make_graph("citation2", device=dev)
G = ob.Graph(g.rowptr, g.col, g.n)
x = g.features(32, device=dev)
T = 65536
n = 16
e_all = g.query_edges(n * T, "stream", device=dev)
ip3 = torch.zeros(3, device=dev)
plan_stream = torch.cuda.Stream(device=dev)
import ocn_b200.cn as cnmod
TIMES = {}
def timed_wrap(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        TIMES[name] = TIMES.get(name, 0.0) + 1e3 * (time.perf_counter() - t)
        return r
    return w
cnmod._hub_workspace = timed_wrap("hub_ws", cnmod._hub_workspace)
cnmod._borrow_colstat = timed_wrap("colstat", cnmod._borrow_colstat)
for mode in ("work-stream", "default-stream"):
    st = torch.cuda.Stream(device=dev) if mode == "work-stream" else torch.cuda.current_stream()
    torch.cuda.synchronize()
    print(mode, "reserved GB", torch.cuda.memory_reserved() / 2**30, flush=True)
    for s in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            a.record()
            e = e_all[:, s * T:(s + 1) * T]
            sess = ob.CNSession(G, e, 2048, 3, 0, plan_stream=plan_stream)
            t1 = time.perf_counter()
            TIMES.clear()
            sess.build(3, True)
            tb = time.perf_counter()
            sess.stats(5, 0.0, ip3, 0)
            out = sess.aggregate(x, 5, 0.0, ip3)
            sess.release()
            b.record()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"  step {s:2d}: plan host {1e3*(t1-t0):6.2f} ms  enqueue {1e3*(t2-t1):6.2f} ms  sync {1e3*(t3-t2):6.2f} ms  device {a.elapsed_time(b):6.2f} ms  reserved {torch.cuda.memory_reserved()/2**30:5.1f} GB  build host {1e3*(tb-t1):6.2f} ({TIMES}) positions {sess.plan_host[11]} entries {sess.plan_host[10]}", flush=True)
