"""Find host stalls in the bench's device-resident leg (diagnostic): per-call host times above a threshold."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
import ocn_b200.cn as cnmod
from ocn_b200 import synth, _lib
dev = "cuda:0"
g = synth.make_graph("citation2", device=dev)
G = ob.Graph(g.rowptr, g.col, g.n)
x = g.features(32, device=dev)
T = 65536
n = 13
e_all = g.query_edges(n * T, "stream", device=dev)
ip3 = torch.zeros(3, device=dev)
plan_stream = torch.cuda.Stream(device=dev)
LOG = []
def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        dt = 1e3 * (time.perf_counter() - t)
        if dt > 2.0:
            LOG.append((label or name, round(dt, 2)))
        return r
    setattr(obj, name, w)
wrap(cnmod, "_hub_workspace"); wrap(cnmod, "_borrow_colstat"); wrap(cnmod, "_return_colstat")
wrap(torch, "empty", "torch.empty"); wrap(torch, "zeros", "torch.zeros")
L = _lib.lib()
class LW:
    def __init__(self, L): self.L = L
    def __getattr__(self, k):
        f = getattr(self.L, k)
        def w(*a):
            t = time.perf_counter(); r = f(*a); dt = 1e3 * (time.perf_counter() - t)
            if dt > 2.0: LOG.append((k, round(dt, 2)))
            return r
        return w
_lib._lib = LW(L)
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    ob.reserve_stream_pool(4 << 30, dev)
with torch.cuda.stream(plan_stream):
    ob.reserve_stream_pool(1 << 30, dev)
import gc
gc.collect(); gc.disable()
NA = torch.cuda.memory_stats()["num_device_alloc"]
for rep in range(3):
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for s in range(n):
        LOG.clear()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            e = e_all[:, s * T:(s + 1) * T]
            sess = ob.CNSession(G, e, 2048, 3, 0, plan_stream=plan_stream)
            sess.build(3, True)
            sess.stats(5, 0.0, ip3, 0)
            out = sess.aggregate(x, 5, 0.0, ip3)
            sess.release()
        dt = 1e3 * (time.perf_counter() - t0)
        na = torch.cuda.memory_stats()["num_device_alloc"]
        if na != globals().get("NA", na):
            LOG.append(("cudaMalloc calls", na - NA, "reserved GB", round(torch.cuda.memory_reserved() / 2**30, 2)))
        NA = na
        if dt > 3.0 or LOG:
            print(f"rep {rep} step {s}: host {dt:.2f} ms  slow calls: {LOG}", flush=True)
    torch.cuda.synchronize()
    print(f"rep {rep}: {1e3*(time.perf_counter()-t_all)/n:.3f} ms/step wall", flush=True)
