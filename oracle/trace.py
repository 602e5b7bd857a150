"""TEST INFRASTRUCTURE ONLY -- record / replay of the third-party sparse API calls the reference's text makes.

``Recorder`` wraps the functions and methods of ``oracle/emul`` (the pure-torch stand-ins of torch_sparse / pygho).
While the reference's own ``utils.adjoverlap``, ``get_cn1_cn2`` and ``multidomainforward`` execute on top of them
(``oracle/make_trace.py``), every call made BY THE REFERENCE'S TEXT (not the stand-ins' internal calls) is logged with
its arguments and result.  ``replay`` (used by ``tests/test_gpu_shim.py``) feeds the same calls, in order, to another
implementation of the same API -- the CUDA shim ``ocn_b200/shim`` -- and compares every result.  Objects keep their
identity across calls (a matrix returned by call 3 and passed to call 7 is the SAME object in the replay), so lazy
objects of the shim travel exactly as they would under the reference's text.
"""
from __future__ import annotations

import functools
import inspect

import torch

# (module path inside the stand-in package, class or None, attribute) of everything the reference touches
TS_METHODS = ["sum", "mul", "coo", "csr", "sizes", "sparse_sizes", "size", "nnz", "has_value", "to_torch_sparse_coo_tensor",
              "to_dense", "coalesce", "to_symmetric", "__add__", "__getitem__", "index_select", "fill_value_", "fill_value",
              "set_value_"]
TS_STORAGE = ["row", "col", "value", "has_value", "rowcount", "rowptr"]
TS_CLASSMETHODS = ["from_edge_index", "from_torch_sparse_coo_tensor", "from_dense"]
PG_METHODS = ["index_select", "to_torch_sparse_coo", "sum"]


class Recorder:
    def __init__(self):
        self.calls = []
        self.depth = 0
        self.keep = []          # recorded objects stay alive so that id() is never reused
        self.seen = set()       # objects the log already holds in full: later ARGUMENTS refer to them by id
        self.kinds = {}

    # -- encoding ----------------------------------------------------------------------------
    def enc(self, o, arg=False):
        ts, pg = self.kinds["ts"], self.kinds["pg"]
        if isinstance(o, (ts, pg)):
            if arg and id(o) in self.seen:
                return {"__kind__": "ref", "oid": id(o)}
            self.seen.add(id(o))
        if isinstance(o, ts):
            self.keep.append(o)
            v = o._value
            return {"__kind__": "ts", "oid": id(o), "row": o._row.clone(), "col": o._col.clone(),
                    "val": None if v is None else v.clone(), "sizes": tuple(o._sizes)}
        if isinstance(o, pg):
            self.keep.append(o)
            return {"__kind__": "pg", "oid": id(o), "indices": o.indices.clone(),
                    "values": None if o.values is None else o.values.clone(), "shape": tuple(o.shape)}
        if isinstance(o, torch.Tensor):
            if o.is_sparse:
                c = o.coalesce()
                return {"__kind__": "coo", "indices": c.indices().clone(), "values": c.values().clone(), "shape": tuple(c.shape)}
            return o.detach().clone()
        if isinstance(o, (list, tuple)):
            return {"__kind__": "seq", "tuple": isinstance(o, tuple), "items": [self.enc(i, arg) for i in o]}
        if isinstance(o, torch.Size):
            return {"__kind__": "seq", "tuple": True, "items": list(o)}
        if o is None or isinstance(o, (bool, int, float, str, torch.dtype, torch.device)):
            return str(o) if isinstance(o, (torch.device,)) else o
        raise TypeError(f"trace: cannot encode {type(o)}")

    # -- wrapping ----------------------------------------------------------------------------
    def wrap(self, name, fn):
        rec = self

        @functools.wraps(fn)
        def inner(*args, **kwargs):
            top = rec.depth == 0
            entry = None
            if top:
                try:
                    entry = {"fn": name, "args": [rec.enc(a, True) for a in args],
                             "kwargs": {k: rec.enc(v, True) for k, v in kwargs.items()}}
                except TypeError:
                    entry = None        # an argument that cannot travel (a lambda): the call is not replayed
            rec.depth += 1
            try:
                out = fn(*args, **kwargs)
            finally:
                rec.depth -= 1
            if entry is not None:
                entry["out"] = rec.enc(out)
                rec.calls.append(entry)
            return out
        return inner

    def install(self, torch_sparse, pygho, spspmm_mod, spmm_mod):
        ts, pg = torch_sparse.SparseTensor, pygho.SparseTensor
        self.kinds = {"ts": ts, "pg": pg}
        rec = self
        storage_cls = type(ts(row=torch.zeros(0, dtype=torch.long), col=torch.zeros(0, dtype=torch.long),
                              sparse_sizes=(1, 1)).storage)
        for m in TS_METHODS:
            setattr(ts, m, self.wrap(f"ts.{m}", getattr(ts, m)))
        for m in TS_CLASSMETHODS:
            self._wrap_classmethod(ts, m, f"ts.{m}")
        for m in TS_STORAGE:
            self._wrap_storage(storage_cls, m)
        for cls, tag in ((ts, "ts"), (pg, "pg")):
            self._wrap_init(cls, tag)
        for m in PG_METHODS:
            setattr(pg, m, self.wrap(f"pg.{m}", getattr(pg, m)))
        torch_sparse.masked_select_nnz = self.wrap("ts.masked_select_nnz", torch_sparse.masked_select_nnz)
        for m in ("spmm_add", "spmm_mean", "spmm_max"):
            w = self.wrap(f"ts.matmul.{m}", getattr(torch_sparse.matmul, m))
            setattr(torch_sparse.matmul, m, w)
            setattr(torch_sparse, m, w)
        for m in ("spsphadamard", "spspmm"):
            setattr(spspmm_mod, m, self.wrap(f"pg.Spspmm.{m}", getattr(spspmm_mod, m)))
        spmm_mod.spmm = self.wrap("pg.Spmm.spmm", spmm_mod.spmm)

    def _wrap_classmethod(self, cls, m, name):
        raw = getattr(cls, m).__func__
        setattr(cls, m, staticmethod(self.wrap(name, lambda *a, **k: raw(cls, *a, **k))))

    def _wrap_storage(self, storage_cls, m):
        raw = getattr(storage_cls, m)
        w = self.wrap(f"ts.storage.{m}", lambda owner, *a, **k: raw(owner.storage, *a, **k))
        setattr(storage_cls, m, lambda self_, *a, **k: w(self_._o, *a, **k))

    def _wrap_init(self, cls, tag):
        init = cls.__init__
        rec = self

        def __init__(self_, *a, **k):
            def build(*aa, **kk):
                init(self_, *aa, **kk)
                return self_
            rec.wrap(f"{tag}.__init__", build)(*a, **k)
        cls.__init__ = __init__

    def take(self):
        out, self.calls = self.calls, []
        self.seen = set()
        return out


# ---- replay -------------------------------------------------------------------------------------

def replay(calls, api, device, check):
    """``api``: dict name -> callable of the implementation under test; ``check(name, index, got, want)`` compares one
    encoded result.  Returns the number of calls replayed."""
    live = {}

    def dec(o):
        if isinstance(o, dict):
            k = o["__kind__"]
            if k == "seq":
                items = [dec(i) for i in o["items"]]
                return tuple(items) if o["tuple"] else items
            if k == "coo":
                return torch.sparse_coo_tensor(o["indices"].to(device), o["values"].to(device), o["shape"]).coalesce()
            if o["oid"] in live:
                return live[o["oid"]]
            if k == "ref":
                raise KeyError("trace refers to an object that no earlier call produced")
            obj = api[f"{k}.__decode__"](o, device)
            live[o["oid"]] = obj
            return obj
        if isinstance(o, torch.Tensor):
            return o.to(device)
        if isinstance(o, str) and o in ("cpu",):
            return device
        return o

    def bind(enc, obj):
        if isinstance(enc, dict):
            if enc["__kind__"] == "seq":
                for e, ob in zip(enc["items"], obj):
                    bind(e, ob)
            elif "oid" in enc and enc["__kind__"] != "ref":
                live[enc["oid"]] = obj

    for n, c in enumerate(calls):
        args = [dec(a) for a in c["args"]]
        kwargs = {k: dec(v) for k, v in c["kwargs"].items()}
        got = api[c["fn"]](*args, **kwargs)
        check(c["fn"], n, got, c["out"])
        bind(c["out"], got)
    return len(calls)
