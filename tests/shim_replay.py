"""Helpers shared by the trace-replay tests: an API table for an implementation of the torch_sparse / pygho subset
and the comparison of one replayed result with the recorded one."""
import torch


def api_table(torch_sparse, pygho, spspmm_mod, spmm_mod):
    import importlib
    import sys
    TS, PG = torch_sparse.SparseTensor, pygho.SparseTensor
    mm = torch_sparse.matmul      # the real package (and the shim) rebind this name to the matmul FUNCTION
    if not hasattr(mm, "spmm_add"):
        mm = importlib.import_module(torch_sparse.__name__ + ".matmul")
    api = {
        "ts.__init__": lambda *a, **k: TS(*a, **k),
        "pg.__init__": lambda *a, **k: PG(*a, **k),
        "ts.from_edge_index": TS.from_edge_index,
        "ts.from_torch_sparse_coo_tensor": TS.from_torch_sparse_coo_tensor,
        "ts.from_dense": TS.from_dense,
        "ts.masked_select_nnz": torch_sparse.masked_select_nnz,
        "ts.matmul.spmm_add": mm.spmm_add,
        "ts.matmul.spmm_mean": mm.spmm_mean,
        "ts.matmul.spmm_max": mm.spmm_max,
        "pg.Spspmm.spsphadamard": spspmm_mod.spsphadamard,
        "pg.Spspmm.spspmm": spspmm_mod.spspmm,
        "pg.Spmm.spmm": spmm_mod.spmm,
        "ts.__decode__": lambda e, dev: TS(row=e["row"].to(dev), col=e["col"].to(dev),
                                           value=None if e["val"] is None else e["val"].to(dev), sparse_sizes=e["sizes"],
                                           is_sorted=True),
        "pg.__decode__": lambda e, dev: PG(e["indices"].to(dev), None if e["values"] is None else e["values"].to(dev),
                                           e["shape"], is_coalesced=True),
    }
    for m in ("row", "col", "value", "has_value", "rowcount", "rowptr"):
        api[f"ts.storage.{m}"] = (lambda m: lambda owner, *a, **k: getattr(owner.storage, m)(*a, **k))(m)

    class _Methods(dict):
        def __missing__(self, name):
            tag, m = name.split(".", 1)
            return lambda obj, *a, **k: getattr(obj, m)(*a, **k)
    out = _Methods()
    out.update(api)
    return out


def _close(name, n, got, want):
    assert isinstance(got, torch.Tensor), f"call {n} {name}: expected a tensor, got {type(got)}"
    got = got.detach().cpu()
    assert tuple(got.shape) == tuple(want.shape), f"call {n} {name}: shape {tuple(got.shape)} != {tuple(want.shape)}"
    if want.dtype.is_floating_point or got.dtype.is_floating_point:
        g, w = got.double(), want.double()
        bound = 1e-5 * (1 + w.abs()) if w.numel() else w
        assert bool(((g - w).abs() <= bound).all()), f"call {n} {name}: max |diff| {(g - w).abs().max().item()}"
    else:
        assert torch.equal(got.long(), want.long()), f"call {n} {name}: integer result differs"


def check(name, n, got, want):
    if isinstance(want, dict):
        k = want["__kind__"]
        if k == "seq":
            assert len(got) == len(want["items"]), f"call {n} {name}: {len(got)} results, expected {len(want['items'])}"
            for g, w in zip(got, want["items"]):
                check(name, n, g, w)
        elif k == "ts":
            row, col, val = got.coo()
            assert tuple(got.sizes()) == tuple(want["sizes"]), f"call {n} {name}: sizes {got.sizes()} != {want['sizes']}"
            _close(name + " row", n, row, want["row"])
            _close(name + " col", n, col, want["col"])
            assert (val is None) == (want["val"] is None), f"call {n} {name}: value presence differs"
            if val is not None:
                _close(name + " value", n, val, want["val"])
        elif k == "pg":
            assert tuple(got.shape) == tuple(want["shape"]), f"call {n} {name}: shape {got.shape} != {want['shape']}"
            _close(name + " indices", n, got.indices, want["indices"])
            _close(name + " values", n, got.values.reshape(want["values"].shape), want["values"])
        elif k == "coo":
            c = got.coalesce()
            assert tuple(c.shape) == tuple(want["shape"]), f"call {n} {name}: shape differs"
            _close(name + " indices", n, c.indices(), want["indices"])
            _close(name + " values", n, c.values(), want["values"])
        else:
            raise AssertionError(f"call {n} {name}: unexpected recorded kind {k}")
    elif isinstance(want, torch.Tensor):
        _close(name, n, got, want)
    elif want is None:
        assert got is None, f"call {n} {name}: expected None"
    elif isinstance(want, float):
        assert abs(float(got) - want) <= 1e-6 * (1 + abs(want)), f"call {n} {name}: {got} != {want}"
    elif isinstance(want, str):
        pass    # a device name
    else:
        assert got == want or list(got) == list(want), f"call {n} {name}: {got} != {want}"
