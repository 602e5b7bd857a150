"""The recorded API traces of the reference's text (tests/golden/ref_trace_*.pt, oracle/make_trace.py):

* CPU: replaying a trace against the stand-ins that produced it reproduces every result (the record / replay
  machinery, object identity included, is sound), the traces cover the method census of SURVEY 7 step 2, and the
  CUDA shim exposes every recorded name and refuses CPU tensors (no fallback).
* GPU (tests/test_gpu_shim.py): the same replay against ocn_b200/shim.
"""
import glob
import importlib
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import trace as T  # noqa: E402
import shim_replay  # noqa: E402

TRACES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_trace_*.pt")))


def _emul():
    """oracle/emul imported under private names (the public names may belong to the shim in this process)."""
    emul = os.path.join(ROOT, "oracle", "emul")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("torch_sparse", "pygho")}
    sys.path.insert(0, emul)
    try:
        mods = [importlib.import_module(n) for n in ("torch_sparse", "pygho", "pygho.backend.Spspmm", "pygho.backend.Spmm")]
    finally:
        sys.path.remove(emul)
        for k in [k for k in sys.modules if k.split(".")[0] in ("torch_sparse", "pygho")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    return mods


def test_traces_exist_and_cover_the_census():
    assert len(TRACES) >= 6
    names = set()
    for p in TRACES:
        names |= {c["fn"] for c in torch.load(p)["calls"]}
    for need in ("ts.__init__", "ts.__getitem__", "ts.sum", "ts.mul", "ts.coo", "ts.sizes", "ts.size", "ts.storage.row",
                 "ts.storage.col", "ts.to_torch_sparse_coo_tensor", "ts.fill_value_", "ts.from_edge_index", "ts.to_symmetric",
                 "ts.from_torch_sparse_coo_tensor", "ts.matmul.spmm_add", "ts.matmul.spmm_mean", "ts.matmul.spmm_max",
                 "ts.masked_select_nnz", "pg.__init__", "pg.index_select", "pg.to_torch_sparse_coo",
                 "pg.Spspmm.spsphadamard", "pg.Spspmm.spspmm"):
        assert need in names, f"no recorded call of {need}"


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p)[10:-3] for p in TRACES])
def test_replay_against_the_stand_ins_that_recorded_it(path):
    ts, pg, spspmm_mod, spmm_mod = _emul()
    fx = torch.load(path)
    n = T.replay(fx["calls"], shim_replay.api_table(ts, pg, spspmm_mod, spmm_mod), torch.device("cpu"), shim_replay.check)
    assert n == len(fx["calls"]) > 0


def test_shim_exposes_every_recorded_name_and_has_no_cpu_path():
    import ocn_b200.shim as shim
    from ocn_b200._lib import OcnError
    sts = importlib.import_module("ocn_b200.shim.torch_sparse")
    spg = importlib.import_module("ocn_b200.shim.pygho")
    sps = importlib.import_module("ocn_b200.shim.pygho.backend.Spspmm")
    spm = importlib.import_module("ocn_b200.shim.pygho.backend.Spmm")
    api = shim_replay.api_table(sts, spg, sps, spm)
    a = sts.SparseTensor(row=torch.tensor([0, 0, 1]), col=torch.tensor([0, 2, 1]), sparse_sizes=(3, 3))
    p = spg.SparseTensor(torch.tensor([[0, 0, 1], [0, 2, 1]]), torch.ones(3), (3, 3), is_coalesced=True)
    for path in TRACES:
        for c in torch.load(path)["calls"]:
            tag, m = c["fn"].split(".", 1)
            if c["fn"] in api.keys():
                continue
            assert hasattr(a if tag == "ts" else p, m), f"the shim lacks {c['fn']}"
    # containers hold CPU data (the drivers build the graph on the host), computing on it is refused
    assert a.coo()[0].tolist() == [0, 0, 1] and a.sizes() == [3, 3]
    for call in (lambda: a.sum(dim=0), lambda: a[torch.tensor([1])], lambda: sts.spmm_add(a, torch.ones(3, 2)),
                 lambda: sts.masked_select_nnz(a, torch.tensor([True, False, True])), lambda: p.sum(dims=1),
                 lambda: sps.spsphadamard(p, p), lambda: spm.spmm(p, 1, torch.ones(3, 2)),
                 lambda: sps.spspmm(p, 1, p, 0).indices):
        with pytest.raises(OcnError):
            call()
    # install() / uninstall() swap the public names
    before = sys.modules.get("torch_sparse")
    shim.install(force=True)
    try:
        import torch_sparse
        from pygho.backend.Spspmm import spsphadamard  # noqa: F401
        from torch_geometric.nn import GCNConv  # noqa: F401
        assert torch_sparse.SparseTensor is sts.SparseTensor
    finally:
        shim.uninstall()
    assert sys.modules.get("torch_sparse") is before


REF = os.environ.get("OCN_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model.py")), reason="the reference's files are not on this machine")
def test_reference_modules_import_and_build_over_the_shim():
    """The reference's own model.py / utils.py import against the stand-ins and its classes construct (no compute: CPU)."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r); import ocn_b200.shim as shim; shim.install(force=True); sys.path.insert(0, %r);\n"
        "import model, utils; shim.accelerate(utils, model)\n"
        "assert model.adjoverlap.__wrapped__ is not None and utils.adjoverlap is model.adjoverlap\n"
        "for k in ('cn2', 'cn3', 'cn4', 'cn5', 'cn6', 'cn7'): model.predictor_dict[k](64, 64, 1, 3, 0.0)\n"
        "for fn in ('gcn', 'gin', 'sage', 'max', 'puregcn'): model.GCN(8, 8, 8, 2, 0.0, conv_fn=fn)\n"
        "print('ok')\n") % (ROOT, REF)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
