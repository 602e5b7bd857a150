"""Fused inference head (csrc/head.cu) against the torch modules that mirror the reference's
nn.Sequential heads (model.py:2192-2235), for every flag combination and served width; the golden fixtures
of tests/test_golden_reference.py (*_f32 / *_f64) pin the same kernel to the reference's own scores."""
import itertools

import pytest
import torch

import ocn_b200 as ob
from ocn_b200 import head

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("in_ch,hid", [(32, 32), (64, 32), (32, 64), (64, 64)])
@pytest.mark.parametrize("ln,tailact,two", list(itertools.product([False, True], repeat=3)))
@pytest.mark.parametrize("cls", ["cn5", "cn6"])
def test_fused_head_matches_modules(in_ch, hid, ln, tailact, two, cls):
    torch.manual_seed(in_ch + hid + 2 * ln + 4 * tailact + 8 * two)
    P = ob.predictor_dict[cls]
    pred = P(in_ch, hid, 2 if two else 1, 3, 0.0, ln=ln, tailact=tailact, twolayerlin=two).to(DEV).eval()
    with torch.no_grad():
        pred.alpha.copy_(torch.tensor([0.3, -0.2, 1.1]))
        pred.beta.fill_(0.7)
    B = 1000 + 3  # not a multiple of the links a warp carries
    xs = [torch.randn(B, in_ch, device=DEV) * s for s in (1.0, 3.0, 0.5, 2.0)]
    x3 = xs[2] if cls == "cn6" else None
    n_params = sum(p.numel() for n, p in pred.named_parameters()
                   if n.split(".")[0] in ("xcn1lin", "xcn2lin", "xijlin", "lin") or (cls == "cn6" and n.startswith("xcn3lin")))
    if 4 * n_params > 200 * 1024:  # does not fit the shared-memory budget: the torch modules serve it
        assert head.supported(pred, in_ch) == -1
        return
    assert head.supported(pred, in_ch) == n_params
    with torch.no_grad():
        got = pred._head(xs[0], xs[1], x3, xs[3])
        pred.fuse_head = False
        want = pred._head(xs[0], xs[1], x3, xs[3])
    assert got.shape == want.shape
    # fp32 fma chains in a different order than cuBLAS; compare in float64 against the modules' own result
    err = (got.double() - want.double()).abs().max().item()
    assert err <= 2e-5 * (1.0 + want.abs().max().item()), err


def test_unserved_widths_and_training_keep_the_modules():
    pred = ob.CNLinkPredictorOringin(256, 256, 1, 3, 0.0).to(DEV).eval()
    assert head.supported(pred, 256) == -1
    x = torch.randn(10, 256, device=DEV)
    with torch.no_grad():
        assert pred._head(x, x, None, x).shape == (10, 1)
    odd = ob.CNLinkPredictorOringin(48, 48, 1, 3, 0.0).to(DEV).eval()      # not a multiple of 32 anywhere: torch modules
    assert head.supported(odd, 48) == -1 and not head.wide_supported(odd, 48)
    x = torch.randn(5000, 48, device=DEV)
    with torch.no_grad():
        assert odd._head(x, x, None, x).shape == (5000, 1)
    pred = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0).to(DEV).train()
    x = torch.randn(10, 32, device=DEV, requires_grad=True)
    out = pred._head(x, x, None, x)
    out.sum().backward()  # autograd through the torch modules
    assert x.grad is not None


def test_parameter_cache_follows_updates():
    pred = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0).to(DEV).eval()
    x = torch.randn(64, 32, device=DEV)
    with torch.no_grad():
        a = pred._head(x, x, None, x)
        pred.lin[-1].bias.add_(1.0)
        b = pred._head(x, x, None, x)
    assert torch.allclose(b - a, torch.ones_like(a), atol=1e-5)


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("cls,ln", [("cn5", False), ("cn6", True)])
def test_head_kernel_variants_at_32(variant, cls, ln, lib_options):
    """in = hidden = 32: tcgen05 with the activations through shared memory (1) / through tensor memory (3, the default)
    and the CUDA-core kernel (2), each against the torch modules evaluated in float64, over several tiles per pipeline."""
    import copy
    lib_options(head_tc=variant)
    torch.manual_seed(variant)
    pred = ob.predictor_dict[cls](32, 32, 1, 3, 0.0, ln=ln).to(DEV).eval()
    B = 148 * 4 * 128 * 2 + 77           # more than two rounds of tiles on every pipeline, ragged tail
    xs = [torch.randn(B, 32, device=DEV) * s for s in (1.0, 3.0, 0.5, 2.0)]
    x3 = xs[2] if cls == "cn6" else None
    p64 = copy.deepcopy(pred).double()
    p64.fuse_head = False
    with torch.no_grad():
        got = pred._head(xs[0], xs[1], x3, xs[3])
        want = p64._head(xs[0].double(), xs[1].double(), None if x3 is None else x3.double(), xs[3].double())
    err = (got.double() - want).abs().max().item()
    assert err <= 5e-6 * (1.0 + want.abs().max().item()), err


@pytest.mark.parametrize("cls,in_ch,hid,ln,tailact,two,out_ch", [
    ("cn5", 64, 64, True, False, False, 1), ("cn6", 64, 64, False, True, True, 1), ("cn5", 128, 128, True, False, False, 1),
    ("cn7", 256, 256, False, False, False, 1), ("cn5", 256, 256, True, False, True, 3), ("cn6", 128, 256, True, True, False, 1),
    ("cn5", 256, 64, False, False, False, 2), ("cn5", 96, 128, True, False, False, 1)])
def test_wide_head_on_the_tensor_cores(cls, in_ch, hid, ln, tailact, two, out_ch):
    """ocn_linear_tc (tcgen05, hi/lo operand split, one launch per layer with its tail fused) against the torch modules in
    float64: hidden 64 / 128 / 256, LayerNorm, tailact, twolayerlin, several outputs, inputs that are not a power of two."""
    import copy
    torch.manual_seed(in_ch + hid + out_ch)
    pred = ob.predictor_dict[cls](in_ch, hid, out_ch, 3, 0.0, ln=ln, tailact=tailact, twolayerlin=two).to(DEV).eval()
    with torch.no_grad():
        pred.alpha.copy_(torch.tensor([0.3, -0.2, 1.1]))
        pred.beta.fill_(0.7)
    assert head.wide_supported(pred, in_ch)
    B = 148 * 128 * 2 + 77                 # more tiles than resident CTAs, ragged tail
    xs = [torch.randn(B, in_ch, device=DEV) * s for s in (1.0, 3.0, 0.5, 2.0)]
    x3 = xs[2] if cls == "cn6" else None
    p64 = copy.deepcopy(pred).double()
    p64.fuse_head = False
    with torch.no_grad():
        got = head.fused_head_wide(pred, xs[0], xs[1], x3, xs[3])
        want = p64._head(xs[0].double(), xs[1].double(), None if x3 is None else x3.double(), xs[3].double())
        if head.supported(pred, in_ch) < 0:     # and it is what _head picks at this size
            from ocn_b200 import _lib
            before = _lib.lib().ocn_launch_count()
            again = pred._head(xs[0], xs[1], x3, xs[3])
            assert _lib.lib().ocn_launch_count() > before and torch.equal(again, got)
    assert got.shape == want.shape == (B, out_ch)
    err = (got.double() - want).abs().max().item()
    assert err <= 1e-5 * (1.0 + want.abs().max().item()), err
    # a parameter update reaches the cached split weights
    with torch.no_grad():
        pred.lin[0].weight.mul_(1.5)
        p64.lin[0].weight.mul_(1.5)
        got = head.fused_head_wide(pred, xs[0], xs[1], x3, xs[3])
        want = p64._head(xs[0].double(), xs[1].double(), None if x3 is None else x3.double(), xs[3].double())
    assert (got.double() - want).abs().max().item() <= 1e-5 * (1.0 + want.abs().max().item())


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 5000])
@pytest.mark.parametrize("k,n", [(32, 32), (32, 256), (256, 32), (64, 128), (96, 64), (256, 256)])
def test_linear_tc_shapes_and_tails(rows, k, n):
    """ocn_linear_tc on its own: fewer rows than a tile, one K chunk (no ring refill), an odd number of chunks, every
    output width; Linear, Linear + LayerNorm + ReLU with the z accumulation and the fused last Linear, against float64."""
    torch.manual_seed(rows + k + n)
    lin = torch.nn.Linear(k, n).to(DEV)
    ln = torch.nn.LayerNorm(n).to(DEV)
    fin = torch.nn.Linear(n, 3).to(DEV)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_()
    holder = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0)      # only carries the cache of split weights
    x = torch.randn(rows, k, device=DEV) * 2
    z0 = torch.randn(rows, n, device=DEV)
    with torch.no_grad():
        out, _ = head.linear_tc(holder, x, lin)
        want = lin.double()(x.double())
        lin.float()
        assert (out.double() - want).abs().max().item() <= 1e-5 * (1 + want.abs().max().item())
        z = z0.clone()
        out2, f2 = head.linear_tc(holder, x, lin, ln, True, want_out=True, z=z, z_scale=0.7, z_accumulate=True, final=fin)
        v = torch.relu(ln.double()(lin.double()(x.double())))
        lin.float(); ln.float()
        tol = 1e-5 * (1 + v.abs().max().item())
        assert (out2.double() - v).abs().max().item() <= tol
        assert (z.double() - (z0.double() + 0.7 * v)).abs().max().item() <= tol
        wf = fin.double()(v)
        fin.float()
        assert f2.shape == (rows, 3) and (f2.double() - wf).abs().max().item() <= 1e-5 * (1 + wf.abs().max().item())
        z = z0.clone()
        head.linear_tc(holder, x, lin, None, False, want_out=False, z=z, z_scale=-1.25, z_accumulate=False)
        assert (z.double() - (-1.25) * want).abs().max().item() <= 1e-5 * (1 + want.abs().max().item())


def test_linear_tc_refuses_what_it_does_not_serve():
    from ocn_b200 import _lib
    L = _lib.lib()
    assert L.ocn_linear_tc_prep_floats(48, 64) == -1 and L.ocn_linear_tc_prep_floats(64, 48) == -1
    assert L.ocn_linear_tc_prep_floats(256, 256) == 2 * 256 * 256
    x = torch.randn(8, 64, device=DEV)
    buf = torch.empty(2 * 64 * 64, device=DEV)
    b = torch.zeros(64, device=DEV)
    rc = L.ocn_linear_tc(x.data_ptr(), 8, 64, 64, buf.data_ptr(), b.data_ptr(), None, None, 0, None, None, 1.0, 0, None, None, 0, None, None)
    assert rc != 0 and b"no output" in L.ocn_last_error()
