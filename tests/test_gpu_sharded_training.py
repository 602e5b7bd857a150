"""Two ranks (sharing cuda:0, gloo for the two tiny exchanges) run one optimiser step of the citation2
driver's predictor loop (NeighborOverlapCitation2.py:131-209) through ``ocn_b200.dist.sharded_train_step``;
the summed gradients, the loss and the inner-product buffer must equal the sequential single-process loop."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(variant):
    import ocn_b200 as ob
    from ocn_b200 import synth
    g = synth.make_graph("citation2", scale=0.002)
    G = ob.Graph(g.rowptr.to(DEV), g.col.to(DEV), g.n)
    torch.manual_seed(1)
    cls = ob.CNLinkPredictorOringin if variant == 5 else ob.CNLinkPredictorbaselearn
    pred = cls(16, 16, 1, 3, 0.0, weighted=True).to(DEV).train()
    h = g.features(16).to(DEV).requires_grad_(True)
    pos = g.query_edges(5 * 96, "pos").cpu()
    neg = torch.stack((pos[0], torch.randint(0, g.n, (pos.shape[1],), generator=torch.Generator().manual_seed(3))))
    subs = [pos[:, k:k + 96].to(DEV) for k in range(0, pos.shape[1], 96)] + \
           [neg[:, k:k + 96].to(DEV) for k in range(0, neg.shape[1], 96)]
    signs = [1.0] * 5 + [-1.0] * 5
    return G, pred, h, subs, signs, pos.shape[1]


def _sequential(variant):
    G, pred, h, subs, signs, total = _setup(variant)
    loss = 0.0
    for e, sg in zip(subs, signs):
        if variant == 5:
            out = pred.multidomainforward(h, G, None, None, e)
        else:
            out = pred.multidomainforward(h, G, None, None, e, 1.0)
        l = -(1.0 / total) * F.logsigmoid(sg * out).sum()
        l.backward()
        loss += float(l.detach())
    return pred, h, loss


def _worker(rank, world, port, variant, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ocn_b200.dist import sharded_train_step
    G, pred, h, subs, signs, total = _setup(variant)
    loss = sharded_train_step(pred, h, G, subs, signs, total, rank, world, fill=1.0 if variant == 7 else 0.0)
    torch.cuda.synchronize()
    if rank == 0:
        # numpy arrays are pickled by value (tensors would be shared through file descriptors of a process
        # that may already have exited)
        out.put(({k: v.grad.cpu().numpy() for k, v in pred.named_parameters() if v.grad is not None},
                 h.grad.cpu().numpy(), float(loss), pred.innerprod.cpu().numpy(), pred.n))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("variant", [5, 7])
def test_sharded_train_step_matches_sequential(variant):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, out)) for r in range(world)]
    for p in procs:
        p.start()
    grads, hgrad, loss, ip, n = out.get(timeout=300)
    grads = {k: torch.from_numpy(v) for k, v in grads.items()}
    hgrad, ip = torch.from_numpy(hgrad), torch.from_numpy(ip)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    pred, h, loss_ref = _sequential(variant)
    assert abs(loss - loss_ref) <= 1e-5 * (1 + abs(loss_ref))
    if variant == 5:
        assert n == pred.n == 10
        assert torch.allclose(ip, pred.innerprod.cpu(), rtol=1e-6, atol=0)
    for k, v in pred.named_parameters():
        if v.grad is None:
            assert k not in grads or bool((grads[k] == 0).all())
            continue
        scale = float(v.grad.abs().max()) + 1e-12
        assert float((grads[k] - v.grad.cpu()).abs().max()) <= 2e-4 * scale + 1e-7, k
    scale = float(h.grad.abs().max()) + 1e-12
    assert float((hgrad - h.grad.cpu()).abs().max()) <= 2e-4 * scale + 1e-7
