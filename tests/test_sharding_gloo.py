"""World-size-2 gloo test of the multi-GPU host logic: whole link batches are dealt to ranks, each
rank scores its batches independently, scores are all-gathered.  The scoring function
is a stand-in (the CUDA path needs a GPU); what is checked is the dealing, the gather layout and that
no batch is ever split."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocn_b200.dist import deal_batches, gather_scores


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, bs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    edges = torch.stack((torch.arange(T), torch.arange(T) * 7 % 1000))
    mine = deal_batches(T, bs, rank, world)
    # stand-in per-batch scorer with a batch-coupled term (like the column statistics)
    local = []
    for (s, e) in mine:
        b = edges[:, s:e]
        local.append(b[0].float() * 2 + b[1].float().sum())
    scores = gather_scores(local, mine, T)
    if rank == 0:
        ref = []
        for s in range(0, T, bs):
            b = edges[:, s:s + bs]
            ref.append(b[0].float() * 2 + b[1].float().sum())
        out.put(bool(torch.equal(scores, torch.cat(ref))))
    dist.destroy_process_group()


def test_batches_dealt_and_scores_gathered():
    T, bs, world = 10 * 64 + 13, 64, 2
    all_b = [deal_batches(T, bs, r, world) for r in range(world)]
    flat = sorted(x for b in all_b for x in b)
    assert flat[0][0] == 0 and flat[-1][1] == T and all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
    assert all((e - s) == bs for (s, e) in flat[:-1])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, bs, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True


def _train_worker(rank, world, port, out):
    """Scalar exchange + running-mean replay + flat gradient all-reduce, against the sequential loop."""
    from ocn_b200.dist import allreduce_gradients, exchange_batch_scalars, replay_running_mean
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    U = 7
    s_seq = torch.randn(U).abs() * 3  # the inner products of the 7 sub-batches of a step
    mine = list(range(rank, U, world))
    s_all = exchange_batch_scalars(s_seq[mine], mine, U)
    ips, n_end = replay_running_mean(s_all, torch.tensor([0.25]), 3)
    # sequential module behaviour (model.py:2245-2248)
    ip, n, want = torch.tensor([0.25]), 3, []
    for u in range(U):
        n += 1
        beta = n ** -1
        ip *= (1 - beta)
        ip += beta * s_seq[u]
        want.append(ip.clone())
    ok = torch.equal(s_all, s_seq) and torch.equal(ips, torch.cat(want)) and n_end == n
    # gradients: every rank back-propagates its sub-batches' share of the loss; the sum is the step gradient
    lin = torch.nn.Linear(5, 3)
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    unused = torch.nn.Parameter(torch.ones(4))  # never touched on any rank: stays zero
    x = torch.randn(U, 11, 5)
    for u in mine:
        (lin(x[u]).sigmoid().sum() / U).backward()
    allreduce_gradients(list(lin.parameters()) + [frozen, unused])
    ref = torch.nn.Linear(5, 3)
    ref.load_state_dict(lin.state_dict())
    for u in range(U):
        (ref(x[u]).sigmoid().sum() / U).backward()
    ok = ok and torch.allclose(lin.weight.grad, ref.weight.grad, rtol=1e-5, atol=1e-6) \
        and torch.allclose(lin.bias.grad, ref.bias.grad, rtol=1e-5, atol=1e-6) \
        and frozen.grad is None and bool((unused.grad == 0).all())
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_training_step_exchange_and_gradient_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True


def _cost_worker(rank, world, port, out):
    """Slices of a stream dealt by predicted cost: every rank computes the dealing on its own replica of the graph
    (no collective), scores its slices batch by batch, and the one gather returns the stream in link order."""
    from ocn_b200 import synth
    from ocn_b200.dist import deal_by_cost, predicted_walk_cost
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = synth.tiny_graph(200, 1500, 9)
    slice_links, bs, per_rank = 96, 32, 3
    T = slice_links * world * per_rank
    edges = g.query_edges(T, "stream")
    cost = predicted_walk_cost(g.rowptr, g.col, edges[0], slice_links, bs).tolist()
    dealt = deal_by_cost(cost, world, per_rank)
    # the ranks agree without talking: compare through one all_gather of the flattened dealing
    mine_flat = torch.tensor([i for o in dealt for i in o])
    seen = [torch.empty_like(mine_flat) for _ in range(world)]
    dist.all_gather(seen, mine_flat)
    ok = all(torch.equal(s, mine_flat) for s in seen)
    owned = [(sl * slice_links + b, sl * slice_links + b + bs) for sl in dealt[rank] for b in range(0, slice_links, bs)]
    score = lambda b: b[0].float() * 2 + b[1].float().sum()      # batch-coupled stand-in for the CUDA path
    scores = gather_scores([score(edges[:, s:e]) for (s, e) in owned], owned, T)
    ref = torch.cat([score(edges[:, s:s + bs]) for s in range(0, T, bs)])
    ok = ok and torch.equal(scores, ref)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_slices_dealt_by_cost_and_scores_gathered():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cost_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True
