"""One optimiser step of the citation2 predictor loop (NeighborOverlapCitation2.py:131-209: 16 384 positive + 16 384
negative links in sub-batches of 2048, cn5 order 2) over NCCL, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/train_step_nccl.py

Every rank runs ``ocn_b200.dist.sharded_train_step`` on its share of the sub-batches (scalar exchange of the inner
products, in-place all-reduce of the 375 MB gradient of h, one flat bucket for the predictor); rank 0 also runs the whole
step alone first, and the summed gradients / loss / inner-product buffer are compared.  Timing: CUDA events, max over
ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocn_b200 as ob
from ocn_b200 import synth
from ocn_b200.dist import sharded_train_step

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
DEV = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(DEV))
g = synth.make_graph("citation2", device=DEV)
G = ob.Graph(g.rowptr, g.col, g.n)
pos = g.query_edges(16384, "pos", device=DEV)
neg = torch.stack((pos[0], synth.hash_randint(16384, g.n, 5, 9, DEV)))
subs = [pos[:, k:k + 2048].contiguous() for k in range(0, 16384, 2048)] + [neg[:, k:k + 2048].contiguous() for k in range(0, 16384, 2048)]
signs = [1.0] * 8 + [-1.0] * 8


def fresh():
    torch.manual_seed(0)
    pred = ob.CNLinkPredictorOringin(32, 32, 1, 3, 0.0, weighted=True).to(DEV).train()
    h = g.features(32, device=DEV).requires_grad_(True)
    return pred, h


def step(pred, h, r, w, reps=4):
    ms = []
    for rep in range(reps):
        pred.zero_grad(set_to_none=True)
        h.grad = None
        n0, ip0 = pred.n, pred.innerprod.clone()
        if w > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = sharded_train_step(pred, h, G, subs, signs, 16384, r, w)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
        if rep < reps - 1:                      # every repetition starts from the same state
            pred.n = n0
            with torch.no_grad():
                pred.innerprod.copy_(ip0)
    return loss, min(ms[1:])


pred1, h1 = fresh()
# single-process reference of the same step (world = 1 code path), every rank computes it: nothing is exchanged
import ocn_b200.dist as D
real_world = D._world
D._world = lambda: 1
loss1, ms1 = step(pred1, h1, 0, 1)
D._world = real_world
pred, h = fresh()
loss, ms = step(pred, h, rank, world)
t = torch.tensor([ms], device=DEV)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
err_h = ((h.grad - h1.grad).abs().max() / (h1.grad.abs().max() + 1e-30)).item()
err_p = max(((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-30)).item()
            for p, q in zip(pred.parameters(), pred1.parameters()) if q.grad is not None)
ok = err_h < 1e-4 and err_p < 1e-3 and abs(float(loss) - float(loss1)) < 1e-5 * (1 + abs(float(loss1))) and pred.n == pred1.n
if rank == 0:
    print(json.dumps({"what": "citation2 optimiser step, cn5 order 2, 32768 links in 16 sub-batches, forward + backward, "
                              "sharded over ranks (NCCL)", "n_gpus": world, "ms_sharded_max_over_ranks": float(t),
                      "ms_one_gpu": ms1, "loss": float(loss), "loss_one_gpu": float(loss1), "rel_err_grad_h": err_h,
                      "rel_err_grad_params": err_p, "innerprod": float(pred.innerprod), "innerprod_one_gpu": float(pred1.innerprod),
                      "equal": bool(ok)}))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
