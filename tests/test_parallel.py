"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, parameter broadcast and the
bucketed asynchronous gradient all-reduce.  The same code runs over NCCL on the GPU box (bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcdgan_b200 import parallel as P


def test_shard_batch_covers_everything():
    for n in (1, 7, 16, 33):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                s = P.shard_batch(n, r, world)
                got += list(range(n))[s]
            assert got == list(range(n))
            sizes = [len(range(n)[P.shard_batch(n, r, world)]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)          # different initial weights per rank
    a = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    b = torch.nn.Linear(5, 2)
    P.broadcast_parameters([a, b])
    w_after_bcast = a[0].weight.detach().clone()
    # rank-dependent gradients
    for i, p in enumerate(list(a.parameters()) + list(b.parameters())):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    sync = P.GradSync()
    sync.start(a)
    sync.start(b)
    sync.finish()
    grads = [p.grad.clone() for p in list(a.parameters()) + list(b.parameters())]
    # the same exchange in its four phases (pack / launch / wait / unpack), the form bench.py replays from CUDA graphs at
    # N > 1; launch() must keep working after unpack() has run (a graph replay does not re-run pack())
    for i, p in enumerate(list(a.parameters()) + list(b.parameters())):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    sync.pack(a); sync.launch(a); sync.pack(b); sync.launch(b); sync.wait(); sync.unpack()
    for g, p in zip(grads, list(a.parameters()) + list(b.parameters())):
        assert torch.equal(g, p.grad)
    sync._buckets[id(a)].fill_(float(rank))            # "replay": the bucket was refilled on the device, pack() not called
    sync.launch(a); sync.wait()
    assert torch.allclose(sync._buckets[id(a)], torch.full_like(sync._buckets[id(a)], float(sum(range(world)))))
    gathered = [None] * world
    dist.all_gather_object(gathered, (w_after_bcast, grads))
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_two_ranks_gloo(tmp_path):
    world = 2
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    (w0, g0), (w1, g1) = res
    assert torch.equal(w0, w1)                                   # broadcast made the replicas identical
    for i, (a, b) in enumerate(zip(g0, g1)):
        assert torch.equal(a, b)
        assert torch.allclose(a, torch.full_like(a, 1.5 * (i + 1)))  # mean of (1, 2) * (i + 1)


def _step_gen(net, opt, x, y):
    """a step body in the shape of fcdgan_b200.steps.*_gen: yields (network, wait) where the gradients are complete"""
    loss = ((net(x) - y) ** 2).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    yield net, True
    opt.step()
    return loss


def _worker_step(rank, world, port, out):
    from fcdgan_b200.steps import drive

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 3, 10, 10, generator=g), torch.randn(8, 4, 8, 8, generator=g)
    torch.manual_seed(3)
    full = torch.nn.Conv2d(3, 4, 3)
    torch.manual_seed(50 + rank)
    mine = torch.nn.Conv2d(3, 4, 3)
    with torch.no_grad():
        if rank == 0:
            for a, b in zip(mine.parameters(), full.parameters()):
                a.copy_(b)
    P.broadcast_parameters([mine])
    opt_full, opt_mine = torch.optim.SGD(full.parameters(), lr=0.1), torch.optim.SGD(mine.parameters(), lr=0.1)
    sync = P.GradSync()
    sl = P.shard_batch(8, rank, world)
    for _ in range(3):
        drive(_step_gen(full, opt_full, x, y))                            # 1-rank full batch, no exchange
        drive(_step_gen(mine, opt_mine, x[sl], y[sl]), sync.on_grads)     # sharded, exchange at the yield
    err = max((a - b).abs().max().item() for a, b in zip(mine.parameters(), full.parameters()))
    gathered = [None] * world
    dist.all_gather_object(gathered, err)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_through_on_grads_equals_full_batch_gloo(tmp_path):
    """N-rank sharded optimizer steps driven through the step generators' exchange points (steps.drive + GradSync.on_grads)
    track the 1-rank full-batch steps (no BatchNorm in the toy net, so the equivalence is exact up to summation order)."""
    world = 2
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_step, args=(world, _free_port(), out), nprocs=world, join=True)
    errs = torch.load(out, weights_only=False)
    assert max(errs) < 1e-6, errs
