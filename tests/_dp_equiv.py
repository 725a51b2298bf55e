"""Runner of tests/test_dp_gpu.py (launched with torchrun, one rank per GPU, NCCL): an N-rank sharded iteration with
`parallel.GradSync` must give every rank the gradients of the 1-rank full-batch iteration.

`--sync-bn`: the networks run in TRAIN mode with `fcdgan_b200.set_sync_bn(True)` — every BatchNorm call all-reduces its batch sums,
so statistics, running statistics and gradients are those of the concatenated batch (tolerance 2e-4: the statistics' summation
order differs).  Without it:
With BatchNorm in eval mode (running statistics) samples are independent, the losses are batch means and the shards are equal,
so mean-over-ranks of the shard gradients IS the full-batch gradient; what is left is fp32 summation order (the weight-gradient
reductions use fp32 atomics; the all-reduce sums in another order than one big batch would): tolerance 2e-5 of each tensor's
maximum.  Checked for the eager exchange (`GradSync.on_grads`) and for the CUDA-graph form bench.py uses at N > 1
(`graph.YieldingStep`: graphs cut at the exchange points, NCCL calls between them)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb  # noqa: E402
from fcdgan_b200 import parallel as P  # noqa: E402
from fcdgan_b200.graph import YieldingStep  # noqa: E402
from fcdgan_b200.steps import drive  # noqa: E402

C, H, W, PER_RANK = 4, 64, 56, 2


SYNC_BN = "--sync-bn" in sys.argv      # train-mode BatchNorm with synchronised statistics instead of eval-mode BatchNorm


def make_nets(dev):
    torch.manual_seed(0)
    netG = fb.Generator(C).to(dev).train(SYNC_BN)
    torch.manual_seed(1)
    netD = fb.Discriminator_SRGAN_simple(C).to(dev).train(SYNC_BN)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():                      # non-trivial running statistics
        for net in (netG, netD):
            for k, v in net.state_dict().items():
                if "running_var" in k:
                    v.copy_((0.5 + torch.rand(v.shape, generator=g)).to(dev))
                elif "running_mean" in k:
                    v.copy_((0.2 * torch.randn(v.shape, generator=g)).to(dev))
    return netG, netD


def make_gen(netG, netD):
    def gen(x, y, cmap):
        zero = torch.zeros_like(cmap)
        y_fake = netG(x)
        gl, _, _, _ = fb.losses._MaskedRecon.apply(y, y_fake, zero, fb.losses.LOSS_L1, False)
        netG.zero_grad()
        gl.backward()
        yield netG, False
        c_out = netD(fb.soft_mask(x, cmap), fb.soft_mask(y, cmap))
        nc_out = netD(fb.soft_mask(x, cmap), fb.soft_mask(x, cmap))
        dl = 1 + fb.mean(nc_out) - fb.mean(c_out)
        netD.zero_grad()
        dl.backward()
        yield netD, True
        return gl, dl
    return gen


def grads(*nets):
    return [(k, p.grad.detach().clone()) for n in nets for k, p in n.named_parameters()]


def main():
    local = P.init_from_env("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    fb.set_precision("parity")
    B = PER_RANK * world
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, H, W, generator=g)
    y = x + 0.3 * torch.randn(B, C, H, W, generator=g)
    cmap = 0.3 * torch.rand(B, 1, H, W, generator=g)
    sl = P.shard_batch(B, rank, world)
    xs, ys, cs = (t[sl].to(dev) for t in (x, y, cmap))
    # 1-rank full batch (every rank computes it for itself; no communication)
    fb.set_sync_bn(False)
    netG, netD = make_nets(dev)
    # (losses are detached at once: a live loss keeps its iteration's autograd graph — and with it the AccumulateGrad nodes
    # created on THIS stream — alive, and a later CUDA-graph capture would re-use those nodes across streams; graph.py docstring)
    gl_full, dl_full = (v.detach() for v in drive(make_gen(netG, netD)(x.to(dev), y.to(dev), cmap.to(dev))))
    print(f"[rank {rank}] full-batch reference done", file=sys.stderr, flush=True)
    want = grads(netG, netD)
    want_stats = {k: v.clone() for n in (netG, netD) for k, v in n.state_dict().items() if "running" in k}
    # N-rank sharded, eager exchange
    fb.set_sync_bn(SYNC_BN)
    netG, netD = make_nets(dev)
    P.broadcast_parameters([netG, netD])
    sync = P.GradSync()
    gl, dl = (v.detach() for v in drive(make_gen(netG, netD)(xs, ys, cs), sync.on_grads))
    got_eager = grads(netG, netD)
    got_stats = {k: v.clone() for n in (netG, netD) for k, v in n.state_dict().items() if "running" in k}
    print(f"[rank {rank}] eager exchange done", file=sys.stderr, flush=True)
    # N-rank sharded, CUDA graphs cut at the exchange points (eval-mode BatchNorm only: SyncBN's all-reduces inside captured
    # segments mixed with eager all-reduces between them hung on this stack; the engine refuses SyncBN under capture)
    if SYNC_BN:
        got_graph = got_eager
    else:
        step = YieldingStep(make_gen(netG, netD), sync, [xs, ys, cs], warmup=2, modules=[netG, netD])
        assert len(step.graphs) == 3
        step()
        torch.cuda.synchronize()
        got_graph = grads(netG, netD)
    print(f"[rank {rank}] graph form done", file=sys.stderr, flush=True)
    if SYNC_BN:      # synchronised statistics == the statistics of the concatenated batch: the running statistics agree too
        for k, v in want_stats.items():
            err = (got_stats[k] - v).abs().max().item() / max(v.abs().max().item(), 1e-6)
            assert err < 1e-5, f"rank {rank} running statistic {k}: {err:.3g}"
    tol = 2e-4 if SYNC_BN else 2e-5      # train-mode BatchNorm amplifies the summation-order noise of its statistics
    worst = 0.0
    # scale of a tensor's error: its own maximum, but not less than 1 % of the largest gradient entry of the run — the last
    # layer's bias gradient is a 1-element sum of sigmoid'(nc) - sigmoid'(c) terms that nearly cancel (nc_out is D(x, x): its
    # feature difference is exactly 0), so its own magnitude says nothing about the accuracy of the sums behind it
    gmax = max(b.abs().max().item() for _, b in want)
    for what, got in (("eager", got_eager), ("graphs", got_graph)):
        for (k, a), (_, b) in zip(got, want):
            scale = max(b.abs().max().item(), 1e-2 * gmax)
            err = (a - b).abs().max().item() / scale
            worst = max(worst, err)
            assert err < tol, f"rank {rank} [{what}] {k}: {err:.3g} (|ref|max {b.abs().max().item():.3g}, run max {gmax:.3g})"
    losses = torch.stack([gl, dl])
    dist.all_reduce(losses)
    losses /= world
    assert abs(losses[0].item() - gl_full.item()) < 1e-6 * max(1, abs(gl_full.item())), (losses, gl_full)
    assert abs(losses[1].item() - dl_full.item()) < 1e-6 * max(1, abs(dl_full.item())), (losses, dl_full)
    dist.barrier()
    if rank == 0:
        print(f"DP_EQUIV_OK world={world} sync_bn={SYNC_BN} worst_rel_err={worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
