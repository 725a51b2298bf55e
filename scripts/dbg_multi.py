import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import bench
import fcdgan_b200 as fb
from fcdgan_b200 import parallel as P
local = P.init_from_env("nccl"); rank = dist.get_rank(); world = dist.get_world_size()
dev = torch.device("cuda", local)
B, C, H, W = 16, 13, 256, 256
netG, netD = fb.Generator(C).to(dev).train(), fb.Discriminator_SRGAN_simple(C).to(dev).train()
P.broadcast_parameters([netG, netD])
optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99)); optD = torch.optim.RMSprop(netD.parameters(), lr=5e-5)
crit = fb.losses._MaskedRecon; sync = P.GradSync()
zero_cmap = torch.zeros(B, 1, H, W, device=dev)
x, y, region, cmap = bench.synth(B, 1 + rank, device=dev)
T = {}
def mark(k, t0):
    T.setdefault(k, []).append((time.perf_counter() - t0) * 1e3)
def step(use_sync=True):
    t = time.perf_counter(); y_fake = netG(x); gl, _, _, _ = crit.apply(y, y_fake, zero_cmap, 0, False); mark("G fwd", t)
    t = time.perf_counter(); optG.zero_grad(set_to_none=True); gl.backward(); mark("G bwd", t)
    t = time.perf_counter()
    if use_sync: sync.start(netG)
    mark("sync.start G", t)
    t = time.perf_counter(); xm, ym = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap); c = netD(xm, ym); yu = fb.soft_mask(y, cmap, other=x, region=region); nc = netD(xm, yu); dl = 1 + fb.mean(nc) - fb.mean(c); mark("D fwd", t)
    t = time.perf_counter(); optD.zero_grad(set_to_none=True); dl.backward(); mark("D bwd", t)
    t = time.perf_counter()
    if use_sync: sync.start(netD); sync.finish()
    mark("sync D+finish", t)
    t = time.perf_counter(); optG.step(); optD.step(); mark("opt", t)
for mode in (True, False):
    for _ in range(3): step(mode)
    torch.cuda.synchronize(); dist.barrier(); T.clear()
    t0 = time.perf_counter()
    for _ in range(8): step(mode)
    host = (time.perf_counter() - t0) * 1e3 / 8
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3 / 8
    print(f"[rank {rank}] sync={mode} host {host:.1f} ms/step wall {wall:.1f} | " + " | ".join(f"{k} {sum(v)/len(v):.1f} (max {max(v):.0f})" for k, v in T.items()), flush=True)
    dist.barrier()
dist.destroy_process_group()
