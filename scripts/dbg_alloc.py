import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fcdgan_b200 as fb
from fcdgan_b200 import parallel as P
use_dist = "RANK" in os.environ
if use_dist:
    import torch.distributed as dist
    P.init_from_env("nccl")
    t = torch.ones(1, device="cuda"); dist.all_reduce(t)
dev = torch.device("cuda", 0)
B, C, H, W = 16, 13, 256, 256
netG, netD = fb.Generator(C).to(dev).train(), fb.Discriminator_SRGAN_simple(C).to(dev).train()
optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99)); optD = torch.optim.RMSprop(netD.parameters(), lr=5e-5)
crit = fb.losses._MaskedRecon
zero_cmap = torch.zeros(B, 1, H, W, device=dev)
x, y, region, cmap = bench.synth(B, 1, device=dev)
def step():
    y_fake = netG(x); gl, _, _, _ = crit.apply(y, y_fake, zero_cmap, 0, False)
    optG.zero_grad(set_to_none=True); gl.backward()
    xm, ym = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap); c = netD(xm, ym); yu = fb.soft_mask(y, cmap, other=x, region=region); nc = netD(xm, yu)
    dl = 1 + fb.mean(nc) - fb.mean(c); optD.zero_grad(set_to_none=True); dl.backward(); optG.step(); optD.step()
rows = []
for i in range(24):
    a0 = torch.cuda.memory_stats()["num_device_alloc"]; t0 = time.perf_counter()
    step()
    host = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    rows.append((i, torch.cuda.memory_stats()["num_device_alloc"] - a0, round(host, 1), round((time.perf_counter() - t0) * 1e3, 1)))
print("dist" if use_dist else "plain", "step, new cudaMallocs, host ms, wall ms:", rows, "reserved GiB", round(torch.cuda.memory_reserved() / 2**30, 1), flush=True)
if use_dist:
    dist.destroy_process_group()
