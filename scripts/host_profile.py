"""cProfile of the host side of one bench step (where does the CPU time of issuing a step go?)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E, parallel as P
dev = torch.device("cuda:0")
B, C, H, W = 16, 13, 256, 256
netG, netD = fb.Generator(C).to(dev).train(), fb.Discriminator_SRGAN_simple(C).to(dev).train()
optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99))
optD = torch.optim.RMSprop(netD.parameters(), lr=5e-5)
crit = fb.losses._MaskedRecon
zero_cmap = torch.zeros(B, 1, H, W, device=dev)
x, y, region, cmap = bench.synth(B, 1, device=dev)
def step():
    y_fake = netG(x)
    gen_loss, _, _, _ = crit.apply(y, y_fake, zero_cmap, 0, False)
    optG.zero_grad(set_to_none=True)
    gen_loss.backward()
    x_mask, y_mask = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap)
    c_out = netD(x_mask, y_mask)
    y_unc = fb.soft_mask(y, cmap, other=x, region=region)
    nc_out = netD(x_mask, y_unc)
    d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
    optD.zero_grad(set_to_none=True)
    d_loss.backward()
    optG.step(); optD.step()
for _ in range(3): step()
torch.cuda.synchronize()
import time
t=time.perf_counter()
for _ in range(5): step()
print("host issue ms/step", (time.perf_counter()-t)*1e3/5)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter()-t)*1e3/5)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])

# ---- which torch.empty calls are slow?
import time as _t
_orig = torch.empty
log = []
def timed_empty(*a, **k):
    t0 = _t.perf_counter(); r = _orig(*a, **k); dt = (_t.perf_counter() - t0) * 1e6
    log.append((dt, r.numel() * r.element_size()))
    return r
torch.empty = timed_empty
ms0 = torch.cuda.memory_stats()
for _ in range(3): step()
torch.cuda.synchronize()
ms1 = torch.cuda.memory_stats()
torch.empty = _orig
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated", "segment.all.allocated"):
    print(k, ms0.get(k), "->", ms1.get(k))
print("reserved GiB", torch.cuda.memory_reserved() / 2**30, "peak allocated GiB", torch.cuda.max_memory_allocated() / 2**30)
log.sort(reverse=True)
print("slowest:", [(round(d), s) for d, s in log[:15]])
print("total us", sum(d for d, _ in log), "n", len(log), "n>50us", sum(1 for d, _ in log if d > 50))
