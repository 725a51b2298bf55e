import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import bench
import fcdgan_b200 as fb
from fcdgan_b200 import parallel as P, engine as E, _lib
local = P.init_from_env("nccl"); rank = dist.get_rank(); world = dist.get_world_size()
dev = torch.device("cuda", local)
B, C, H, W = 16, 13, 256, 256
netG = fb.Generator(C).to(dev).train()
P.broadcast_parameters([netG])
optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99))
crit = fb.losses._MaskedRecon; sync = P.GradSync()
zero_cmap = torch.zeros(B, 1, H, W, device=dev)
x, y, region, cmap = bench.synth(B, 1 + rank, device=dev)
calls = []
orig_call = _lib.call
def timed_call(name, *a):
    t = time.perf_counter(); r = orig_call(name, *a); calls.append((name, (time.perf_counter() - t) * 1e3)); return r
_lib.call = timed_call
orig_empty = torch.empty
def timed_empty(*a, **k):
    t = time.perf_counter(); r = orig_empty(*a, **k); calls.append(("torch.empty", (time.perf_counter() - t) * 1e3)); return r
torch.empty = timed_empty
def step(mode):
    y_fake = netG(x); gl, _, _, _ = crit.apply(y, y_fake, zero_cmap, 0, False)
    optG.zero_grad(set_to_none=True); gl.backward()
    if mode == 1:
        sync.start(netG); sync.finish()
    elif mode == 2:
        flat = torch.cat([p.grad.reshape(-1) for p in netG.parameters()]); dist.all_reduce(flat)
    optG.step()
for mode in (1, 2, 0):
    for _ in range(3): step(mode)
    torch.cuda.synchronize(); dist.barrier(); calls.clear()
    ms0 = torch.cuda.memory_stats()["num_device_alloc"]
    t0 = time.perf_counter()
    for _ in range(6): step(mode)
    host = (time.perf_counter() - t0) * 1e3 / 6
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3 / 6
    slow = sorted(calls, key=lambda c: -c[1])[:6]
    print(f"[rank {rank}] mode={mode} host {host:.1f} wall {wall:.1f} dev_allocs +{torch.cuda.memory_stats()['num_device_alloc'] - ms0} slowest {[(n, round(t, 1)) for n, t in slow]}", flush=True)
    dist.barrier()
dist.destroy_process_group()
