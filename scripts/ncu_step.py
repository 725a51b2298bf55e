"""One eager iteration of the bench.py workload (config 2: G + D fwd/bwd + optimizer steps, batch 16 of 13 x 256 x 256) between
cudaProfilerStart/Stop, for the ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file launches.csv python scripts/ncu_step.py
(bench.py replays the same launches from one CUDA graph; the list is taken in eager mode so every kernel is visible)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fcdgan_b200 as fb

dev = torch.device("cuda:0")
fb.set_precision(sys.argv[1] if len(sys.argv) > 1 else "parity")
B, C = bench.BATCH_PER_GPU, bench.C
torch.manual_seed(0)
netG, netD = fb.Generator(C).to(dev).train(), fb.Discriminator_SRGAN_simple(C).to(dev).train()
optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99))
optD = torch.optim.RMSprop(netD.parameters(), lr=5e-5)
zero_cmap = torch.zeros(B, 1, bench.H, bench.W, device=dev)
x, y, region, cmap = bench.synth(B, 1234, device=dev)


def step():
    y_fake = netG(x)
    gen_loss, _, _, _ = fb.losses._MaskedRecon.apply(y, y_fake, zero_cmap, fb.losses.LOSS_L1, False)
    optG.zero_grad(set_to_none=True)
    gen_loss.backward()
    x_mask, y_mask = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap)
    c_out = netD(x_mask, y_mask)
    nc_out = netD(x_mask, fb.soft_mask(y, cmap, other=x, region=region))
    d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
    optD.zero_grad(set_to_none=True)
    d_loss.backward()
    optG.step()
    optD.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
