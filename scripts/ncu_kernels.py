"""One launch of every kernel family at Generator-layer size (B=16, 256x256, 64 channels) for an `ncu --set full`
capture: conv fwd / dgrad / wgrad (tcgen05), BN stats / apply / backward, staging, masked loss, SSIM level."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E, _lib
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "parity"
fb.set_precision(prec)
N, C, H, W = 16, 64, 256, 256
torch.manual_seed(0)
tape = E.Tape(dev, True)
x = torch.randn(N, 13, H, W, device=dev)
a = E.stage_input(tape, x, False)                                   # stage (13 -> 64 channels)
xp = E.PackedAct(x)                                                 # pack4 staging
w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
b = torch.zeros(64, device=dev)
act = tape.new_act(N, H, W, 64)
act.hi.normal_(); 
if act.lo is not None: act.lo.zero_()
for rep in range(2):                                                 # first pass warms up (attribute setup, packing)
    tape = E.Tape(dev, True); tape.track(act)
    z = E.conv(tape, act, w, b, 1, 1, stats=True)                    # conv_tc fwd + bn_stats
    bn = E.BN(torch.ones(64, device=dev), torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.ones(64, device=dev), None)
    out = E.bn_act(tape, z, bn, True, E.ACT_PRELU, slope=torch.full((1,), 0.25, device=dev))   # finalize + bn_act_fwd
    out.grad.normal_(); out.mark_ready()
    tape.ops[-1](tape)                                               # bwd reduce / finalize / apply
    tape.ops[-2](tape)                                               # wgrad_tc + dgrad (conv_tc)
# losses at boundary size
t = torch.randn(N, 13, H, W, device=dev); g = (t + 0.1 * torch.randn_like(t)).requires_grad_(True)
cm = torch.rand(N, 1, H, W, device=dev).requires_grad_(True)
crit = fb.CNetLoss(channel=13)
for rep in range(2):
    gl, l1, _, sl = crit(t, g, cm)
    (gl + l1 + sl).backward()
torch.cuda.synchronize()
print("done")
