"""One 3x3 64->64 forward convolution with bias + fused BatchNorm statistics (B = 16, 256 x 256) per precision, for an
`ncu --set full --import-source on` capture of conv_halo_kernel:  python scripts/ncu_conv_fwd.py [fast|parity]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E, _lib
dev = torch.device("cuda:0")
N, C, H, W = 16, 64, 256, 256
prec = sys.argv[1] if len(sys.argv) > 1 else "fast"
fb.set_precision(prec)
tape = E.Tape(dev, False)
act = tape.new_act(N, H, W, C)
act.hi.normal_()
if act.lo is not None:
    act.lo.normal_().mul_(2 ** -9)
w = torch.randn(C, C, 3, 3, device=dev) * 0.05
w_hi, w_lo = E._packed(w, C, C, 0, "v")
bias = torch.randn(C, device=dev)
z = torch.zeros(N, H, W, C, device=dev)
st = torch.zeros(2, C, dtype=torch.float64, device=dev)
for _ in range(3):
    _lib.call("fcd_conv2d_fwd", act.p_hi(), act.p_lo(), act.ld, w_hi.data_ptr(), _lib.ptr(w_lo), bias.data_ptr(), None, C, z.data_ptr(), C,
              N, H, W, C, C, 3, 3, 1, 1, st[0].data_ptr(), st[1].data_ptr(), 0, E._raw_stream())
torch.cuda.synchronize()
print("done")
