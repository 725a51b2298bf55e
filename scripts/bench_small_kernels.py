"""HBM-bound glue kernels at production shapes, timed alone with CUDA events (inputs far larger than L2), against the bytes
each one has to move: the staging kernels either side of the few-band layers, the soft mask, OutConv, max-pool and bilinear
up-sampling.  `python scripts/bench_small_kernels.py [B]` (default B = 16 tiles of 13 x 256 x 256)."""
import sys

import torch

sys.path.insert(0, ".")
from fcdgan_b200 import _lib  # noqa: E402

DEV = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, nbytes, name, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:58s} {us:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / us / 1e3:7.0f} GB/s", flush=True)


def bf(*shape):
    return torch.empty(shape, dtype=torch.bfloat16, device=DEV)


C, H, W = 13, 256, 256
x = torch.randn(B, C, H, W, device=DEV)
# staging
M, P, Kp = 4, 9, 128
hi, lo = bf(B, H, W + M, Kp), bf(B, H, W + M, Kp)
timeit(lambda: _lib.call("fcd_stage_nchw_to_split_rowpack", x.data_ptr(), B, C, H, W, M, P, Kp, hi.data_ptr(), lo.data_ptr(), S()),
       x.numel() * 4 + hi.numel() * 4, "stage rowpack 13 bands x 9 px -> K=128")
h4, l4 = bf(B, H, W + M, 64), bf(B, H, W + M, 64)
timeit(lambda: _lib.call("fcd_stage_nchw_to_split_pack4", x.data_ptr(), B, C, H, W, M, h4.data_ptr(), l4.data_ptr(), S()),
       x.numel() * 4 + h4.numel() * 4, "stage pack4 13 bands x 4 px -> K=64")
hi2, lo2 = bf(B, H // 2, W // 2, 128), bf(B, H // 2, W // 2, 128)
timeit(lambda: _lib.call("fcd_stage_im2col3x3s2", x.data_ptr(), B, C, H, W, hi2.data_ptr(), lo2.data_ptr(), 128, S()),
       x.numel() * 4 + hi2.numel() * 4, "stage im2col 3x3 s2 13 bands -> K=128")
hs, ls = bf(B, H, W, 64), bf(B, H, W, 64)
timeit(lambda: _lib.call("fcd_stage_nchw_to_split", x.data_ptr(), None, B, C, H, W, hs.data_ptr(), ls.data_ptr(), 64, 64, S()),
       x.numel() * 4 + hs.numel() * 4, "stage NCHW -> split NHWC (64 slots)")
# soft mask
m = torch.rand(B, 1, H, W, device=DEV)
y = torch.randn_like(x)
out = torch.empty_like(x)
timeit(lambda: _lib.call("fcd_mask_fwd", x.data_ptr(), None, None, m.data_ptr(), B, C, H, W, out.data_ptr(), S()),
       x.numel() * 8 + m.numel() * 4, "soft mask x*(1-m)")
timeit(lambda: _lib.call("fcd_mask_fwd", x.data_ptr(), y.data_ptr(), m.data_ptr(), m.data_ptr(), B, C, H, W, out.data_ptr(), S()),
       x.numel() * 12 + m.numel() * 8, "soft mask (x*(1-r)+y*r)*(1-m)")
# OutConv 64 -> 1
Cin = 64
a_hi, a_lo = torch.randn(B, H, W, Cin, device=DEV).bfloat16(), (torch.randn(B, H, W, Cin, device=DEV) * 1e-3).bfloat16()
w, b = torch.randn(1, Cin, device=DEV) * 0.1, torch.zeros(1, device=DEV)
o = torch.empty(B, 1, H, W, device=DEV)
timeit(lambda: _lib.call("fcd_outconv_sigmoid_fwd", a_hi.data_ptr(), a_lo.data_ptr(), Cin, Cin, w.data_ptr(), b.data_ptr(), 1, B, H, W,
                         o.data_ptr(), S()), a_hi.numel() * 4 + o.numel() * 4, "OutConv 64->1 + sigmoid forward")
do = torch.randn_like(o)
dx = torch.empty(B, H, W, Cin, device=DEV)
dw, db = torch.empty(1, Cin, device=DEV), torch.empty(1, device=DEV)
scratch = torch.empty(Cin + 1, dtype=torch.float64, device=DEV)
timeit(lambda: _lib.call("fcd_outconv_sigmoid_bwd", do.data_ptr(), o.data_ptr(), a_hi.data_ptr(), a_lo.data_ptr(), Cin, Cin, w.data_ptr(),
                         1, B, H, W, dx.data_ptr(), Cin, dw.data_ptr(), db.data_ptr(), 0, scratch.data_ptr(), S()),
       a_hi.numel() * 4 + dx.numel() * 4 + o.numel() * 8, "OutConv backward")
# max-pool / up-sampling on a 64-channel 256^2 tensor
p_hi, p_lo = bf(B, H // 2, W // 2, Cin), bf(B, H // 2, W // 2, Cin)
timeit(lambda: _lib.call("fcd_maxpool2_fwd", a_hi.data_ptr(), a_lo.data_ptr(), Cin, B, H, W, Cin, p_hi.data_ptr(), p_lo.data_ptr(), Cin, S()),
       a_hi.numel() * 4 + p_hi.numel() * 4, "maxpool2 forward 64 ch")
dpo = torch.randn(B, H // 2, W // 2, Cin, device=DEV)
timeit(lambda: _lib.call("fcd_maxpool2_bwd", dpo.data_ptr(), Cin, a_hi.data_ptr(), a_lo.data_ptr(), Cin, B, H, W, Cin, dx.data_ptr(), Cin, 0, S()),
       a_hi.numel() * 4 + dx.numel() * 4 + dpo.numel() * 4, "maxpool2 backward 64 ch")
timeit(lambda: _lib.call("fcd_upsample2x_bilinear_fwd", p_hi.data_ptr(), p_lo.data_ptr(), Cin, B, H // 2, W // 2, Cin, a_hi.data_ptr(),
                         a_lo.data_ptr(), Cin, H, W, 0, 0, S()), a_hi.numel() * 4 + p_hi.numel() * 4, "bilinear x2 forward 64 ch")
timeit(lambda: _lib.call("fcd_upsample2x_bilinear_bwd", dx.data_ptr(), Cin, B, H // 2, W // 2, Cin, H, W, 0, 0, dpo.data_ptr(), Cin, S()),
       dx.numel() * 4 + dpo.numel() * 4, "bilinear x2 backward 64 ch")
