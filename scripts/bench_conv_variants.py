"""Times fcd_conv2d_fwd (3x3, 64 -> 64, B = 16, 256 x 256) with the epilogue features switched on one at a time:
bias, BatchNorm statistics, in-place accumulation (reduce-add), plus the TMA-store epilogue on/off."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E, _lib
dev = torch.device("cuda:0")
N, C, H, W = 16, 64, 256, 256
FL = 2.0 * N * H * W * C * C * 9


def run(prec):
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

    def call(use_bias, use_stats, use_addend):
        _lib.call("fcd_conv2d_fwd", act.p_hi(), act.p_lo(), act.ld, w_hi.data_ptr(), _lib.ptr(w_lo),
                  bias.data_ptr() if use_bias else None, z.data_ptr() if use_addend else None, C, z.data_ptr(), C, N, H, W, C, C,
                  3, 3, 1, 1, st[0].data_ptr() if use_stats else None, st[1].data_ptr() if use_stats else None, 0, E._raw_stream())

    for tma, slots in ((1, 0), (1, 2), (0, 0)):
        _lib.call("fcd_set_option", b"conv_tma_out", tma)
        _lib.call("fcd_set_option", b"conv_halo_slots", slots)
        for name, args in (("plain", (0, 0, 0)), ("bias", (1, 0, 0)), ("bias+stats", (1, 1, 0)), ("stats", (0, 1, 0)),
                           ("accumulate (dgrad form)", (0, 0, 1))):
            for _ in range(3):
                call(*args)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call(*args)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"{prec:6s} tma_out={tma} x_slots={slots or 'auto'}  {name:26s} {ms * 1e3:7.1f} us   {FL / ms / 1e9:7.1f} TFLOP/s algorithmic", flush=True)
    _lib.call("fcd_set_option", b"conv_tma_out", 1)
    _lib.call("fcd_set_option", b"conv_halo_slots", 0)


for prec in ("parity", "fast"):
    run(prec)
