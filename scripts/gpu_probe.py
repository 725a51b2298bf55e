"""GPU bring-up probe (run under gpurun): pins tcgen05 descriptor semantics and checks the convolution
engines (tcgen05 + SIMT) against an fp64 torch reference on the same split-bf16 operands.

    gpurun -- python scripts/gpu_probe.py [--skip-umma] [--time]
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcdgan_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def joined(hi, lo):
    return hi.double() + lo.double()


# ------------------------------------------------------------------------------------------------
def umma_probe():
    print("=== UMMA descriptor probe ===", flush=True)
    g = torch.Generator(device="cpu").manual_seed(7)

    def run(mn_major, n, ksteps, a_rows, b_rows, a_blocks, b_blocks, a_shift=0, a_bo=0, b_shift=0, b_bo=0,
            a_sbo=0, b_sbo=0):
        A = (torch.randint(-4, 5, (a_blocks, a_rows, 64), generator=g).float()).to(dev)
        B = (torch.randint(-4, 5, (b_blocks, b_rows, 64), generator=g).float()).to(dev)
        Ab, Bb = A.to(torch.bfloat16).contiguous(), B.to(torch.bfloat16).contiguous()
        D = torch.full((128, n), float("nan"), device=dev)
        _lib.call_probe("fcd_debug_umma_probe", Ab.data_ptr(), Bb.data_ptr(), D.data_ptr(), a_rows, b_rows, a_blocks,
                  b_blocks, mn_major, n, ksteps, a_shift, a_bo, b_shift, b_bo, a_sbo, b_sbo, 0, 0, None)
        torch.cuda.synchronize()
        return A.double(), B.double(), D.double()

    # 1. K-major sanity: D[m][n] = sum_k A[m][k] B[n][k], K = 64
    A, B, D = run(0, 64, 4, 128, 64, 1, 1)
    ref = A[0] @ B[0].t()
    print(f"K-major  shift=0            max|err| = {(D - ref).abs().max().item():.3g}", flush=True)
    # 2. K-major with row-shifted A start (halo view): rows m+shift
    for shift in (1, 2, 3, 5, 7, 8, 9):
        for bo in sorted({0, shift & 7}):
            A, B, D = run(0, 64, 4, 144, 64, 1, 1, a_shift=shift, a_bo=bo)
            ref = A[0][shift:shift + 128] @ B[0].t()
            err = (D - ref).abs().max().item()
            print(f"K-major  A shift={shift} base_off={bo}  max|err| = {err:.3g}", flush=True)
    # 2b. B shifted as well
    A, B, D = run(0, 64, 4, 128, 80, 1, 1, b_shift=3, b_bo=0)
    ref = A[0] @ B[0][3:67].t()
    print(f"K-major  B shift=3 base_off=0  max|err| = {(D - ref).abs().max().item():.3g}", flush=True)
    # 3. MN-major sanity: A blocks [k rows][64 m], M = 128 from two blocks; B [k rows][64 n]
    A, B, D = run(1, 64, 4, 64, 64, 2, 1)
    Am = torch.cat([A[0], A[1]], dim=1)  # [k][128]
    ref = Am.t() @ B[0]
    print(f"MN-major shift=0            max|err| = {(D - ref).abs().max().item():.3g}", flush=True)
    # 3b. MN-major, N = 128 from two B blocks
    A, B, D = run(1, 128, 4, 64, 64, 2, 2)
    Am = torch.cat([A[0], A[1]], dim=1)
    Bm = torch.cat([B[0], B[1]], dim=1)
    ref = Am.t() @ Bm
    print(f"MN-major N=128              max|err| = {(D - ref).abs().max().item():.3g}", flush=True)
    # 4. MN-major with row (K) shift
    for shift in (1, 3, 8, 9):
        for bo in sorted({0, shift & 7}):
            A, B, D = run(1, 64, 4, 80, 64, 2, 1, a_shift=shift, a_bo=bo)
            Am = torch.cat([A[0], A[1]], dim=1)[shift:shift + 64]
            ref = Am.t() @ B[0]
            print(f"MN-major A shift={shift} base_off={bo}  max|err| = {(D - ref).abs().max().item():.3g}", flush=True)


# ------------------------------------------------------------------------------------------------
def pack(w, Cout_p, Cin_p, mode, want_lo=True):
    Cout, Cin, KH, KW = w.shape
    rows, cols = (Cout_p, Cin_p) if mode == 0 else (Cin_p, Cout_p)
    hi = torch.empty(KH * KW, rows, cols, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi) if want_lo else None
    _lib.call("fcd_pack_conv_weight", w.contiguous().data_ptr(), Cout, Cin, KH, KW, Cout_p, Cin_p, mode,
              hi.data_ptr(), _lib.ptr(lo), None)
    return hi, lo


def ref_weight(w):
    hi, lo = split(w)
    return joined(hi, lo)


def conv_case(N, H, W, Cin, Cout, K, stride, pad, engine, fast=False, stats=False, tag=""):
    torch.manual_seed(0)
    Cin_p = (Cin + 15) // 16 * 16
    Cout_p = (Cout + 15) // 16 * 16
    x = torch.randn(N, H, W, Cin_p, device=dev)
    x[..., Cin:] = 0
    w = torch.randn(Cout, Cin, K, K, device=dev) * 0.1
    b = torch.randn(Cout_p, device=dev)
    b[Cout:] = 0
    xh, xl = split(x)
    wh, wl = pack(w, Cout_p, Cin_p, 0)
    OH = (H + 2 * pad - K) // stride + 1
    OW = (W + 2 * pad - K) // stride + 1
    z = torch.full((N, OH, OW, Cout_p), float("nan"), device=dev)
    ssum = torch.zeros(Cout_p, dtype=torch.float64, device=dev) if stats else None
    ssq = torch.zeros(Cout_p, dtype=torch.float64, device=dev) if stats else None
    _lib.call("fcd_conv2d_fwd", xh.data_ptr(), None if fast else xl.data_ptr(), Cin_p, wh.data_ptr(),
              None if fast else wl.data_ptr(), b.data_ptr(), z.data_ptr(), Cout_p, N, H, W, Cin_p, Cout_p, K, K,
              stride, pad, _lib.ptr(ssum), _lib.ptr(ssq), engine, None)
    torch.cuda.synchronize()
    xr = (xh.double() if fast else joined(xh, xl))[..., :Cin].permute(0, 3, 1, 2)
    wr = w.to(torch.bfloat16).double() if fast else ref_weight(w)
    ref = F.conv2d(xr, wr, b[:Cout].double(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    got = z[..., :Cout].double()
    err = (got - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    msg = f"conv{tag} eng={engine} fast={int(fast)} N{N} {H}x{W} {Cin}->{Cout} k{K}s{stride}p{pad}: max|err|={err:.3g} rel={rel:.3g}"
    if stats:
        s_ref = ref.sum(dim=(0, 1, 2))
        q_ref = (ref * ref).sum(dim=(0, 1, 2))
        msg += f" stat_err={((ssum[:Cout] - s_ref).abs().max() / s_ref.abs().max()).item():.2g}/{((ssq[:Cout] - q_ref).abs().max() / q_ref.abs().max()).item():.2g}"
    print(msg, flush=True)
    return rel


def wgrad_case(N, H, W, Cin, Cout, K, stride, pad, engine, tag=""):
    torch.manual_seed(1)
    Cin_p = (Cin + 15) // 16 * 16
    Cout_p = (Cout + 15) // 16 * 16
    OH = (H + 2 * pad - K) // stride + 1
    OW = (W + 2 * pad - K) // stride + 1
    x = torch.randn(N, H, W, Cin_p, device=dev)
    x[..., Cin:] = 0
    dz = torch.randn(N, OH, OW, Cout_p, device=dev)
    dz[..., Cout:] = 0
    xh, xl = split(x)
    gh, gl = split(dz)
    dw = torch.full((Cout, Cin, K, K), float("nan"), device=dev)
    db = torch.full((Cout,), float("nan"), device=dev)
    nbytes = _lib.load().fcd_conv2d_wgrad_workspace(N, H, W, Cin_p, Cout_p, K, K, stride, pad, engine)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    _lib.call("fcd_conv2d_wgrad", xh.data_ptr(), xl.data_ptr(), Cin_p, gh.data_ptr(), gl.data_ptr(), Cout_p,
              dw.data_ptr(), db.data_ptr(), N, H, W, Cin, Cin_p, Cout, Cout_p, K, K, stride, pad, 0, ws.data_ptr(),
              nbytes, engine, None)
    torch.cuda.synchronize()
    xr = joined(xh, xl)[..., :Cin].permute(0, 3, 1, 2).requires_grad_(False)
    gr = joined(gh, gl)[..., :Cout].permute(0, 3, 1, 2)
    wref = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, device=dev, requires_grad=True)
    out = F.conv2d(xr, wref, None, stride=stride, padding=pad)
    (gw,) = torch.autograd.grad(out, wref, gr)
    err = (dw.double() - gw).abs().max().item()
    rel = err / gw.abs().max().item()
    dberr = (db.double() - gr.sum(dim=(0, 2, 3))).abs().max().item()
    print(f"wgrad{tag} eng={engine} N{N} {H}x{W} {Cin}->{Cout} k{K}s{stride}p{pad}: max|err|={err:.3g} rel={rel:.3g} db_err={dberr:.3g}",
          flush=True)


def dgrad_strided_case(N, H, W, Cin, Cout, K, stride, pad):
    torch.manual_seed(2)
    Cin_p = (Cin + 15) // 16 * 16
    Cout_p = (Cout + 15) // 16 * 16
    OH = (H + 2 * pad - K) // stride + 1
    OW = (W + 2 * pad - K) // stride + 1
    w = torch.randn(Cout, Cin, K, K, device=dev) * 0.1
    dz = torch.randn(N, OH, OW, Cout_p, device=dev)
    dz[..., Cout:] = 0
    gh, gl = split(dz)
    wh, wl = pack(w, Cout_p, Cin_p, 0)
    dx = torch.full((N, H, W, Cin_p), float("nan"), device=dev)
    _lib.call("fcd_conv2d_dgrad_strided", gh.data_ptr(), gl.data_ptr(), Cout_p, wh.data_ptr(), wl.data_ptr(),
              dx.data_ptr(), Cin_p, N, H, W, Cin_p, Cout_p, K, K, stride, pad, None)
    torch.cuda.synchronize()
    xr = torch.zeros(N, Cin, H, W, dtype=torch.float64, device=dev, requires_grad=True)
    out = F.conv2d(xr, ref_weight(w), None, stride=stride, padding=pad)
    (gx,) = torch.autograd.grad(out, xr, joined(gh, gl)[..., :Cout].permute(0, 3, 1, 2))
    err = (dx[..., :Cin].double() - gx.permute(0, 2, 3, 1)).abs().max().item()
    print(f"dgrad_strided N{N} {H}x{W} {Cin}->{Cout} k{K}s{stride}p{pad}: max|err|={err:.3g} rel={err / gx.abs().max().item():.3g}",
          flush=True)


def dgrad_s1_case(N, H, W, Cin, Cout, K, pad, engine):
    """stride-1 dgrad = forward conv of dz with the mode-1 packed weights."""
    torch.manual_seed(3)
    w = torch.randn(Cout, Cin, K, K, device=dev) * 0.1
    dz = torch.randn(N, H, W, Cout, device=dev)
    gh, gl = split(dz)
    wh, wl = pack(w, Cout, Cin, 1)
    dx = torch.full((N, H, W, Cin), float("nan"), device=dev)
    _lib.call("fcd_conv2d_fwd", gh.data_ptr(), gl.data_ptr(), Cout, wh.data_ptr(), wl.data_ptr(), None, dx.data_ptr(),
              Cin, N, H, W, Cout, Cin, K, K, 1, K - 1 - pad, None, None, engine, None)
    torch.cuda.synchronize()
    xr = torch.zeros(N, Cin, H, W, dtype=torch.float64, device=dev, requires_grad=True)
    out = F.conv2d(xr, ref_weight(w), None, stride=1, padding=pad)
    (gx,) = torch.autograd.grad(out, xr, joined(gh, gl).permute(0, 3, 1, 2))
    err = (dx.double() - gx.permute(0, 2, 3, 1)).abs().max().item()
    print(f"dgrad_s1 eng={engine} N{N} {H}x{W} {Cin}->{Cout} k{K}p{pad}: max|err|={err:.3g} rel={err / gx.abs().max().item():.3g}",
          flush=True)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timing():
    print("=== timing (G-sized layer: N=16, 256x256, 64->64, 3x3) ===", flush=True)
    N, H, W, C = 16, 256, 256, 64
    x = torch.randn(N, H, W, C, device=dev)
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    xh, xl = split(x)
    wh, wl = pack(w, C, C, 0)
    z = torch.empty(N, H, W, C, device=dev)
    flops = 2.0 * N * H * W * C * C * 9
    ssum = torch.zeros(C, dtype=torch.float64, device=dev)
    ssq = torch.zeros(C, dtype=torch.float64, device=dev)
    for name, lo, eng, st in (("tc split", True, 2, False), ("tc split+stats", True, 2, True), ("tc fast", False, 2, False),
                              ("simt", True, 1, False)):
        def f():
            _lib.call("fcd_conv2d_fwd", xh.data_ptr(), xl.data_ptr() if lo else None, C, wh.data_ptr(),
                      wl.data_ptr() if lo else None, None, z.data_ptr(), C, N, H, W, C, C, 3, 3, 1, 1,
                      ssum.data_ptr() if st else None, ssq.data_ptr() if st else None, eng, None)
        ms = timeit(f)
        print(f"fwd {name:16s}: {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
    gh, gl = split(torch.randn(N, H, W, C, device=dev))
    dw = torch.empty(C, C, 3, 3, device=dev)
    for name, eng in (("tc", 2), ("simt", 1)):
        nbytes = _lib.load().fcd_conv2d_wgrad_workspace(N, H, W, C, C, 3, 3, 1, 1, eng)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)

        def f():
            _lib.call("fcd_conv2d_wgrad", xh.data_ptr(), xl.data_ptr(), C, gh.data_ptr(), gl.data_ptr(), C,
                      dw.data_ptr(), None, N, H, W, C, C, C, C, 3, 3, 1, 1, 0, ws.data_ptr(), nbytes, eng, None)
        ms = timeit(f, 3)
        print(f"wgrad {name:14s}: {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
    # cuDNN bar on the same box (fp32 / TF32), NCHW and channels_last
    xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        ms = timeit(lambda: F.conv2d(xc, w, None, padding=1))
        print(f"cudnn fwd tf32={int(tf32)} : {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
    xb = xc.to(torch.bfloat16)
    wb = w.to(torch.bfloat16)
    ms = timeit(lambda: F.conv2d(xb, wb, None, padding=1))
    print(f"cudnn fwd bf16   : {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


def sec_simt():
    print("=== SIMT conv ===", flush=True)
    conv_case(2, 20, 24, 13, 64, 9, 1, 4, 1, tag="[G head]")
    conv_case(2, 20, 24, 64, 13, 9, 1, 4, 1, tag="[G tail]")
    conv_case(2, 27, 27, 64, 64, 3, 1, 1, 1, stats=True)
    conv_case(2, 32, 32, 13, 64, 3, 2, 1, 1, tag="[D0]")
    conv_case(2, 16, 16, 128, 1, 1, 1, 0, 1, tag="[1x1]")
    dgrad_strided_case(2, 32, 32, 13, 64, 3, 2, 1)
    dgrad_strided_case(2, 27, 27, 64, 128, 3, 2, 1)
    dgrad_s1_case(2, 20, 24, 64, 64, 3, 1, 1)
    wgrad_case(2, 20, 24, 13, 64, 9, 1, 4, 1)
    wgrad_case(2, 27, 27, 64, 128, 3, 2, 1, 1)
    wgrad_case(2, 32, 48, 64, 64, 3, 1, 1, 1)


def sec_tc():
    print("=== tcgen05 conv ===", flush=True)
    conv_case(2, 32, 48, 64, 64, 3, 1, 1, 2, fast=True)
    conv_case(2, 32, 48, 64, 64, 3, 1, 1, 2)
    conv_case(2, 32, 48, 64, 64, 3, 1, 1, 2, stats=True)
    conv_case(3, 27, 27, 128, 128, 3, 1, 1, 2, stats=True)
    conv_case(1, 55, 55, 256, 128, 3, 1, 1, 2)
    conv_case(2, 24, 40, 64, 64, 9, 1, 4, 2, tag="[9x9]")
    conv_case(8, 64, 64, 64, 64, 3, 1, 1, 2, tag="[multi-tile/CTA]")
    dgrad_s1_case(2, 20, 24, 64, 128, 3, 1, 2)


def sec_wgrad():
    print("=== tcgen05 wgrad ===", flush=True)
    wgrad_case(2, 32, 48, 64, 64, 3, 1, 1, 2)
    wgrad_case(3, 27, 27, 128, 128, 3, 1, 1, 2)
    wgrad_case(2, 24, 40, 64, 64, 9, 1, 4, 2, tag="[9x9]")
    wgrad_case(8, 64, 64, 64, 64, 3, 1, 1, 2)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("section", choices=["umma", "simt", "tc", "wgrad", "time"])
    args = ap.parse_args()
    print(torch.cuda.get_device_name(0), "lib version", _lib.load().fcd_version(), flush=True)
    t0 = time.time()
    {"umma": umma_probe, "simt": sec_simt, "tc": sec_tc, "wgrad": sec_wgrad, "time": timing}[args.section]()
    print(f"[{args.section}] done in {time.time() - t0:.1f}s", flush=True)
