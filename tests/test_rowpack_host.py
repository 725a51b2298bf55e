"""Host-side index algebra of the tight row-packed operand layout (engine.row_pack_pixels / PackedAct with P = KW,
fcd_stage_nchw_to_split_rowpack): the weight reshuffles of engine.conv_small_in / conv_small_out and the tap geometry they
pass to fcd_conv2d_taps_fwd / fcd_conv2d_taps_wgrad, emulated on the CPU in float64 straight from the semantics
include/fcd_b200.h documents for those entry points, must reproduce nn.Conv2d (Module.py:146,158) and its autograd."""
import pytest
import torch
import torch.nn.functional as F

from fcdgan_b200 import engine as E


def stage_rowpack(img, M, P, Kp):
    """dst[n,h,w'',j*C+c] = src[n,c,h,w''-M+j], j < P (include/fcd_b200.h: fcd_stage_nchw_to_split_rowpack)."""
    N, C, H, W = img.shape
    v = torch.zeros(N, H, W + M, Kp, dtype=img.dtype)
    for j in range(P):
        for wq in range(W + M):
            w = wq - M + j
            if 0 <= w < W:
                v[:, :, wq, j * C:(j + 1) * C] = img[:, :, :, w].permute(0, 2, 1)
    return v


def taps_fwd(x, w, OH, OW, n_r, n_s, dh0, dh_step, dw0, dw_step):
    """z[n,oh,ow,:] = sum_{i,j} x[n, oh+dh0+i*dh_step, ow+dw0+j*dw_step, :] . w[i*n_s+j]  (fcd_conv2d_taps_fwd, csh = csw = 1)."""
    N, XH, XW, _ = x.shape
    z = torch.zeros(N, OH, OW, w.shape[0], dtype=x.dtype)
    for i in range(n_r):
        for j in range(n_s):
            for oh in range(OH):
                h = oh + dh0 + i * dh_step
                if not 0 <= h < XH:
                    continue
                for ow in range(OW):
                    ww = ow + dw0 + j * dw_step
                    if 0 <= ww < XW:
                        z[:, oh, ow, :] += x[:, h, ww, :] @ w[:, :, i, j].T
    return z


def taps_wgrad(x, dz, n_r, n_s, dh0, dw0, dw_step):
    """dw[co][ci][r][s] = sum_{n,h<GH,w<GW} x[n,h+r+dh0,w+s*dw_step+dw0,ci] * dz[n,h,w,co]  (fcd_conv2d_taps_wgrad)."""
    _, XH, XW, Cin = x.shape
    _, GH, GW, Cout = dz.shape
    dw = torch.zeros(Cout, Cin, n_r, n_s, dtype=x.dtype)
    for r in range(n_r):
        for s in range(n_s):
            for h in range(GH):
                hh = h + r + dh0
                if not 0 <= hh < XH:
                    continue
                for w in range(GW):
                    ww = w + s * dw_step + dw0
                    if 0 <= ww < XW:
                        dw[:, :, r, s] += dz[:, h, w, :].T @ x[:, hh, ww, :]
    return dw


def test_row_pack_pixels_rule():
    assert E.row_pack_pixels(13, 9) == 9      # 117 -> 2 chunks of 64 instead of 3 taps of 64
    assert E.row_pack_pixels(3, 9) == 9 and E.row_pack_pixels(4, 9) == 9      # 27 / 36 -> 1 chunk instead of 3
    assert E.row_pack_pixels(13, 3) == 4      # 3x3 filters: one tap either way
    assert E.row_pack_pixels(16, 9) == 4      # 144 -> 3 chunks: no gain
    E.set_rowpack(False)
    try:
        assert E.row_pack_pixels(13, 9) == 4
    finally:
        E.set_rowpack(True)


@pytest.mark.parametrize("C,H,W", [(13, 10, 12), (3, 9, 8)])
def test_rowpack_geometry_reproduces_conv2d_and_its_gradients(C, H, W):
    torch.manual_seed(C)
    dt = torch.float64
    KH = KW = 9
    pad, M = 4, E.PACK_M
    P = E.row_pack_pixels(C, KW)
    Kp = E.pad_ch(P * C)
    N, Cout = 2, 5
    # engine.conv_small_in: forward and weight gradient on the row-packed INPUT
    img = torch.randn(N, C, H, W, dtype=dt)
    w = torch.randn(Cout, C, KH, KW, dtype=dt, requires_grad=True)
    ref = F.conv2d(img, w, padding=pad)
    xp = stage_rowpack(img, M, P, Kp)
    z = taps_fwd(xp, E._w_rowpack_in(w.detach()), H, W, KH, 1, -pad, 1, -pad + M, 1)
    assert (z.permute(0, 3, 1, 2) - ref).abs().max() < 1e-11
    g = torch.randn_like(ref)
    (gw,) = torch.autograd.grad(ref, w, g)
    dwp = taps_wgrad(xp, g.permute(0, 2, 3, 1).contiguous(), KH, 1, -pad, -pad + M, 1)
    assert (E._w_unrowpack_in(dwp, C, KW) - gw).abs().max() < 1e-11
    # engine.conv_small_out backward: weight / bias / data gradient on the row-packed OUTPUT gradient
    Cin, Co = 6, C
    xa = torch.randn(N, Cin, H, W, dtype=dt, requires_grad=True)
    w2 = torch.randn(Co, Cin, KH, KW, dtype=dt, requires_grad=True)
    out = F.conv2d(xa, w2, padding=pad)
    go = torch.randn_like(out)
    gx, gw2 = torch.autograd.grad(out, (xa, w2), go)
    dzp = stage_rowpack(go, M, P, Kp)
    x_nhwc = xa.detach().permute(0, 2, 3, 1).contiguous()
    T = taps_wgrad(x_nhwc, dzp, KH, 1, -pad, (KW - 1) - M - pad, 1)
    assert (E._w_unrowpack_out(T, Co, KW) - gw2).abs().max() < 1e-11
    assert (dzp.sum((0, 1, 2))[:Co] - go.sum((0, 2, 3))).abs().max() < 1e-11
    wd = E._w_rowpack_in(w2.detach().flip(2, 3).permute(1, 0, 2, 3))
    dx = taps_fwd(dzp, wd, H, W, KH, 1, -(KH - 1 - pad), 1, -(KW - 1 - pad) + M, 1)
    assert (dx.permute(0, 3, 1, 2) - gx).abs().max() < 1e-11
