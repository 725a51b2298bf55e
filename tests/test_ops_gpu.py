"""Per-kernel GPU parity tests, called through the C ABI (fcdgan_b200._lib / engine) and compared with torch
fp64 on the SAME operands (split-bf16 operands are re-joined for the reference, so the comparison isolates the
kernel's arithmetic).  These are element-wise tight (no activation-kink ambiguity: the reference sees the same
pre-activations).  Tolerances are relative to the reference tensor's max |value| and written per test."""
import pytest
import torch
import torch.nn.functional as F

from fcdgan_b200 import _lib
from fcdgan_b200 import engine as E

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def S():
    return torch.cuda.current_stream().cuda_stream


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def joined(hi, lo):
    return hi.double() + (lo.double() if lo is not None else 0)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def pack(w, Cout_p, Cin_p, mode):
    Cout, Cin, KH, KW = w.shape
    rows, cols = (Cout_p, Cin_p) if mode == 0 else (Cin_p, Cout_p)
    hi = torch.empty(KH * KW, rows, cols, dtype=torch.bfloat16, device=DEV)
    lo = torch.empty_like(hi)
    _lib.call("fcd_pack_conv_weight", w.contiguous().data_ptr(), Cout, Cin, KH, KW, Cout_p, Cin_p, mode, hi.data_ptr(),
              lo.data_ptr(), S())
    return hi, lo


def act_from(x_nchw, Cp=None, ld=None, off=0):
    """NCHW fp32 -> engine Act (optionally a channel slice at `off` of a wider buffer with pitch `ld`)."""
    N, C, H, W = x_nchw.shape
    Cp = Cp or E.pad_ch(C)
    ld = ld or Cp
    buf = E.Act.empty(N, H, W, ld, DEV, ld)
    buf.hi.zero_()
    if buf.lo is not None:
        buf.lo.zero_()
    a = buf if (ld == Cp and off == 0) else buf.slice(off, Cp)
    a.C = C
    _lib.call("fcd_stage_nchw_to_split", x_nchw.contiguous().data_ptr(), None, N, C, H, W, a.p_hi(), a.p_lo(), a.ld, Cp, S())
    return a


def act_value(a):
    """joined value of an Act as NCHW fp64 (logical channels)."""
    return joined(a.hi, a.lo)[..., :a.C].permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # N, H, W, Cin, Cout, K, stride, pad, engine
    (2, 20, 24, 13, 64, 9, 1, 4, E.ENGINE_SIMT),     # Generator head  Module.py:146
    (2, 20, 24, 64, 13, 9, 1, 4, E.ENGINE_SIMT),     # Generator tail  Module.py:158
    (2, 32, 32, 13, 64, 3, 2, 1, E.ENGINE_SIMT),     # Discriminator layer 0  Module.py:196
    (2, 27, 27, 64, 128, 3, 2, 1, E.ENGINE_SIMT),    # odd size, stride 2
    (2, 32, 48, 64, 64, 3, 1, 1, E.ENGINE_TC),       # residual-block conv  Module.py:177
    (3, 27, 27, 128, 128, 3, 1, 1, E.ENGINE_TC),     # odd size, partial tiles
    (1, 55, 55, 256, 128, 3, 1, 1, E.ENGINE_TC),
    (2, 24, 40, 64, 64, 9, 1, 4, E.ENGINE_TC),
    (8, 64, 64, 64, 64, 3, 1, 1, E.ENGINE_TC),       # several tiles per CTA
    (2, 5, 4, 64, 64, 3, 1, 1, E.ENGINE_TC),         # tensor smaller than one TMA box (U-Net bottleneck)
    (2, 2, 2, 128, 64, 3, 1, 1, E.ENGINE_TC),
    (2, 20, 24, 13, 64, 9, 1, 4, E.ENGINE_TC),       # Generator head on the tcgen05 engine (13 bands zero-padded to 64)
    (2, 20, 24, 64, 13, 9, 1, 4, E.ENGINE_TC),       # Generator tail
    (2, 19, 21, 4, 64, 3, 1, 1, E.ENGINE_TC),        # Segmentor first layer, 4 bands
    (2, 32, 32, 13, 64, 3, 2, 1, E.ENGINE_TC),       # Discriminator layers on the tcgen05 engine: TMA element strides
    (2, 44, 36, 64, 128, 3, 2, 1, E.ENGINE_TC),      #   (forward / wgrad) and parity-class dgrad
    (3, 27, 27, 128, 256, 3, 2, 1, E.ENGINE_TC),     #   odd sizes
    (2, 11, 9, 256, 512, 3, 2, 1, E.ENGINE_TC),
    (2, 20, 24, 64, 128, 3, 1, 1, E.ENGINE_TC),      # halo kernel: two 64-wide weight blocks (split) / one 128-wide block (fast)
    (1, 33, 17, 64, 256, 3, 1, 1, E.ENGINE_TC),      # halo kernel, n_blocks = 4 (split) / 2 (fast), ragged tiles
    (2, 16, 16, 128, 64, 1, 1, 0, E.ENGINE_TC),      # 1x1 convolution (im2col first layer, convT pixel shuffles) on the halo kernel
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("fast", [False, True])
def test_conv_fwd_wgrad_dgrad(case, fast):
    """fcd_conv2d_fwd / fcd_conv2d_wgrad / dgrad vs F.conv2d autograd in fp64.  Tolerance: 2e-5 in parity
    precision (three bf16 MMAs, fp32 accumulate), exact-operand comparison 1e-5 in fast precision (the
    reference is given the bf16-rounded operands)."""
    N, H, W, Cin, Cout, K, stride, pad, engine = case
    torch.manual_seed(0)
    E.set_precision("fast" if fast else "parity")
    try:
        Cin_p, Cout_p = E.pad_ch(Cin, engine == E.ENGINE_TC), E.pad_ch(Cout, engine == E.ENGINE_TC)
        x = torch.randn(N, Cin, H, W, device=DEV)
        w = torch.randn(Cout, Cin, K, K, device=DEV) * 0.1
        b = torch.randn(Cout, device=DEV)
        a = act_from(x, Cp=Cin_p)
        wh, wl = pack(w, Cout_p, Cin_p, 0)
        OH = (H + 2 * pad - K) // stride + 1
        OW = (W + 2 * pad - K) // stride + 1
        z = torch.full((N, OH, OW, Cout_p), float("nan"), device=DEV)
        bias = E._padded_vec(b, Cout_p)
        st = torch.zeros(2, Cout_p, dtype=torch.float64, device=DEV)
        _lib.call("fcd_conv2d_fwd", a.p_hi(), a.p_lo(), a.ld, wh.data_ptr(), None if fast else wl.data_ptr(), bias.data_ptr(),
                  None, 0, z.data_ptr(), Cout_p, N, H, W, Cin_p, Cout_p, K, K, stride, pad, st[0].data_ptr(),
                  st[1].data_ptr(), engine, S())
        xr = act_value(a).clone().requires_grad_(True)
        wr = (wh.double() + (0 if fast else wl.double())).view(K, K, Cout_p, Cin_p)[:, :, :Cout, :Cin].permute(2, 3, 0, 1)
        wr = wr.clone().requires_grad_(True)
        ref = F.conv2d(xr, wr, b.double(), stride=stride, padding=pad)
        got = z[..., :Cout].permute(0, 3, 1, 2)
        tol = 1e-5 if fast else 3e-5
        assert rel(got, ref) < tol
        assert torch.isfinite(z).all()
        assert rel(st[0, :Cout], ref.sum(dim=(0, 2, 3))) < 1e-4 and rel(st[1, :Cout], (ref * ref).sum(dim=(0, 2, 3))) < 1e-4
        # ---- backward operands
        dz = torch.randn(N, Cout, OH, OW, device=DEV)
        g = act_from(dz, Cp=Cout_p)
        gr = act_value(g)
        gx, gw = torch.autograd.grad(ref, (xr, wr), gr)
        dw = torch.full((Cout, Cin, K, K), float("nan"), device=DEV)
        db = torch.full((Cout,), float("nan"), device=DEV)
        nbytes = _lib.load().fcd_conv2d_wgrad_workspace(N, H, W, Cin_p, Cout_p, K, K, stride, pad, engine)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=DEV)
        _lib.call("fcd_conv2d_wgrad", a.p_hi(), a.p_lo(), a.ld, g.p_hi(), g.p_lo(), g.ld, dw.data_ptr(), db.data_ptr(), N,
                  H, W, Cin, Cin_p, Cout, Cout_p, K, K, stride, pad, 0, ws.data_ptr(), nbytes, engine, S())
        assert rel(dw, gw) < 3e-5
        assert rel(db, gr.sum(dim=(0, 2, 3))) < 3e-5
        # accumulate = 1 doubles it
        _lib.call("fcd_conv2d_wgrad", a.p_hi(), a.p_lo(), a.ld, g.p_hi(), g.p_lo(), g.ld, dw.data_ptr(), db.data_ptr(), N,
                  H, W, Cin, Cin_p, Cout, Cout_p, K, K, stride, pad, 1, ws.data_ptr(), nbytes, engine, S())
        assert rel(dw, 2 * gw) < 3e-5
        # dgrad (+ addend)
        dx = torch.full((N, H, W, Cin_p), float("nan"), device=DEV)
        addend = torch.randn(N, H, W, Cin_p, device=DEV)
        if stride == 1:
            wdh, wdl = pack(w, Cout_p, Cin_p, 1)
            _lib.call("fcd_conv2d_fwd", g.p_hi(), g.p_lo(), g.ld, wdh.data_ptr(), None if fast else wdl.data_ptr(), None,
                      addend.data_ptr(), Cin_p, dx.data_ptr(), Cin_p, N, OH, OW, Cout_p, Cin_p, K, K, 1, K - 1 - pad, None,
                      None, engine, S())
        else:
            wdh, wdl = pack(w, Cout_p, Cin_p, 1)
            _lib.call("fcd_conv2d_dgrad_strided", g.p_hi(), g.p_lo(), g.ld, wh.data_ptr(), None if fast else wl.data_ptr(),
                      wdh.data_ptr(), None if fast else wdl.data_ptr(), addend.data_ptr(), Cin_p, dx.data_ptr(), Cin_p, N, H, W,
                      Cin_p, Cout_p, K, K, stride, pad, engine, S())
        want = gx + addend[..., :Cin].permute(0, 3, 1, 2).double()
        assert rel(dx[..., :Cin].permute(0, 3, 1, 2), want) < 3e-5
        if stride == 1:
            # in-place accumulation (addend IS the output, the form engine.conv's backward uses: TMA reduce-add epilogue)
            _lib.call("fcd_conv2d_fwd", g.p_hi(), g.p_lo(), g.ld, wdh.data_ptr(), None if fast else wdl.data_ptr(), None,
                      dx.data_ptr(), Cin_p, dx.data_ptr(), Cin_p, N, OH, OW, Cout_p, Cin_p, K, K, 1, K - 1 - pad, None,
                      None, engine, S())
            assert rel(dx[..., :Cin].permute(0, 3, 1, 2), want + gx) < 3e-5
    finally:
        E.set_precision("parity")


def test_conv_reads_concat_slice():
    """virtual concat: a convolution reading a channel slice of a wider NHWC buffer (pitch > channels)."""
    torch.manual_seed(1)
    N, H, W, C = 2, 20, 20, 64
    x = torch.randn(N, C, H, W, device=DEV)
    a = act_from(x, Cp=64, ld=192, off=64)
    w = torch.randn(64, 64, 3, 3, device=DEV) * 0.1
    wh, wl = pack(w, 64, 64, 0)
    z = torch.empty(N, H, W, 64, device=DEV)
    _lib.call("fcd_conv2d_fwd", a.p_hi(), a.p_lo(), a.ld, wh.data_ptr(), wl.data_ptr(), None, None, 0, z.data_ptr(), 64, N, H,
              W, 64, 64, 3, 3, 1, 1, None, None, E.ENGINE_TC, S())
    wr = joined(wh, wl).view(3, 3, 64, 64).permute(2, 3, 0, 1)
    assert rel(z.permute(0, 3, 1, 2), F.conv2d(act_value(a), wr, None, padding=1)) < 2e-5


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act,use_bn,training,residual", [
    (E.ACT_RELU, True, True, False), (E.ACT_PRELU, True, True, False), (E.ACT_NONE, True, True, True),
    (E.ACT_LEAKY, True, True, False), (E.ACT_PRELU, False, True, False), (E.ACT_LEAKY, False, True, False),
    (E.ACT_RELU, True, False, False), (E.ACT_NONE, True, False, True)])
def test_bn_act_fwd_bwd(act, use_bn, training, residual):
    """BatchNorm2d (train: batch stats + running-stat update; eval: running stats) + activation + residual,
    forward and backward, through the engine primitives, vs torch autograd in fp64.  Tolerance 2e-5
    (output is re-split to bf16 pairs: 2^-17 relative)."""
    torch.manual_seed(2)
    N, C, H, W = 3, 64, 13, 11
    tape = E.Tape(DEV, True)
    zt = torch.randn(N, H, W, C, device=DEV) * 1.5 + 0.3
    z = E.Z(zt, N, H, W, C, C)
    st = torch.zeros(2, C, dtype=torch.float64, device=DEV)
    z.sum, z.sqsum = st[0], st[1]
    _lib.call("fcd_bn_stats", zt.data_ptr(), C, z.npix, C, st[0].data_ptr(), st[1].data_ptr(), S())
    gamma = (1 + 0.2 * torch.randn(C, device=DEV)).requires_grad_(True)
    beta = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    rm, rv = 0.1 * torch.randn(C, device=DEV), 1 + 0.2 * torch.rand(C, device=DEV)
    rm0, rv0 = rm.clone(), rv.clone()
    slope = torch.full((1,), 0.25, device=DEV, requires_grad=True)
    bn = E.BN(gamma, beta, rm, rv, torch.zeros((), dtype=torch.long, device=DEV)) if use_bn else None
    res_nchw = torch.randn(N, C, H, W, device=DEV)
    res = tape.track(act_from(res_nchw)) if residual else None
    out = E.bn_act(tape, z, bn, training, act, slope=slope if act == E.ACT_PRELU else None, slope_const=0.2, residual=res)
    # reference
    zr = zt.double().permute(0, 3, 1, 2).clone().requires_grad_(True)
    g64, b64, s64 = gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True), \
        slope.detach().double().requires_grad_(True)
    rmr, rvr = rm0.double().clone(), rv0.double().clone()
    u = F.batch_norm(zr, rmr, rvr, g64, b64, training, 0.1, 1e-5) if use_bn else zr
    if act == E.ACT_RELU:
        r = F.relu(u)
    elif act == E.ACT_PRELU:
        r = F.prelu(u, s64)
    elif act == E.ACT_LEAKY:
        r = F.leaky_relu(u, 0.2)
    else:
        r = u
    rres = act_value(res).clone().requires_grad_(True) if residual else None
    if residual:
        r = r + rres
    assert rel(act_value(out), r) < 2e-5
    if use_bn and training:
        assert rel(rm, rmr) < 1e-6 and rel(rv, rvr) < 1e-6
        assert int(bn.num_batches_tracked.item()) == 1
    # backward: upstream gradient arrives as fp32 NHWC in out.grad
    dr = torch.randn(N, C, H, W, device=DEV)
    out.grad.copy_(dr.permute(0, 2, 3, 1))
    out.mark_ready()
    tape.ops[-1](tape)
    wants = [zr] + ([g64, b64] if use_bn else []) + ([s64] if act == E.ACT_PRELU else []) + ([rres] if residual else [])
    grads = torch.autograd.grad(r, wants, dr.double())
    dz_ref = grads[0]
    assert rel(act_value(z.dz), dz_ref) < 3e-5
    k = 1
    if use_bn:
        assert rel(tape.pgrads[id(gamma)], grads[k]) < 3e-5 and rel(tape.pgrads[id(beta)], grads[k + 1]) < 3e-5
        k += 2
    if act == E.ACT_PRELU:
        assert rel(tape.pgrads[id(slope)], grads[k]) < 3e-5
        k += 1
    if residual:
        assert res.ready and rel(res.grad.permute(0, 3, 1, 2), grads[k]) < 1e-6


@pytest.mark.parametrize("accumulate", [False, True])
@pytest.mark.parametrize("H,W", [(16, 16), (11, 9), (27, 55), (8, 13), (5, 2)])
def test_maxpool(H, W, accumulate):
    """nn.MaxPool2d(2) floor semantics + first-max gradient routing (Module.py:44); exact (selection only).  The gradient is
    written (pixels of a dropped odd row / column get 0) or added to one that is already there (the skip connection's)."""
    torch.manual_seed(3)
    N, C = 2, 64
    x = torch.randn(N, C, H, W, device=DEV)
    x[0, :, 0:2, 0:2] = 1.5          # ties inside a window: torch routes the gradient to the first maximum
    tape = E.Tape(DEV, True)
    a = tape.track(act_from(x))
    out = E.maxpool2(tape, a)
    xr = act_value(a).clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 2)
    assert rel(act_value(out), ref) < 1e-7
    d = torch.randn_like(ref)
    out.grad.copy_(d.permute(0, 2, 3, 1).float())
    out.mark_ready()
    pre = torch.randn(N, H, W, a.Cp, device=DEV)
    if accumulate:
        a.grad.copy_(pre)
        a.mark_ready()
    else:
        a.grad.fill_(float("nan"))   # every element must be written
    tape.ops[-1](tape)
    (gx,) = torch.autograd.grad(ref, xr, d)
    if accumulate:
        gx = gx + pre[..., :C].permute(0, 3, 1, 2).double()
    assert rel(a.grad[..., :C].permute(0, 3, 1, 2), gx) < 1e-6


@pytest.mark.parametrize("h,w,H,W", [(8, 8, 16, 16), (13, 5, 27, 11), (12, 12, 25, 25), (1, 2, 2, 4)])
def test_upsample_into_concat(h, w, H, W):
    """bilinear x2 align_corners=True + F.pad (extra row/col bottom/right) written into a concat slot
    (Module.py:60,70-78).  Tolerance 2e-5 (bf16-pair output)."""
    torch.manual_seed(4)
    N, C = 2, 64
    x = torch.randn(N, C, h, w, device=DEV)
    tape = E.Tape(DEV, True)
    a = tape.track(act_from(x))
    cat = tape.new_act(N, H, W, 192, 192)
    slot = tape.track(cat.slice(128, 64))
    E.upsample2x_into(tape, a, slot)
    xr = act_value(a).clone().requires_grad_(True)
    up = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    dY, dX = H - up.shape[2], W - up.shape[3]
    ref = F.pad(up, [dX // 2, dX - dX // 2, dY // 2, dY - dY // 2])
    assert rel(act_value(slot), ref) < 2e-5
    d = torch.randn(N, H, W, 192, device=DEV)
    cat.grad.copy_(d)
    cat.mark_ready()
    tape.ops[-1](tape)
    (gx,) = torch.autograd.grad(ref, xr, d[..., 128:192].permute(0, 3, 1, 2).double())
    assert rel(a.grad.permute(0, 3, 1, 2), gx) < 1e-5


def test_conv_transpose_into_concat():
    """nn.ConvTranspose2d(k2,s2) + pad + concat slot (Module.py:63,70-78) fwd/bwd.  Tolerance 3e-5."""
    torch.manual_seed(5)
    N, Cin, Cout, h, w, H, W = 2, 128, 64, 6, 5, 13, 11
    x = torch.randn(N, Cin, h, w, device=DEV)
    wt = (torch.randn(Cin, Cout, 2, 2, device=DEV) * 0.1).requires_grad_(True)
    b = torch.randn(Cout, device=DEV).requires_grad_(True)
    tape = E.Tape(DEV, True)
    a = tape.track(act_from(x))
    cat = tape.new_act(N, H, W, 128, 128)
    slot = tape.track(cat.slice(64, 64))
    E.conv_transpose2x2_into(tape, a, wt, b, slot)
    xr = act_value(a).clone().requires_grad_(True)
    w64 = joined(*split(wt.detach())).requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True)
    up = F.conv_transpose2d(xr, w64, b64, stride=2)
    dY, dX = H - up.shape[2], W - up.shape[3]
    ref = F.pad(up, [dX // 2, dX - dX // 2, dY // 2, dY - dY // 2])
    assert rel(act_value(slot), ref) < 3e-5
    d = torch.randn(N, H, W, 128, device=DEV)
    cat.grad.copy_(d)
    cat.mark_ready()
    tape.ops[-1](tape)
    gx, gw, gb = torch.autograd.grad(ref, (xr, w64, b64), d[..., 64:].permute(0, 3, 1, 2).double())
    assert rel(a.grad.permute(0, 3, 1, 2), gx) < 3e-5
    assert rel(tape.pgrads[id(wt)], gw) < 3e-5 and rel(tape.pgrads[id(b)], gb) < 3e-5


def test_outconv_sigmoid():
    """OutConv 1x1 + sigmoid (Module.py:82-90) fwd/bwd.  Tolerance 1e-5."""
    torch.manual_seed(6)
    N, C, H, W = 2, 128, 19, 23
    x = torch.randn(N, C, H, W, device=DEV)
    w = (torch.randn(1, C, 1, 1, device=DEV) * 0.2).requires_grad_(True)
    b = torch.randn(1, device=DEV).requires_grad_(True)
    tape = E.Tape(DEV, True)
    a = tape.track(act_from(x))
    slot = {}
    out = E.outconv_sigmoid(tape, a, w, b, slot)
    xr = act_value(a).clone().requires_grad_(True)
    w64, b64 = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    ref = torch.sigmoid(F.conv2d(xr, w64, b64))
    assert rel(out, ref) < 1e-5
    d = torch.randn_like(out)
    slot["dout"] = d
    tape.ops[-1](tape)
    gx, gw, gb = torch.autograd.grad(ref, (xr, w64, b64), d.double())
    assert rel(a.grad.permute(0, 3, 1, 2), gx) < 1e-5
    assert rel(tape.pgrads[id(w)], gw) < 1e-5 and rel(tape.pgrads[id(b)], gb) < 1e-5


def test_disc_head():
    """classifier(GAP(fx - fy)) -> sigmoid (Module.py:212-223) fwd/bwd.  Tolerance 2e-5."""
    torch.manual_seed(7)
    N, C, H, W = 3, 512, 3, 2
    fx, fy = torch.randn(N, C, H, W, device=DEV), torch.randn(N, C, H, W, device=DEV)
    w1 = (torch.randn(1024, 512, 1, 1, device=DEV) * 0.05).requires_grad_(True)
    b1 = torch.randn(1024, device=DEV).requires_grad_(True)
    w2 = (torch.randn(1, 1024, 1, 1, device=DEV) * 0.05).requires_grad_(True)
    b2 = torch.randn(1, device=DEV).requires_grad_(True)
    tape = E.Tape(DEV, True)
    ax, ay = tape.track(act_from(fx)), tape.track(act_from(fy))
    slot = {}
    out = E.disc_head(tape, ax, ay, w1, b1, w2, b2, slot)
    xr, yr = act_value(ax).clone().requires_grad_(True), act_value(ay).clone().requires_grad_(True)
    P = [p.detach().double().requires_grad_(True) for p in (w1, b1, w2, b2)]
    hcl = F.adaptive_avg_pool2d(xr - yr, 1)
    hcl = F.leaky_relu(F.conv2d(hcl, P[0], P[1]), 0.2)
    ref = torch.sigmoid(F.conv2d(hcl, P[2], P[3]).view(N))
    assert out.shape == (N,) and rel(out, ref) < 2e-5
    d = torch.randn(N, device=DEV)
    slot["dout"] = d
    tape.ops[-1](tape)
    g = torch.autograd.grad(ref, [xr, yr] + P, d.double())
    assert rel(ax.grad.permute(0, 3, 1, 2), g[0]) < 2e-5 and rel(ay.grad.permute(0, 3, 1, 2), g[1]) < 2e-5
    for p, gr in zip((w1, b1, w2, b2), g[2:]):
        assert rel(tape.pgrads[id(p)], gr) < 2e-5


def test_stage_roundtrip_and_mask():
    """NCHW -> split NHWC (+ fused soft mask x*(1-m), Demo_RSSS.py:290) -> NCHW."""
    torch.manual_seed(8)
    N, C, H, W = 2, 13, 9, 7
    x = torch.randn(N, C, H, W, device=DEV)
    m = torch.rand(N, 1, H, W, device=DEV)
    a = E.Act.empty(N, H, W, C, DEV)
    _lib.call("fcd_stage_nchw_to_split", x.data_ptr(), m.data_ptr(), N, C, H, W, a.p_hi(), a.p_lo(), a.ld, a.Cp, S())
    assert rel(act_value(a), (x * (1 - m)).double()) < 1e-5
    assert (joined(a.hi, a.lo)[..., C:] == 0).all()
    out = torch.empty_like(x)
    _lib.call("fcd_unstage_split_to_nchw", a.p_hi(), a.p_lo(), a.ld, N, C, H, W, out.data_ptr(), S())
    assert rel(out, x * (1 - m)) < 1e-5


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rowpack", [False, True])
@pytest.mark.parametrize("C,Cout,K,pad,H,W", [(13, 64, 9, 4, 20, 24), (4, 64, 3, 1, 19, 21), (3, 64, 9, 4, 33, 40), (16, 128, 3, 1, 8, 8),
                                              (13, 64, 9, 4, 9, 256), (4, 128, 9, 4, 21, 19)])
def test_conv_small_in_pack4(C, Cout, K, pad, H, W, rowpack):
    """<= 16-band input convolution on the channel-packed input (Generator head Module.py:146, Segmentor first
    layer Module.py:26), 4-pixel form and tight row-packed form (one tap per filter row, engine.row_pack_pixels):
    forward, BN statistics, wgrad, db vs F.conv2d autograd in fp64.  Tolerance 3e-5."""
    torch.manual_seed(9)
    N = 2
    x = torch.randn(N, C, H, W, device=DEV)
    w = (torch.randn(Cout, C, K, K, device=DEV) * 0.1).requires_grad_(True)
    b = torch.randn(Cout, device=DEV).requires_grad_(True)
    tape = E.Tape(DEV, True)
    P = E.row_pack_pixels(C, K) if rowpack else 4
    if rowpack and P == 4:
        pytest.skip("the row-packed form does not apply to this shape (same launch as the 4-pixel case)")
    xp = E.PackedAct(x, P=P)
    # the staged operand itself: v[n, h, w'', j*cs + c] = x[n, c, h, w'' - M + j]
    cs = 16 if P == 4 else C
    v = joined(xp.hi, xp.lo)
    want = torch.zeros_like(v)
    xs = joined(*split(x))
    for j in range(P):
        lo_w, hi_w = max(0, xp.M - j), min(xp.Wp, W + xp.M - j)
        want[:, :, lo_w:hi_w, j * cs:j * cs + C] = xs[:, :, :, lo_w - xp.M + j:hi_w - xp.M + j].permute(0, 2, 3, 1)
    assert rel(v, want) < 1e-6 and bool((v[want == 0] == 0).all())
    z = E.conv_small_in(tape, xp, w, b, pad, stats=True)
    # reference on the bf16-pair-rounded operands
    xr = joined(*split(x))
    wr = joined(*split(w.detach())).requires_grad_(True)
    br = b.detach().double().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=pad)
    got = z.t[..., :Cout].permute(0, 3, 1, 2)
    assert rel(got, ref) < 3e-5
    assert rel(z.sum[:Cout], ref.sum(dim=(0, 2, 3))) < 1e-4 and rel(z.sqsum[:Cout], (ref * ref).sum(dim=(0, 2, 3))) < 1e-4
    dz = torch.randn_like(ref).float()
    z.dz = act_from(dz, Cp=z.Cp)
    gr = act_value(z.dz)
    tape.ops[-1](tape)
    gw, gb = torch.autograd.grad(ref, (wr, br), gr)
    assert rel(tape.pgrads[id(w)], gw) < 3e-5 and rel(tape.pgrads[id(b)], gb) < 3e-5


@pytest.fixture
def rowpack_switch(request):
    E.set_rowpack(request.param)
    yield request.param
    E.set_rowpack(True)


@pytest.mark.parametrize("rowpack_switch", [False, True], indirect=True)
@pytest.mark.parametrize("Cin,Cout,K,pad,H,W", [(64, 13, 9, 4, 20, 24), (64, 4, 9, 4, 17, 32), (128, 3, 3, 1, 12, 8), (64, 13, 9, 4, 10, 256)])
def test_conv_small_out_pack4(Cin, Cout, K, pad, H, W, rowpack_switch):
    """<= 16-channel output convolution (Generator tail Module.py:158): four output pixels per MMA row; dgrad and wgrad on
    the channel-packed output gradient (4-pixel form / tight row-packed form).  Tolerance 3e-5."""
    torch.manual_seed(10)
    N = 2
    x = torch.randn(N, Cin, H, W, device=DEV)
    w = (torch.randn(Cout, Cin, K, K, device=DEV) * 0.05).requires_grad_(True)
    b = torch.randn(Cout, device=DEV).requires_grad_(True)
    tape = E.Tape(DEV, True)
    a = tape.track(act_from(x))
    z = E.conv_small_out(tape, a, w, b, pad)
    xr = act_value(a).clone().requires_grad_(True)
    wr = joined(*split(w.detach())).requires_grad_(True)
    br = b.detach().double().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=pad)
    got = z.t[..., :Cout].permute(0, 3, 1, 2)
    assert rel(got, ref) < 3e-5
    slot = {}
    out = E.z_to_nchw(tape, z, slot)
    assert rel(out, ref) < 3e-5
    dy = torch.randn_like(out)
    slot["dout"] = dy
    pre = torch.randn(N, H, W, a.Cp, device=DEV)          # an already accumulated gradient (addend path)
    a.grad.copy_(pre)
    a.mark_ready()
    tape.ops[-1](tape)      # z_to_nchw backward: packs dy
    tape.ops[-2](tape)      # conv backward
    gyr = joined(*split(dy))
    gx, gw, gb = torch.autograd.grad(ref, (xr, wr, br), gyr)
    assert rel(a.grad[..., :Cin].permute(0, 3, 1, 2), gx + pre[..., :Cin].permute(0, 3, 1, 2).double()) < 3e-5
    assert rel(tape.pgrads[id(w)], gw) < 3e-5 and rel(tape.pgrads[id(b)], gb) < 3e-5
