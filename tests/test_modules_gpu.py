"""The exported building blocks as STAND-ALONE drop-ins (Module.py:18-90, 174-190): `DoubleConv`, `Down`, `Up` (bilinear and
transposed-convolution), `OutConv`, `ResidualBlock` called on NCHW tensors like any nn.Module — forward, input gradients,
parameter gradients and running statistics against the CPU oracle's restatement of the same blocks driven by the module's own
state_dict.  Also: the supported channel widths are validated in the constructors with a clear error, and packed-weight
caches follow `.data` writes after `invalidate_weight_cache()`."""
import re

import pytest
import torch
import torch.nn.functional as F

import fcdgan_b200 as fb
from oracle import fcd_oracle as O
from tests._util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sd(mod, prefix):
    """module state_dict -> oracle dict under `prefix`, requiring gradients."""
    return {f"{prefix}.{k}" if prefix else k: (v.detach().cpu().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
                                                else v.detach().cpu().clone()) for k, v in mod.state_dict().items()}


def _randomise(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, v in mod.state_dict().items():
            if "running_var" in k:
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
            elif "running_mean" in k or (v.is_floating_point() and v.dim() == 1):
                v.copy_(0.3 * torch.randn(v.shape, generator=g) + (1.0 if "weight" in k and v.numel() > 1 else 0.0))


# conv bias in front of a train-mode BatchNorm: analytically zero gradient (rounding noise on both sides)
_ZERO_GRAD = re.compile(r".*double_conv\.[03]\.bias|conv[12]\.bias")


# End-to-end gradient tolerance of a block: 2e-2 in relative L2.  Like for the whole networks (tests/test_networks_gpu.py,
# profiles/r02_gradient_conditioning.log) this is the conditioning of conv -> BatchNorm -> ReLU chains on 24 x 20 tiles — a ReLU
# whose pre-activation lies within the 1e-5 forward error of zero takes the other one-sided derivative — not slack in a kernel:
# every backward kernel is held to 3e-5 in tests/test_ops_gpu.py.  Measured over many weight draws: 3e-3 ... 1.2e-2.  The
# weights are seeded so that the test is the same test on every run.
def _compare(mod, sd, prefix, out, out_o, ins, ins_o, tol_g=2e-2, train=True):
    assert rel_err(out, out_o) < 1e-3, rel_err(out, out_o)
    for a, b in zip(ins, ins_o):
        assert rel_l2(a.grad, b.grad) < tol_g, rel_l2(a.grad, b.grad)
    gmax = max(float(p.grad.abs().max()) for p in mod.parameters())
    for k, p in mod.named_parameters():
        r = sd[f"{prefix}.{k}" if prefix else k].grad
        if train and _ZERO_GRAD.fullmatch(k):
            assert float(p.grad.abs().max()) < 1e-3 * gmax, (k, float(p.grad.abs().max()))
            continue
        assert rel_l2(p.grad, r) < tol_g, (k, rel_l2(p.grad, r))


@pytest.mark.parametrize("train", [True, False])
def test_double_conv_down_residual(train):
    fb.set_precision("parity")
    torch.manual_seed(1234)
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(2, 64, 24, 20, generator=g)
    for name, mod, run, prefix in (
            ("DoubleConv", fb.DoubleConv(64, 128), lambda sd, x: O.double_conv(sd, "m", x, train), "m"),
            ("DoubleConv mid", fb.DoubleConv(64, 64, 128), lambda sd, x: O.double_conv(sd, "m", x, train), "m"),
            ("Down", fb.Down(64, 128), lambda sd, x: O.down(sd, "m", x, train), "m"),
            ("ResidualBlock", fb.ResidualBlock(64), lambda sd, x: O.residual_block(sd, "m", x, train), "m")):
        _randomise(mod, 7)
        mod.to(DEV).train(train)
        sd = _sd(mod, prefix)
        x = x0.to(DEV).requires_grad_(True)
        out = mod(x)
        r = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
        (out * r.to(DEV)).sum().backward()
        xo = x0.clone().requires_grad_(True)
        out_o = run(sd, xo)
        (out_o * r).sum().backward()
        _compare(mod, sd, prefix, out, out_o, [x], [xo], train=train)
        if train:
            own = mod.state_dict()
            for k, v in sd.items():
                if "running" in k:
                    assert rel_err(own[k[len(prefix) + 1:]].float(), v.float()) < 1e-4, (name, k)


@pytest.mark.parametrize("bilinear", [True, False])
def test_up_and_outconv(bilinear):
    fb.set_precision("parity")
    torch.manual_seed(1235)
    g = torch.Generator().manual_seed(3)
    # odd skip size: 11 -> 22 is zero-padded to 23 on the right / bottom (Module.py:70-74)
    c1 = 128
    x1_0 = torch.randn(2, c1, 11, 9, generator=g)
    x2_0 = torch.randn(2, 128 if bilinear else 64, 23, 19, generator=g)
    up = fb.Up(256 if bilinear else 128, 64, bilinear)
    _randomise(up, 8)
    up.to(DEV).train()
    sd = _sd(up, "u")
    x1, x2 = x1_0.to(DEV).requires_grad_(True), x2_0.to(DEV).requires_grad_(True)
    out = up(x1, x2)
    r = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
    (out * r.to(DEV)).sum().backward()
    a, b = x1_0.clone().requires_grad_(True), x2_0.clone().requires_grad_(True)
    out_o = O.up(sd, "u", a, b, bilinear, True)
    (out_o * r).sum().backward()
    _compare(up, sd, "u", out, out_o, [x1, x2], [a, b])
    oc = fb.OutConv(128, 1).to(DEV)
    x0 = torch.randn(2, 128, 17, 21, generator=g)
    x = x0.to(DEV).requires_grad_(True)
    out = oc(x)
    (out * out).sum().backward()
    xo = x0.clone().requires_grad_(True)
    w, bias = oc.conv.weight.detach().cpu().requires_grad_(True), oc.conv.bias.detach().cpu().requires_grad_(True)
    out_o = torch.sigmoid(F.conv2d(xo, w, bias))
    (out_o * out_o).sum().backward()
    assert rel_err(out, out_o) < 1e-3 and rel_l2(x.grad, xo.grad) < 2e-3      # no kink in 1x1 + sigmoid: tight
    assert rel_l2(oc.conv.weight.grad, w.grad) < 2e-3 and rel_l2(oc.conv.bias.grad, bias.grad) < 2e-3


def test_unsupported_widths_raise_in_the_constructor():
    """Channel widths the reduction kernels do not take are rejected where the module is built, not deep inside
    a kernel call."""
    with pytest.raises(ValueError, match="channel"):
        fb.DoubleConv(64, 192)
    with pytest.raises(ValueError, match="channel"):
        fb.Up(100, 64, True)
    with pytest.raises(ValueError, match="OutConv"):
        fb.OutConv(256, 1)
    fb.DoubleConv(13, 64); fb.DoubleConv(2048, 512, 1024); fb.OutConv(128, 1)      # the reference's own widths


def test_weight_cache_follows_data_writes_after_invalidate():
    """Writes through `.data` do not move Tensor._version (WGAN clip `p.data.clamp_`, EMA copies, dist.broadcast(p.data));
    `invalidate_weight_cache()` makes the packed conv weights follow them."""
    fb.set_precision("parity")
    torch.manual_seed(0)
    net = fb.Generator(4).to(DEV).eval()
    x = torch.randn(1, 4, 40, 36, device=DEV)
    with torch.no_grad():
        y0 = net(x)
        for p in net.parameters():
            p.data.mul_(0.5)                   # version counters do not move
        fb.invalidate_weight_cache()
        y1 = net(x)
        fresh = fb.Generator(4).to(DEV).eval()
        fresh.load_state_dict(net.state_dict())
        y2 = fresh(x)
    assert not torch.allclose(y0, y1)
    assert torch.equal(y1, y2)


def test_weight_cache_follows_fused_optimizer_steps():
    """torch's fused optimizers update parameters without moving Tensor._version; the engine invalidates its packed conv weights
    after every optimizer step (global post-step hook), so an eager training loop with `Adam(fused=True)` sees its own updates."""
    fb.set_precision("parity")
    torch.manual_seed(0)
    net = fb.Generator(4).to(DEV).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, fused=True)
    x = torch.randn(2, 4, 40, 36, device=DEV)
    net(x).square().mean().backward()
    v0 = net.block2.conv1.weight._version
    opt.step()
    assert net.block2.conv1.weight._version == v0, "torch changed: fused Adam now moves the version counter (hook still harmless)"
    with torch.no_grad():
        y1 = net.eval()(x)
        fresh = fb.Generator(4).to(DEV).eval()
        fresh.load_state_dict(net.state_dict())
        y2 = fresh(x)
    assert torch.equal(y1, y2)
