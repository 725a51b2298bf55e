"""GPU parity of the perception loss on the conv engine (fcdgan_b200.losses.PerceptionLoss) against
golden vectors recorded from the UNMODIFIED reference's PerceptionLoss / CNetLoss / CGeneratorLoss (Loss.py:17-124) run with
torchvision's VGG16 initialised under a fixed seed (oracle/make_golden_perception.py; the ImageNet weights of Loss.py:25
cannot be downloaded here, and a seeded random VGG16 pins the arithmetic just as well).

Tolerance: 1e-3 relative on the loss values (13 convolution layers deep, parity precision = split-bf16 operands; measured
< 1e-4).  Gradients: 2e-2 in relative L2, 1e-1 of the tensor maximum element-wise — the conditioning of the function, not of
the kernels: in the fp64 oracle itself a 1e-5 relative perturbation of the generated image (the size of the CUDA path's forward
error) moves d loss / d image by 5.3e-3 in relative L2 and 3.3e-2 of its maximum (13 ReLU layers + 4 max-pools: every kink
crossed flips a whole receptive field's worth of gradient; tests/test_oracle_golden.py::test_perception_gradient_conditioning,
profiles/r02_gradient_conditioning.log).  Measured here: 2e-3 ... 8e-3 in L2, 1e-2 ... 4e-2 element-wise."""
import pytest
import torch

import fcdgan_b200 as fb
from oracle import fcd_oracle as O
from tests._util import load_golden, rel_err, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VTOL, G_L2, G_MAX = 1e-3, 2e-2, 1e-1


@pytest.fixture(scope="module")
def vgg():
    f = load_golden("perception.pt")
    net = O.vgg16_features(1234)
    chk = float(sum(p.double().abs().sum() for p in net.parameters()))
    if abs(chk - f["vgg_checksum"]) > 1e-6 * chk:
        pytest.skip("this torchvision initialises VGG16 differently from the fixture's")
    return net.to(DEV)


def _close(a, ref, tol=VTOL):
    return abs(float(a) - ref) <= tol * abs(ref)


def _check_grads(g, cmap, d, what):
    assert rel_l2(g.grad, d["dg"]) < G_L2 and rel_err(g.grad, d["dg"]) < G_MAX, (what, rel_l2(g.grad, d["dg"]), rel_err(g.grad, d["dg"]))
    if cmap is not None:
        assert rel_l2(cmap.grad, d["dcmap"]) < G_L2 and rel_err(cmap.grad, d["dcmap"]) < G_MAX, \
            (what, rel_l2(cmap.grad, d["dcmap"]), rel_err(cmap.grad, d["dcmap"]))


@pytest.mark.parametrize("key,layers,per_band", [("perband", 1, True), ("rgb5", 5, False)])
def test_perception_loss_golden(vgg, key, layers, per_band):
    fb.set_precision("parity")
    d = load_golden("perception.pt")[key]
    pl = fb.PerceptionLoss(feature_layer=layers, perception_perBand=per_band, vgg_features=vgg)
    g = d["g"].to(DEV).requires_grad_(True)
    cmap = d["cmap"].to(DEV).requires_grad_(True)
    from fcdgan_b200 import engine as E
    E.PROFILE = []
    try:
        v = pl(d["t"].to(DEV), g, cmap)
        tags = [t[1] for t in E.PROFILE]
    finally:
        E.PROFILE = None
    n_conv = {1: 13, 5: 13}[layers]
    assert sum(t.startswith("conv_fwd_tc") for t in tags) == n_conv, tags      # the whole VGG stack ran on the tcgen05 engine
    assert _close(v, d["value"]), (float(v), d["value"])
    v.backward()
    _check_grads(g, cmap, d, key)
    assert all(p.grad is None for p in vgg.parameters())                        # frozen (Loss.py:26-27)


def test_perception_inside_the_criteria(vgg):
    """CNetLoss (per band, weight 0.4: Demo_USSS.py:40) and CGeneratorLoss (RGB, two feature layers, weight 0.5:
    Demo_WSSS.py:43) with a live perception term; hard mask (generator_mask_switch=True) sends no gradient to cmap."""
    fb.set_precision("parity")
    f = load_golden("perception.pt")
    d = f["cnet"]
    t = d["t"].to(DEV)
    g = d["g"].to(DEV).requires_grad_(True)
    cmap = d["cmap"].to(DEV).requires_grad_(True)
    crit = fb.CNetLoss(channel=3, perception_layer=1, perception_perBand=True, vgg_features=vgg)
    gl, l1, perc, sl = crit(t, g, cmap)
    for got, ref in zip((gl, l1, perc, sl), d["values"]):
        assert _close(got, ref), (float(got), ref)
    (gl + 0.65 * l1 + 0.4 * perc + 0.3 * sl).backward()
    _check_grads(g, cmap, d, "cnet")
    g2 = d["g"].to(DEV).requires_grad_(True)
    cmap2 = d["cmap"].to(DEV).requires_grad_(True)
    crit2 = fb.CGeneratorLoss(channel=3, perception_layer=2, perception_perBand=False, vgg_features=vgg)
    gl2, sl2, perc2 = crit2(t, g2, cmap2)
    for got, ref in zip((gl2, sl2, perc2), f["cgen"]["values"]):
        assert _close(got, ref), (float(got), ref)
    (gl2 + 0.5 * perc2).backward()
    _check_grads(g2, cmap2, f["cgen"], "cgen")
    g3 = d["g"].to(DEV).requires_grad_(True)
    cm3 = d["cmap"].to(DEV).requires_grad_(True)
    _, _, perc3, _ = crit(t, g3, cm3, generator_mask_switch=True)
    assert _close(perc3, f["cnet_hard"]["value"])
    perc3.backward()
    _check_grads(g3, None, f["cnet_hard"], "hard mask")
    assert cm3.grad is None or float(cm3.grad.abs().max()) == 0.0


def test_perception_disabled_without_weights_and_steps_refuse_a_weight():
    """No VGG16 supplied: nothing is downloaded, the term is a constant 0, and the step bodies refuse a non-zero perception
    weight instead of silently training another objective than the reference."""
    crit = fb.CNetLoss(channel=4)
    assert not crit.loss_perception.enabled
    B, C, H, W = 1, 4, 176, 168
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, W, generator=g).to(DEV)
    y = (x.cpu() + 0.3 * torch.randn(B, C, H, W, generator=g)).to(DEV)
    cmap = torch.rand(B, 1, H, W, generator=g).to(DEV)
    assert float(crit(x, y, cmap)[2]) == 0.0
    netG = fb.Generator(C).to(DEV)
    with pytest.raises(ValueError, match="perception_weight"):
        fb.usss_g_step(netG, x, y, crit, perception_weight=0.4)


def test_usss_generator_stage_with_perception_against_the_oracle(vgg):
    """Stage 1 of Demo_USSS (lines 142-159) with the demo's perception weight 0.4: loss values and the generator's parameter
    gradients against the CPU oracle."""
    fb.set_precision("parity")
    C, B, H, W = 4, 1, 176, 168
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, C, H, W, generator=g)
    y = x + 0.3 * torch.randn(B, C, H, W, generator=g)
    sdG = O.make_state_dict(O.generator_spec(C), 11)
    netG = fb.Generator(C)
    netG.load_state_dict(sdG)
    netG.to(DEV).train()
    crit = fb.CNetLoss(channel=C, vgg_features=vgg)
    out = fb.usss_g_step(netG, x.to(DEV), y.to(DEV), crit, perception_weight=0.4, ssim_weight=0.3)
    oG = O.clone_sd(sdG, requires_grad=True)
    y_fake = O.generator(oG, x, train=True)
    zero = torch.zeros(B, 1, H, W)
    gl, l1, sl = O.cnet_loss(y, y_fake, zero)
    vsd = {k: v.cpu() for k, v in vgg.state_dict().items()}
    perc = O.perception_loss(vsd, y, y_fake, zero, 1, True)
    (gl + 0.4 * perc + 0.3 * sl).backward()
    assert _close(out["generator_loss"], gl.item()) and _close(out["perception_loss"], perc.item()) and _close(out["ssim_loss"], sl.item())
    dot = n1 = n2 = 0.0
    for k, p in netG.named_parameters():
        a, b = p.grad.detach().double().cpu().flatten(), oG[k].grad.double().flatten()
        dot += float(a @ b); n1 += float(a @ a); n2 += float(b @ b)
    assert dot / (n1 ** 0.5 * n2 ** 0.5) > 1 - 1e-4 and abs(n1 ** 0.5 / n2 ** 0.5 - 1) < 1e-2
