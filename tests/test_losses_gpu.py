"""GPU parity of the fused loss kernels (csrc/losses.cu through fcdgan_b200.losses / .ssim) against the golden
vectors produced by the UNMODIFIED reference Loss.py / ssim.py (tests/golden/losses.pt), plus live checks against
the CPU oracle at other shapes (odd sizes, win skipping, per-image reduction).

Tolerance: 1e-4 relative on loss values, 2e-4 of the tensor max on gradients (fp32 arithmetic, different
summation order than oneDNN)."""
import pytest
import torch
import torch.nn as nn

import fcdgan_b200 as fb
from oracle import fcd_oracle as O
from tests._util import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VTOL = 1e-4
GTOL = 2e-4


def close(a, ref, tol=VTOL):
    a = a.item() if torch.is_tensor(a) else a
    return abs(a - ref) <= tol * max(abs(ref), 1e-3)


def test_cnet_loss_golden():
    f = load_golden("losses.pt")
    t = f["t"].to(DEV)
    g = f["g"].to(DEV).requires_grad_(True)
    cmap = f["cmap"].to(DEV).requires_grad_(True)
    crit = fb.CNetLoss(channel=3)
    gl, l1, perc, sl = crit(t, g, cmap)
    assert float(perc) == 0.0 or crit.loss_perception.enabled
    for got, ref in zip((gl, l1, sl), f["cnet"]):
        assert close(got, ref), (got.item(), ref)
    (gl + 0.65 * l1 + 0.7 * sl).backward()
    assert rel_err(g.grad, f["cnet_dg"]) < GTOL and rel_err(cmap.grad, f["cnet_dcmap"]) < GTOL


def test_cgenerator_loss_golden_with_skipped_sample():
    f = load_golden("losses.pt")
    t = f["t"].to(DEV)
    g = f["g"].to(DEV).requires_grad_(True)
    cmap2 = f["cmap2"].to(DEV).requires_grad_(True)     # sample 1 is all-changed: sum(1-cmap) == 0 -> skipped
    crit = fb.CGeneratorLoss(channel=3)
    gl, sl, perc = crit(t, g, cmap2)
    for got, ref in zip((gl, sl), f["cgen"]):
        assert close(got, ref), (got.item(), ref)
    (gl + 0.3 * sl).backward()
    assert torch.isfinite(g.grad).all() and torch.isfinite(cmap2.grad).all()
    assert rel_err(g.grad, f["cgen_dg"]) < GTOL and rel_err(cmap2.grad, f["cgen_dcmap"]) < GTOL


def test_region_loss_golden_with_empty_region():
    f = load_golden("losses.pt")
    cm = f["cmap"].to(DEV).requires_grad_(True)
    region = f["region"].to(DEV)
    r1 = fb.region_loss(cm, region, nn.L1Loss())
    r2 = fb.region_loss(cm, 1 - region, nn.MSELoss())
    assert close(r1, f["region_l1"]) and close(r2, f["region_mse"])
    (0.02 * r1 + 2 * r2).backward()
    assert rel_err(cm.grad, f["region_dcmap"]) < GTOL
    with pytest.raises(NotImplementedError):
        fb.region_loss(cm, region, nn.BCELoss())


def test_ssim_family_golden():
    f = load_golden("losses.pt")
    X, Y = f["X"].to(DEV), f["Y"].to(DEV)
    assert close(fb.SSIM(data_range=1.0, channel=3)(X, Y), f["ssim"])
    assert rel_err(fb.SSIM(data_range=1.0, channel=3, size_average=False, nonnegative_ssim=True)(X, Y), f["ssim_nsa"]) < VTOL
    assert close(fb.MS_SSIM(data_range=1.0, channel=3)(X, Y), f["msssim"])
    assert rel_err(fb.MS_SSIM(data_range=1.0, channel=3, size_average=False)(X, Y), f["msssim_nsa"]) < VTOL
    assert close(fb.MS_SSIM(data_range=1.0, channel=3)(X, 1 - X), f["msssim_anti"], 1e-3)
    with pytest.raises(ValueError):
        fb.ms_ssim(X, Y[:, :2], data_range=1.0)
    with pytest.raises(AssertionError):
        fb.ms_ssim(X[..., :160, :160], Y[..., :160, :160], data_range=1.0)
    with pytest.raises(ValueError):
        fb.ssim(X, Y, data_range=1.0, win_size=10)


@pytest.mark.parametrize("shape", [(2, 13, 176, 200), (1, 4, 220, 220), (2, 3, 161, 163)])
def test_ms_ssim_grad_vs_oracle(shape):
    """values and gradients at odd pyramid sizes (avg_pool2d padding = size % 2) vs the oracle's autograd."""
    g = torch.Generator().manual_seed(3)
    X = torch.randn(shape, generator=g)
    Y = X + 0.3 * torch.randn(shape, generator=g)
    Xo, Yo = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
    ref = O.ms_ssim(Xo, Yo, data_range=1.0)
    ref.backward()
    Xc, Yc = X.to(DEV).requires_grad_(True), Y.to(DEV).requires_grad_(True)
    got = fb.ms_ssim(Xc, Yc, data_range=1.0)
    assert close(got, ref.item())
    got.backward()
    assert rel_err(Xc.grad, Xo.grad) < GTOL and rel_err(Yc.grad, Yo.grad) < GTOL


def test_ssim_grad_small_and_skipped_dimension():
    """single-scale SSIM incl. a dimension smaller than the window (blur skipped along it, ssim.py:45-50)."""
    g = torch.Generator().manual_seed(4)
    for shape in [(2, 3, 40, 37), (1, 2, 8, 30)]:
        X = torch.rand(shape, generator=g)
        Y = (X + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
        Xo, Yo = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.ssim(Xo, Yo, data_range=1.0, size_average=False)
            (ref * torch.arange(1, shape[0] + 1)).sum().backward()
            Xc, Yc = X.to(DEV).requires_grad_(True), Y.to(DEV).requires_grad_(True)
            got = fb.ssim(Xc, Yc, data_range=1.0, size_average=False)
        assert rel_err(got, ref) < VTOL
        (got * torch.arange(1, shape[0] + 1, device=DEV)).sum().backward()
        assert rel_err(Xc.grad, Xo.grad) < GTOL and rel_err(Yc.grad, Yo.grad) < GTOL


def test_inline_terms():
    """fused mean / mean|x| / mean x^2 and soft masking (Demo_RSSS.py:290-300, Demo_WSSS.py:299,315)."""
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(2, 4, 20, 24, generator=g), torch.randn(2, 4, 20, 24, generator=g)
    region = (torch.rand(2, 1, 20, 24, generator=g) > 0.5).float()
    cm = torch.rand(2, 1, 20, 24, generator=g)
    cmo = cm.clone().requires_grad_(True)
    xm = x * (1 - cmo.repeat(1, 4, 1, 1))
    yu = (y * (1 - region) + x * region) * (1 - cmo.repeat(1, 4, 1, 1))
    ref = (xm * y).sum() + (yu * x).sum() + 0.3 * cmo.abs().mean() + 0.7 * (cmo ** 2).mean() + cmo.mean()
    ref.backward()
    cmc = cm.to(DEV).requires_grad_(True)
    xd, yd, rd = x.to(DEV), y.to(DEV), region.to(DEV)
    xm_c = fb.soft_mask(xd, cmc)
    yu_c = fb.soft_mask(yd, cmc, other=xd, region=rd)
    assert rel_err(xm_c, xm) < 1e-6 and rel_err(yu_c, yu) < 1e-6
    got = (xm_c * yd).sum() + (yu_c * xd).sum() + 0.3 * fb.mean_abs(cmc) + 0.7 * fb.mean_sq(cmc) + fb.mean(cmc)
    assert close(got, ref.item())
    got.backward()
    assert rel_err(cmc.grad, cmo.grad) < 1e-5


def test_cpu_input_fails_loudly():
    from fcdgan_b200._lib import FcdError
    with pytest.raises(FcdError):
        fb.ms_ssim(torch.rand(1, 1, 200, 200), torch.rand(1, 1, 200, 200), data_range=1.0)
    with pytest.raises(FcdError):
        fb.Generator(3)(torch.rand(1, 3, 16, 16))
