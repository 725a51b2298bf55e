"""GPU parity of the training-loop bodies (the callers of the hot path) against golden vectors recorded from
the UNMODIFIED reference running the same loop bodies (oracle/make_golden.py):
  * USSS joint iteration, Demo_USSS.py:305-341 (double backward with retain_graph; live SSIM gradient),
  * RSSS adversarial iteration, Demo_RSSS.py:285-331 (D update with gradients flowing into S through the soft
    masks, D re-run, CGeneratorLoss + region losses).
Loss values within 2e-4 relative, change-density map within 1e-3, gradients: global cosine / norm ratio and
per-tensor L2 (activation-kink tolerant, see tests/_util.check_grad_summary_l2)."""
import pytest
import torch
import torch.nn as nn

import fcdgan_b200 as fb
from oracle import fcd_oracle as O
from tests._util import check_grad_summary_l2, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nets(C):
    G = fb.Generator(C); G.load_state_dict(O.make_state_dict(O.generator_spec(C), 11))
    S = fb.Segmentor(C, 1, True); S.load_state_dict(O.make_state_dict(O.segmentor_spec(C, 1, True), 12))
    D = fb.Discriminator_SRGAN_simple(C); D.load_state_dict(O.make_state_dict(O.discriminator_spec(C), 13))
    return G.to(DEV), S.to(DEV), D.to(DEV)


def _grads(net):
    return {k: p.grad for k, p in net.named_parameters()}


def test_step_usss():
    f = load_golden("step_usss.pt")
    C = f["C"]
    netG, netS, _ = _nets(C)
    netG.train(); netS.train()
    x, y = f["x"].to(DEV), f["y"].to(DEV)
    crit = fb.CNetLoss(channel=C)
    ssim_w, l1_w = f["ssim_w"], f["l1_w"]
    # --- Demo_USSS.py:320-338 verbatim (perception weight 0)
    y_fake = netG(x)
    cmap = netS(x, y)
    gl, l1, _, sl = crit(y, y_fake, cmap)
    Loss = gl + ssim_w * sl
    netG.zero_grad()
    Loss.backward(retain_graph=True)
    NetLoss = gl + l1_w * l1 + ssim_w * sl
    netS.zero_grad()
    NetLoss.backward()
    for got, ref in zip((gl, l1, sl), f["losses"]):
        assert abs(got.item() - ref) <= 2e-4 * max(abs(ref), 1e-3), (got.item(), ref)
    assert rel_err(cmap, f["cmap"]) < 1e-3
    check_grad_summary_l2(_grads(netG), f["gradsG"], 1e-2, 0.25, what="usss G")
    check_grad_summary_l2(_grads(netS), f["gradsS"], 5e-2, 0.5, what="usss S")


def test_step_rsss():
    f = load_golden("step_rsss.pt")
    C = f["C"]
    netG, netS, netD = _nets(C)
    netG.eval(); netS.train(); netD.train()
    x, y, region = f["x"].to(DEV), f["y"].to(DEV), f["region"].to(DEV)
    # --- Demo_RSSS.py:285-331 with the fused inline terms (soft_mask / mean)
    cmap = netS(x, y)
    x_mask = fb.soft_mask(x, cmap)
    y_mask = fb.soft_mask(y, cmap)
    c_out = netD(x_mask, y_mask)
    x_unc = fb.soft_mask(x, cmap)
    y_unc = fb.soft_mask(y, cmap, other=x, region=region)
    nc_out = netD(x_unc, y_unc)
    netD.zero_grad()
    d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
    d_loss.backward(retain_graph=True)
    assert abs(d_loss.item() - f["d_loss"]) < 2e-4
    assert rel_err(c_out, f["c_out"]) < 1e-3 and rel_err(nc_out, f["nc_out"]) < 1e-3
    assert rel_err(cmap, f["cmap"]) < 1e-3
    check_grad_summary_l2(_grads(netD), f["gradsD"], 5e-2, 0.5, what="rsss D")
    c_out2 = netD(x_mask, y_mask)
    y_fake = netG(x)
    gcrit = fb.CGeneratorLoss(channel=C)
    gl, sl, _ = gcrit(y, y_fake, cmap)
    g_loss = gl + 0.0 * sl
    rl1 = fb.region_loss(cmap, region, nn.L1Loss())
    rl2 = fb.region_loss(cmap, 1 - region, nn.MSELoss())
    s_loss = 1.0 * fb.mean(c_out2) + 0.02 * rl1 + 0.5 * g_loss + 2.0 * rl2
    netS.zero_grad()
    s_loss.backward()
    assert abs(s_loss.item() - f["s_loss"]) < 2e-4 * max(1.0, abs(f["s_loss"]))
    check_grad_summary_l2(_grads(netS), f["gradsS"], 5e-2, 0.5, what="rsss S")
