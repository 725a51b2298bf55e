"""The three training-loop bodies of the reference, as plain functions over the reference-named
modules: the order of forward / zero_grad / backward(retain_graph) / optimizer-step calls is what decides which
gradients reach which network, so it is kept exactly.

    usss_step   Demo_USSS.py:305-341   joint G + S iteration (CNetLoss, double backward)
    rsss_step   Demo_RSSS.py:270-332   D update, then S update through the re-run D (region supervision)
    wsss_step   Demo_WSSS.py:240-323   D update on a changed + an unchanged pair, then S update (nc_loss)

Each returns a dict of the scalar losses (device tensors, no host sync) and the change-density map.  The inline terms
of the reference (`x * (1 - cmap.repeat(...))`, `.mean()`, `torch.mean(abs(cmap))`, `torch.mean(torch.pow(ncmap, 2))`)
are the fused kernels `soft_mask`, `mean`, `mean_abs`, `mean_sq`.  Optimizers are optional (None = gradients only, as
the parity tests use them).  The perception term is out of scope: its weight multiplies the passthrough
`PerceptionLoss`, 0 by default.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .losses import CGeneratorLoss, CNetLoss, mean, mean_abs, mean_sq, region_loss, soft_mask


def _zero(opt, net):
    if opt is not None:
        opt.zero_grad(set_to_none=True)
    else:
        net.zero_grad()


def usss_step(netG, netS, x, y, criterion: CNetLoss, optG=None, optS=None, perception_weight: float = 0.0,
              ssim_weight: float = 0.0, l1_weight: float = 0.65) -> Dict[str, torch.Tensor]:
    """Demo_USSS.py:320-341 (weights: Demo_USSS.py:40-42).  G receives d(Loss) + d(NetLoss) (two backward sweeps over
    the same graph, `.grad` accumulation), S only d(NetLoss) because its gradients are zeroed in between."""
    y_fake = netG(x)
    cmap = netS(x, y)
    gen, l1, perc, ss = criterion(y, y_fake, cmap)
    loss = gen + perception_weight * perc + ssim_weight * ss
    _zero(optG, netG)
    loss.backward(retain_graph=True)
    net_loss = loss + l1_weight * l1
    _zero(optS, netS)
    net_loss.backward()
    if optG is not None:
        optG.step()
    if optS is not None:
        optS.step()
    return {"generator_loss": gen, "l1_loss": l1, "ssim_loss": ss, "perception_loss": perc, "Loss": loss,
            "NetLoss": net_loss, "cmap": cmap}


def _d_update(netD, optD, c_out, nc_out):
    _zero(optD, netD)
    d_loss = 1 + mean(nc_out) - mean(c_out)                  # Demo_RSSS.py:303, Demo_WSSS.py:283
    d_loss.backward(retain_graph=True)
    if optD is not None:
        optD.step()
    return d_loss


def rsss_step(netG, netS, netD, x, y, region, g_criterion: CGeneratorLoss, optS=None, optD=None, d_weight: float = 1.0,
              l1_weight: float = 0.02, g_weight: float = 0.5, r_weight: float = 2.0, perception_weight: float = 0.0,
              ssim_weight: float = 0.0) -> Dict[str, torch.Tensor]:
    """Demo_RSSS.py:285-331 with discriminator_continuous=True (weights: Demo_RSSS.py:45-53)."""
    cmap = netS(x, y)
    x_mask = soft_mask(x, cmap)
    y_mask = soft_mask(y, cmap)
    c_out = netD(x_mask, y_mask)
    y_unc = soft_mask(y, cmap, other=x, region=region)       # (y*(1-region) + x*region) * (1-cmap), Demo_RSSS.py:297-300
    nc_out = netD(x_mask, y_unc)
    d_loss = _d_update(netD, optD, c_out, nc_out)
    c_out = netD(x_mask, y_mask)                             # rebuilt with the updated D, Demo_RSSS.py:311
    if g_weight != 0:
        y_fake = netG(x)
        gen, ss, perc = g_criterion(y, y_fake, cmap)
        g_loss = gen + perception_weight * perc + ssim_weight * ss
    else:
        g_loss = torch.zeros((), device=x.device)
    l1 = region_loss(cmap, region, nn.L1Loss())
    r = region_loss(cmap, 1 - region, nn.MSELoss())
    s_d = mean(c_out)
    s_loss = d_weight * s_d + l1_weight * l1 + g_weight * g_loss + r_weight * r
    _zero(optS, netS)
    s_loss.backward()
    if optS is not None:
        optS.step()
    return {"d_loss": d_loss, "s_d_loss": s_d, "g_loss": g_loss, "l1_loss": l1, "r_loss": r, "s_loss": s_loss, "cmap": cmap}


def wsss_step(netG, netS, netD, x, y, x_nc, y_nc, g_criterion: CGeneratorLoss, optS=None, optD=None,
              d_weight: float = 1.0, l1_weight: float = 1.6, g_weight: float = 0.2, nc_weight: float = 1.5,
              perception_weight: float = 0.0, ssim_weight: float = 0.0) -> Dict[str, torch.Tensor]:
    """Demo_WSSS.py:247-319 with discriminator_continuous=True (weights: Demo_WSSS.py:43-52).  The unchanged pair is
    masked with the CHANGED pair's map (Demo_WSSS.py:276-277)."""
    cmap = netS(x, y)
    x_mask = soft_mask(x, cmap)
    y_mask = soft_mask(y, cmap)
    c_out = netD(x_mask, y_mask)
    ncmap = netS(x_nc, y_nc)
    nc_out = netD(soft_mask(x_nc, cmap), soft_mask(y_nc, cmap))
    d_loss = _d_update(netD, optD, c_out, nc_out)
    nc_loss = mean_sq(ncmap)
    c_out = netD(x_mask, y_mask)
    if g_weight != 0:
        y_fake = netG(x)
        gen, ss, perc = g_criterion(y, y_fake, cmap)
        g_loss = gen + perception_weight * perc + ssim_weight * ss
    else:
        g_loss = torch.zeros((), device=x.device)
    l1 = mean_abs(cmap)
    s_d = mean(c_out)
    s_loss = d_weight * s_d + l1_weight * l1 + g_weight * g_loss + nc_weight * nc_loss
    _zero(optS, netS)
    s_loss.backward()
    if optS is not None:
        optS.step()
    return {"d_loss": d_loss, "s_d_loss": s_d, "g_loss": g_loss, "l1_loss": l1, "nc_loss": nc_loss, "s_loss": s_loss,
            "cmap": cmap, "ncmap": ncmap}
