"""The reference's training-loop bodies over the UNMODIFIED reference classes — BASELINE / TEST INFRASTRUCTURE ONLY.

`bench.py --impl reference` (host CPU), `bench.py --impl cudnn` / `gpu_baseline` (the same classes moved to the B200 with
`.cuda()`: torch + cuDNN, what the reference itself runs on a GPU) and the step parity tests drive these.  The classes come
from oracle/ref_import.py (mounted reference or the byte-identical staged copy oracle/_ref/); the loop bodies are the demos'
own statements (the demo scripts cannot be imported: everything is under `__main__` with hard-coded paths):

    gd_step     BASELINE configs[1]: generator iteration Demo_USSS.py:142-159 + discriminator update Demo_RSSS.py:285-307
    g_step      north-star configuration: generator iteration only
    usss_step   Demo_USSS.py:305-341
    rsss_step   Demo_RSSS.py:285-331
    wsss_step   Demo_WSSS.py:247-319

The perception term has weight 0 in every BASELINE workload; `Loss.CNetLoss.forward` would still run its
VGG16 on every band, so `stub_perception` replaces the criterion's `loss_perception` attribute by a constant 0 — this REMOVES
work from the reference arm (conservative for any speed-up quoted against it) and is said so in bench.py's JSON line.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _ZeroPerception(nn.Module):
    def forward(self, target_image, generate_image, cmask):
        return 0


def stub_perception(criterion):
    criterion.loss_perception = _ZeroPerception()
    return criterion


def masked_l1(target_image, generate_image, cmap):
    """The generator term of CNetLoss.forward, Loss.py:76-84, on its own (configs[1] / north star use no other term)."""
    num_pixel = target_image.size()[2] * target_image.size()[3]
    num_wnc = torch.sum(1 - cmap, (1, 2, 3))
    target_image_mask = target_image * (1 - cmap.repeat((1, target_image.size()[1], 1, 1)))
    generate_image_mask = generate_image * (1 - cmap.repeat((1, generate_image.size()[1], 1, 1)))
    loss_generator = nn.L1Loss()
    generator_loss = 0
    for i in range(target_image.shape[0]):
        generator_loss += loss_generator(target_image_mask[i], generate_image_mask[i]) * num_pixel / num_wnc[i]
    return generator_loss / target_image.shape[0]


def g_step(netG, optG, x, y, zero_cmap):
    """Demo_USSS.py:142-159 with perception / ssim weight 0 and cmap = zeros (Demo_USSS.py:151)."""
    y_fake = netG(x)
    gen_loss = masked_l1(y, y_fake, zero_cmap)
    optG.zero_grad()
    gen_loss.backward()
    optG.step()
    return gen_loss


def gd_step(netG, netD, optG, optD, x, y, region, cmap, zero_cmap):
    """configs[1]: the generator iteration, then the discriminator update of Demo_RSSS.py:285-307 on a given density map."""
    y_fake = netG(x)
    gen_loss = masked_l1(y, y_fake, zero_cmap)
    optG.zero_grad()
    gen_loss.backward()
    C = x.shape[1]
    x_mask = x * (1 - cmap.repeat((1, C, 1, 1)))
    y_mask = y * (1 - cmap.repeat((1, C, 1, 1)))
    c_out = netD(x_mask, y_mask)
    y_unc = y * (1 - region) + x * region
    y_unc = y_unc * (1 - cmap.repeat((1, C, 1, 1)))
    nc_out = netD(x_mask, y_unc)
    optD.zero_grad()
    d_loss = 1 + nc_out.mean() - c_out.mean()
    d_loss.backward()
    optG.step()
    optD.step()
    return gen_loss, d_loss


def usss_step(netG, netS, criterion, optG, optS, x, y, perception_weight=0.0, ssim_weight=0.0, l1_weight=0.65):
    """Demo_USSS.py:320-341."""
    y_fake = netG(x)
    cmap = netS(x, y)
    generator_loss, l1_loss, perception_loss, ssim_loss = criterion(y, y_fake, cmap)
    Loss = generator_loss + perception_weight * perception_loss + ssim_weight * ssim_loss
    optG.zero_grad()
    Loss.backward(retain_graph=True)
    NetLoss = generator_loss + l1_weight * l1_loss + perception_weight * perception_loss + ssim_weight * ssim_loss
    optS.zero_grad()
    NetLoss.backward()
    optG.step()
    optS.step()
    return Loss, NetLoss


def rsss_step(netG, netS, netD, g_criterion, region_loss, optS, optD, x, y, region, d_weight=1.0, l1_weight=0.02,
              g_weight=0.5, r_weight=2.0, perception_weight=0.0, ssim_weight=0.0):
    """Demo_RSSS.py:285-331 (discriminator_continuous = True)."""
    C = x.shape[1]
    cmap = netS(x, y)
    cmask = cmap
    x_mask = x * (1 - cmask.repeat((1, C, 1, 1)))
    y_mask = y * (1 - cmask.repeat((1, C, 1, 1)))
    c_out = netD(x_mask, y_mask)
    x_unc = x
    y_unc = y * (1 - region) + x * region
    x_unc = x_unc * (1 - cmask.repeat((1, C, 1, 1)))
    y_unc = y_unc * (1 - cmask.repeat((1, C, 1, 1)))
    nc_out = netD(x_unc, y_unc)
    optD.zero_grad()
    d_loss = 1 + nc_out.mean() - c_out.mean()
    d_loss.backward(retain_graph=True)
    optD.step()
    c_out = netD(x_mask, y_mask)
    y_fake = netG(x)
    generator_loss, ssim_loss, perception_loss = g_criterion(y, y_fake, cmap)
    g_loss = generator_loss + perception_weight * perception_loss + ssim_weight * ssim_loss
    l1_loss = region_loss(cmap, region, nn.L1Loss())
    r_loss = region_loss(cmap, 1 - region, nn.MSELoss())
    s_d_loss = c_out.mean()
    s_loss = d_weight * s_d_loss + l1_weight * l1_loss + g_weight * g_loss + r_weight * r_loss
    optS.zero_grad()
    s_loss.backward()
    optS.step()
    return d_loss, s_loss


def wsss_step(netG, netS, netD, g_criterion, optS, optD, x, y, x_nc, y_nc, d_weight=1.0, l1_weight=1.6, g_weight=0.2,
              nc_weight=1.5, perception_weight=0.0, ssim_weight=0.0):
    """Demo_WSSS.py:247-319 (discriminator_continuous = True)."""
    C = x.shape[1]
    cmap = netS(x, y)
    cmask = cmap
    x_mask = x * (1 - cmask.repeat((1, C, 1, 1)))
    y_mask = y * (1 - cmask.repeat((1, C, 1, 1)))
    c_out = netD(x_mask, y_mask)
    ncmap = netS(x_nc, y_nc)
    x_mask_nc = x_nc * (1 - cmask.repeat((1, C, 1, 1)))
    y_mask_nc = y_nc * (1 - cmask.repeat((1, C, 1, 1)))
    nc_out = netD(x_mask_nc, y_mask_nc)
    optD.zero_grad()
    d_loss = 1 + nc_out.mean() - c_out.mean()
    d_loss.backward(retain_graph=True)
    optD.step()
    nc_loss = torch.mean(torch.pow(ncmap, 2))
    c_out = netD(x_mask, y_mask)
    y_fake = netG(x)
    generator_loss, ssim_loss, perception_loss = g_criterion(y, y_fake, cmap)
    g_loss = generator_loss + perception_weight * perception_loss + ssim_weight * ssim_loss
    l1_loss = torch.mean(abs(cmap))
    s_d_loss = c_out.mean()
    s_loss = d_weight * s_d_loss + l1_weight * l1_loss + g_weight * g_loss + nc_weight * nc_loss
    optS.zero_grad()
    s_loss.backward()
    optS.step()
    return d_loss, s_loss
