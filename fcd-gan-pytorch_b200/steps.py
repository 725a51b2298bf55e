"""The training-loop bodies of the reference, as plain functions over the reference-named modules: the
order of forward / zero_grad / backward(retain_graph) / optimizer-step calls is what decides which gradients reach which
network, so it is kept exactly.

    usss_g_step  Demo_USSS.py:142-159   stage 1: generator warm-up (cmap = zeros)
    usss_s_step  Demo_USSS.py:211-228   stage 2: segmentor warm-up (only optimizerS steps)
    usss_step    Demo_USSS.py:305-341   stage 3: joint G + S iteration (CNetLoss, double backward)
    rsss_g_step  Demo_RSSS.py:190-208, Demo_WSSS.py:157-176   generator pre-training (mask = region / zeros)
    rsss_step    Demo_RSSS.py:270-332   D update, then S update through the re-run D (region supervision)
    wsss_step    Demo_WSSS.py:240-323   D update on a changed + an unchanged pair, then S update (nc_loss)

Each returns a dict of the scalar losses (device tensors, no host sync) and the change-density map.  The inline terms
of the reference (`x * (1 - cmap.repeat(...))`, `.mean()`, `torch.mean(abs(cmap))`, `torch.mean(torch.pow(ncmap, 2))`)
are the fused kernels `soft_mask`, `mean`, `mean_abs`, `mean_sq`.  Optimizers are optional (None = gradients only, as
the parity tests use them).  `perception_weight` multiplies `PerceptionLoss` (VGG16 on the conv engine when the criterion was
given `vgg_features`, see losses.PerceptionLoss); the steps REFUSE a non-zero weight when the criterion has no VGG, instead
of silently training another objective than the reference (whose defaults are 0.4 / 0.1 / 0.5, Demo_USSS.py:40,
Demo_RSSS.py:45, Demo_WSSS.py:43 — here the default is 0, the value every BASELINE workload uses).

`lean=True` (usss_step, rsss_step, wsss_step) skips the backward sweeps whose results the reference's own loop body
discards, and nothing else — every forward pass, every loss value, the change-density map, every BatchNorm running statistic
and every gradient that reaches an optimizer step are the same (tests/test_steps_gpu.py::test_lean_steps_*):
  * USSS: `Loss.backward(retain_graph=True)` + `NetLoss.backward()` give G the gradient of Loss TWICE (NetLoss = Loss + w*l1
    and l1 does not depend on G) and S only NetLoss's (its first set is zeroed, Demo_USSS.py:337) — one sweep of NetLoss, then
    G's gradients doubled;
  * RSSS / WSSS: `d_loss.backward()` also back-propagates through the Segmentor, whose gradients are zeroed before they are
    used (Demo_RSSS.py:330, Demo_WSSS.py:321) — the discriminator update runs on masks built from the DETACHED map; the
    generator is frozen after its pre-training (eval mode, never stepped: Demo_RSSS.py:240, Demo_WSSS.py:207) — its forward
    runs without a tape; the discriminator gradients of `s_loss.backward()` are zeroed by the next iteration's
    `optimizerD.zero_grad()` — that pass only carries the gradient THROUGH D (no weight-gradient launches).
SURVEY.md §8(d) counts exactly this algorithmic minimum for configs 3 / 4 / 5 ("reference executes more: 2 S-bwd, 3 full D-bwd,
1 G-bwd").  What differs afterwards is only what the reference never reads: the stale `.grad` of G, and D's until its next
zero_grad.  The default (`lean=False`) replays the reference's sequence call for call.

Data-parallel runs: every body is written once as a generator that YIELDS `(network, wait)` at its exchange points — the
moment a network's gradients are complete, before the optimizer step that consumes them.  `on_grads(network, wait)` (e.g.
`parallel.GradSync.on_grads`) is called there; `graph.YieldingStep` cuts the CUDA graph at the same points.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

from .losses import CGeneratorLoss, CNetLoss, mean, mean_abs, mean_sq, region_loss, soft_mask


def _zero(opt, net):
    if opt is not None:
        opt.zero_grad(set_to_none=True)
    else:
        net.zero_grad()


def drive(gen, on_grads: Optional[Callable] = None):
    """Run a step generator to completion, calling `on_grads(network, wait)` at every exchange point."""
    try:
        while True:
            net, wait = next(gen)
            if on_grads is not None:
                on_grads(net, wait)
    except StopIteration as e:
        return e.value


@contextlib.contextmanager
def _through(net):
    """Inside: `net` only carries gradients THROUGH itself (its parameters take none — the engine then skips every
    weight-gradient launch); the flags are restored afterwards."""
    ps = [p for p in net.parameters() if p.requires_grad]
    for p in ps:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in ps:
            p.requires_grad_(True)


def _check_perception(criterion, weight: float):
    if weight != 0 and not getattr(criterion.loss_perception, "enabled", True):
        raise ValueError("perception_weight != 0 but the criterion was built without VGG16 features (pass vgg_features= to "
                         "CNetLoss / CGeneratorLoss): refusing to train a different objective than the reference")


# ---- stage 1 / 2 of Demo_USSS and the generator pre-training of Demo_RSSS / Demo_WSSS ----------------------------------
def usss_g_gen(netG, x, y, criterion: CNetLoss, optG=None, perception_weight: float = 0.0, ssim_weight: float = 0.0):
    _check_perception(criterion, perception_weight)
    _zero(optG, netG)
    y_fake = netG(x)
    cmap = torch.zeros((x.size(0), 1, x.size(2), x.size(3)), device=x.device)
    gen, l1, perc, ss = criterion(y, y_fake, cmap)
    loss = gen + perception_weight * perc + ssim_weight * ss
    loss.backward()
    yield netG, True
    if optG is not None:
        optG.step()
    return {"generator_loss": gen, "l1_loss": l1, "perception_loss": perc, "ssim_loss": ss, "Loss": loss}


def usss_g_step(netG, x, y, criterion, optG=None, perception_weight=0.0, ssim_weight=0.0, on_grads=None):
    """Demo_USSS.py:142-159 (weights: Demo_USSS.py:40-42)."""
    return drive(usss_g_gen(netG, x, y, criterion, optG, perception_weight, ssim_weight), on_grads)


def usss_s_gen(netG, netS, x, y, criterion: CNetLoss, optS=None, perception_weight: float = 0.0, ssim_weight: float = 0.0,
               l1_weight: float = 0.65):
    _check_perception(criterion, perception_weight)
    y_fake = netG(x)
    cmap = netS(x, y)
    gen, l1, perc, ss = criterion(y, y_fake, cmap)
    net_loss = gen + l1_weight * l1 + perception_weight * perc + ssim_weight * ss
    _zero(optS, netS)
    net_loss.backward()          # G's .grad accumulates unused, like in the reference (only optimizerS steps)
    yield netS, True
    if optS is not None:
        optS.step()
    return {"generator_loss": gen, "l1_loss": l1, "perception_loss": perc, "ssim_loss": ss, "NetLoss": net_loss, "cmap": cmap}


def usss_s_step(netG, netS, x, y, criterion, optS=None, perception_weight=0.0, ssim_weight=0.0, l1_weight=0.65, on_grads=None):
    """Demo_USSS.py:211-228."""
    return drive(usss_s_gen(netG, netS, x, y, criterion, optS, perception_weight, ssim_weight, l1_weight), on_grads)


def rsss_g_gen(netG, x, y, mask, g_criterion: CGeneratorLoss, optG=None, perception_weight: float = 0.0,
               ssim_weight: float = 0.0):
    _check_perception(g_criterion, perception_weight)
    _zero(optG, netG)
    y_fake = netG(x)
    gen, ss, perc = g_criterion(y, y_fake, mask)
    g_loss = gen + perception_weight * perc + ssim_weight * ss
    g_loss.backward()
    yield netG, True
    if optG is not None:
        optG.step()
    return {"generator_loss": gen, "ssim_loss": ss, "perception_loss": perc, "g_loss": g_loss}


def rsss_g_step(netG, x, y, mask, g_criterion, optG=None, perception_weight=0.0, ssim_weight=0.0, on_grads=None):
    """Demo_RSSS.py:190-208 (`mask` = the supervised region) and Demo_WSSS.py:157-176 (`mask` = zeros)."""
    return drive(rsss_g_gen(netG, x, y, mask, g_criterion, optG, perception_weight, ssim_weight), on_grads)


# ---- the joint / adversarial iterations --------------------------------------------------------------------------------
def usss_gen(netG, netS, x, y, criterion: CNetLoss, optG=None, optS=None, perception_weight: float = 0.0,
             ssim_weight: float = 0.0, l1_weight: float = 0.65, lean: bool = False):
    _check_perception(criterion, perception_weight)
    y_fake = netG(x)
    cmap = netS(x, y)
    gen, l1, perc, ss = criterion(y, y_fake, cmap)
    loss = gen + perception_weight * perc + ssim_weight * ss
    net_loss = loss + l1_weight * l1
    if lean:
        _zero(optG, netG)
        _zero(optS, netS)
        net_loss.backward()             # d NetLoss / d G  ==  d Loss / d G  (l1 does not depend on G)
        torch._foreach_mul_([p.grad for p in netG.parameters() if p.grad is not None], 2.0)
    else:
        _zero(optG, netG)
        loss.backward(retain_graph=True)
        _zero(optS, netS)
        net_loss.backward()
    yield netG, False
    yield netS, True
    if optG is not None:
        optG.step()
    if optS is not None:
        optS.step()
    return {"generator_loss": gen, "l1_loss": l1, "ssim_loss": ss, "perception_loss": perc, "Loss": loss,
            "NetLoss": net_loss, "cmap": cmap}


def usss_step(netG, netS, x, y, criterion: CNetLoss, optG=None, optS=None, perception_weight: float = 0.0,
              ssim_weight: float = 0.0, l1_weight: float = 0.65, on_grads=None, lean: bool = False) -> Dict[str, torch.Tensor]:
    """Demo_USSS.py:320-341 (weights: Demo_USSS.py:40-42).  G receives d(Loss) + d(NetLoss) (two backward sweeps over
    the same graph, `.grad` accumulation), S only d(NetLoss) because its gradients are zeroed in between."""
    return drive(usss_gen(netG, netS, x, y, criterion, optG, optS, perception_weight, ssim_weight, l1_weight, lean), on_grads)


def _d_update(netD, optD, c_out, nc_out, retain: bool = True):
    _zero(optD, netD)
    d_loss = 1 + mean(nc_out) - mean(c_out)                  # Demo_RSSS.py:303, Demo_WSSS.py:283
    d_loss.backward(retain_graph=retain)
    return d_loss


def _generator_term(netG, x, y, cmap, g_criterion, g_weight, perception_weight, ssim_weight, lean):
    if g_weight == 0:
        return torch.zeros((), device=x.device)
    if lean:                            # G is frozen in eval mode from the adversarial stage on: no tape, no backward sweep
        with torch.no_grad():
            y_fake = netG(x)
    else:
        y_fake = netG(x)
    gen, ss, perc = g_criterion(y, y_fake, cmap)
    return gen + perception_weight * perc + ssim_weight * ss


def rsss_gen(netG, netS, netD, x, y, region, g_criterion: CGeneratorLoss, optS=None, optD=None, d_weight: float = 1.0,
             l1_weight: float = 0.02, g_weight: float = 0.5, r_weight: float = 2.0, perception_weight: float = 0.0,
             ssim_weight: float = 0.0, lean: bool = False):
    _check_perception(g_criterion, perception_weight)
    cmap = netS(x, y)
    x_mask = soft_mask(x, cmap)
    y_mask = soft_mask(y, cmap)
    if lean:                            # the D update on masks built from the detached map: no sweep through S
        cd = cmap.detach()
        xm_d = soft_mask(x, cd)
        c_out = netD(xm_d, soft_mask(y, cd))
        nc_out = netD(xm_d, soft_mask(y, cd, other=x, region=region))
    else:
        c_out = netD(x_mask, y_mask)
        y_unc = soft_mask(y, cmap, other=x, region=region)   # (y*(1-region) + x*region) * (1-cmap), Demo_RSSS.py:297-300
        nc_out = netD(x_mask, y_unc)
    d_loss = _d_update(netD, optD, c_out, nc_out, retain=not lean)
    yield netD, True
    if optD is not None:
        optD.step()
    if lean:
        with _through(netD):
            c_out = netD(x_mask, y_mask)
    else:
        c_out = netD(x_mask, y_mask)                         # rebuilt with the updated D, Demo_RSSS.py:311
    g_loss = _generator_term(netG, x, y, cmap, g_criterion, g_weight, perception_weight, ssim_weight, lean)
    l1 = region_loss(cmap, region, nn.L1Loss())
    r = region_loss(cmap, 1 - region, nn.MSELoss())
    s_d = mean(c_out)
    s_loss = d_weight * s_d + l1_weight * l1 + g_weight * g_loss + r_weight * r
    _zero(optS, netS)
    s_loss.backward()
    yield netS, True
    if optS is not None:
        optS.step()
    return {"d_loss": d_loss, "s_d_loss": s_d, "g_loss": g_loss, "l1_loss": l1, "r_loss": r, "s_loss": s_loss, "cmap": cmap}


def rsss_step(netG, netS, netD, x, y, region, g_criterion: CGeneratorLoss, optS=None, optD=None, d_weight: float = 1.0,
              l1_weight: float = 0.02, g_weight: float = 0.5, r_weight: float = 2.0, perception_weight: float = 0.0,
              ssim_weight: float = 0.0, on_grads=None, lean: bool = False) -> Dict[str, torch.Tensor]:
    """Demo_RSSS.py:285-331 with discriminator_continuous=True (weights: Demo_RSSS.py:45-53)."""
    return drive(rsss_gen(netG, netS, netD, x, y, region, g_criterion, optS, optD, d_weight, l1_weight, g_weight, r_weight,
                          perception_weight, ssim_weight, lean), on_grads)


def wsss_gen(netG, netS, netD, x, y, x_nc, y_nc, g_criterion: CGeneratorLoss, optS=None, optD=None, d_weight: float = 1.0,
             l1_weight: float = 1.6, g_weight: float = 0.2, nc_weight: float = 1.5, perception_weight: float = 0.0,
             ssim_weight: float = 0.0, lean: bool = False):
    _check_perception(g_criterion, perception_weight)
    cmap = netS(x, y)
    x_mask = soft_mask(x, cmap)
    y_mask = soft_mask(y, cmap)
    ncmap = netS(x_nc, y_nc)
    if lean:
        cd = cmap.detach()
        c_out = netD(soft_mask(x, cd), soft_mask(y, cd))
        nc_out = netD(soft_mask(x_nc, cd), soft_mask(y_nc, cd))
    else:
        c_out = netD(x_mask, y_mask)
        nc_out = netD(soft_mask(x_nc, cmap), soft_mask(y_nc, cmap))
    d_loss = _d_update(netD, optD, c_out, nc_out, retain=not lean)
    yield netD, True
    if optD is not None:
        optD.step()
    nc_loss = mean_sq(ncmap)
    if lean:
        with _through(netD):
            c_out = netD(x_mask, y_mask)
    else:
        c_out = netD(x_mask, y_mask)
    g_loss = _generator_term(netG, x, y, cmap, g_criterion, g_weight, perception_weight, ssim_weight, lean)
    l1 = mean_abs(cmap)
    s_d = mean(c_out)
    s_loss = d_weight * s_d + l1_weight * l1 + g_weight * g_loss + nc_weight * nc_loss
    _zero(optS, netS)
    s_loss.backward()
    yield netS, True
    if optS is not None:
        optS.step()
    return {"d_loss": d_loss, "s_d_loss": s_d, "g_loss": g_loss, "l1_loss": l1, "nc_loss": nc_loss, "s_loss": s_loss,
            "cmap": cmap, "ncmap": ncmap}


def wsss_step(netG, netS, netD, x, y, x_nc, y_nc, g_criterion: CGeneratorLoss, optS=None, optD=None,
              d_weight: float = 1.0, l1_weight: float = 1.6, g_weight: float = 0.2, nc_weight: float = 1.5,
              perception_weight: float = 0.0, ssim_weight: float = 0.0, on_grads=None, lean: bool = False) -> Dict[str, torch.Tensor]:
    """Demo_WSSS.py:247-319 with discriminator_continuous=True (weights: Demo_WSSS.py:43-52).  The unchanged pair is
    masked with the CHANGED pair's map (Demo_WSSS.py:276-277)."""
    return drive(wsss_gen(netG, netS, netD, x, y, x_nc, y_nc, g_criterion, optS, optD, d_weight, l1_weight, g_weight,
                          nc_weight, perception_weight, ssim_weight, lean), on_grads)
