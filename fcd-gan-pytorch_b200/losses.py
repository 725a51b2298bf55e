"""Loss stack of the reference's Loss.py with the same class names / signatures / return orders, on fused
sm_100a kernels (csrc/losses.cu):

    CNetLoss(channel=4, perception_layer=1, perception_perBand=True)(target, generate, cmap, generator_mask_switch=False)
        -> (generator_loss, l1_loss, perception_loss, ssim_loss)                      Loss.py:64-95
    CGeneratorLoss(channel=3, perception_layer=1, perception_perBand=False)(target, generate, cmap)
        -> (generator_loss, ssim_loss, perception_loss)                               Loss.py:100-124
    region_loss(cmap, region, criterion)                                              Loss.py:127-141

plus fused forms of the inline terms of the training loops (`mean_abs`, `mean_sq`, `mean`, `soft_mask`).

The reference loops over the batch in Python with a host sync per sample (Loss.py:82-84,115-118,135-138) and
materialises `cmap.repeat(...)`; here one kernel reads (target, generate, cmap) once and produces the per-sample
sums, mean|cmap| and the masked images for MS-SSIM, and one backward kernel produces every gradient.

PerceptionLoss (VGG16, Loss.py:17-61) is OUT OF SCOPE for the CUDA path (its ImageNet weights
cannot be downloaded here and it is not in the north star): it is a plain-PyTorch passthrough used only when
weights are available; otherwise the perception term is a constant 0 and a warning is issued once.
"""
from __future__ import annotations

import warnings
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .engine import _call
from .ssim import MS_SSIM

LOSS_L1, LOSS_MSE = 0, 1


def _check(t: torch.Tensor, what: str):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise _lib.FcdError(f"{what}: fcdgan_b200 losses take fp32 CUDA tensors (there is no CPU path)")


class _MaskedRecon(torch.autograd.Function):
    """(generator_loss, l1_loss, target*(1-cmap), generate*(1-cmap)) in one pass; Loss.py:75-87 / 109-119."""

    @staticmethod
    def forward(ctx, target, generate, cmap, kind: int, want_masked: bool):
        for t in (target, generate, cmap):
            _check(t, "masked reconstruction loss")
        B, C, H, W = target.shape
        if generate.shape != target.shape or tuple(cmap.shape) != (B, 1, H, W):
            raise ValueError(f"loss: expected generate {tuple(target.shape)} and cmap {(B, 1, H, W)}, got "
                             f"{tuple(generate.shape)} and {tuple(cmap.shape)}")
        target, generate, cmap = target.contiguous(), generate.contiguous(), cmap.contiguous()
        dev = target.device
        sums = torch.empty(3 * B, dtype=torch.float64, device=dev)
        out2 = torch.empty(2, dtype=torch.float32, device=dev)
        tm = torch.empty_like(target) if want_masked else None
        gm = torch.empty_like(target) if want_masked else None
        _call("fcd_masked_recon_fwd", target.data_ptr(), generate.data_ptr(), cmap.data_ptr(), B, C, H, W, kind,
              sums.data_ptr(), out2.data_ptr(), _lib.ptr(tm), _lib.ptr(gm))
        ctx.state = (target, generate, cmap, sums, kind)
        if not want_masked:
            tm = gm = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(tm, gm)
        return out2[0], out2[1], tm, gm

    @staticmethod
    def backward(ctx, g_gen, g_l1, g_tm, g_gm):
        target, generate, cmap, sums, kind = ctx.state
        B, C, H, W = target.shape
        need_t, need_g, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dt = torch.empty_like(target) if need_t else None
        dg = torch.empty_like(generate) if need_g else None
        dc = torch.empty_like(cmap) if need_c else None

        def scalar(g):
            return None if g is None else g.contiguous().to(torch.float32)

        g_gen, g_l1 = scalar(g_gen), scalar(g_l1)
        if (g_tm is None) != (g_gm is None):
            z = torch.zeros_like(target)
            g_tm = z if g_tm is None else g_tm
            g_gm = z if g_gm is None else g_gm
        if g_tm is not None:
            g_tm, g_gm = g_tm.contiguous(), g_gm.contiguous()
        _call("fcd_masked_recon_bwd", target.data_ptr(), generate.data_ptr(), cmap.data_ptr(), B, C, H, W, kind,
              sums.data_ptr(), _lib.ptr(g_gen), _lib.ptr(g_l1), _lib.ptr(g_tm), _lib.ptr(g_gm), _lib.ptr(dt), _lib.ptr(dg),
              _lib.ptr(dc))
        return dt, dg, dc, None, None


class _RegionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cmap, region, kind: int):
        _check(cmap, "region_loss")
        _check(region, "region_loss")
        if cmap.shape != region.shape:
            raise ValueError(f"region_loss: cmap {tuple(cmap.shape)} and region {tuple(region.shape)} differ")
        cmap, region = cmap.contiguous(), region.contiguous()
        B = cmap.shape[0]
        n = cmap[0].numel()
        HW = cmap.shape[2] * cmap.shape[3]
        sums = torch.empty(2 * B, dtype=torch.float64, device=cmap.device)
        out = torch.empty((), dtype=torch.float32, device=cmap.device)
        _call("fcd_region_loss_fwd", cmap.data_ptr(), region.data_ptr(), B, n, HW, kind, sums.data_ptr(), out.data_ptr())
        ctx.state = (cmap, region, sums, kind, B, n, HW)
        return out

    @staticmethod
    def backward(ctx, gout):
        cmap, region, sums, kind, B, n, HW = ctx.state
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("region_loss: gradient w.r.t. the region mask is not implemented (it is data)")
        dc = torch.empty_like(cmap)
        gout = gout.contiguous().to(torch.float32)
        _call("fcd_region_loss_bwd", cmap.data_ptr(), region.data_ptr(), B, n, HW, kind, sums.data_ptr(), gout.data_ptr(),
              dc.data_ptr())
        return dc, None, None


class _Mean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode: int):
        _check(x, "mean")
        x = x.contiguous()
        acc = torch.empty(1, dtype=torch.float64, device=x.device)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        _call("fcd_mean_fwd", x.data_ptr(), x.numel(), mode, acc.data_ptr(), out.data_ptr())
        ctx.state = (x, mode)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, mode = ctx.state
        dx = torch.empty_like(x)
        gout = gout.contiguous().to(torch.float32)
        _call("fcd_mean_bwd", x.data_ptr(), x.numel(), mode, gout.data_ptr(), dx.data_ptr())
        return dx, None


class _SoftMask(torch.autograd.Function):
    """out = (a*(1-region) + b*region) * (1 - mask); gradient flows to `mask` only (a, b, region are data)."""

    @staticmethod
    def forward(ctx, mask, a, b, region):
        for t in (mask, a) + ((b, region) if b is not None else ()):
            _check(t, "soft_mask")
        N, C, H, W = a.shape
        if tuple(mask.shape) != (N, 1, H, W):
            raise ValueError(f"soft_mask: mask must be {(N, 1, H, W)}, got {tuple(mask.shape)}")
        mask, a = mask.contiguous(), a.contiguous()
        if b is not None:
            b, region = b.contiguous(), region.contiguous()
        out = torch.empty_like(a)
        _call("fcd_mask_fwd", a.data_ptr(), _lib.ptr(b), _lib.ptr(region), mask.data_ptr(), N, C, H, W, out.data_ptr())
        ctx.state = (a, b, region)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b, region = ctx.state
        if any(ctx.needs_input_grad[1:]):
            raise NotImplementedError("soft_mask: only the mask (change-density map) receives a gradient")
        N, C, H, W = a.shape
        gout = gout.contiguous()
        dm = torch.empty((N, 1, H, W), dtype=torch.float32, device=a.device)
        _call("fcd_mask_bwd", gout.data_ptr(), a.data_ptr(), _lib.ptr(b), _lib.ptr(region), N, C, H, W, dm.data_ptr(), 0)
        return dm, None, None, None


# ---- fused forms of the inline terms of the training loops --------------------------------------------
def mean(x):
    """x.mean()  — WGAN terms, Demo_RSSS.py:304,324."""
    return _Mean.apply(x, 0)


def mean_abs(x):
    """torch.mean(abs(x))  — Demo_WSSS.py:315, Loss.py:87."""
    return _Mean.apply(x, 1)


def mean_sq(x):
    """torch.mean(x ** 2)  — nc_loss, Demo_WSSS.py:299."""
    return _Mean.apply(x, 2)


def soft_mask(image, cmap, other=None, region=None):
    """image * (1 - cmap.repeat(1, C, 1, 1))  (Demo_RSSS.py:290-291); with `other` and `region`:
    (image*(1-region) + other*region) * (1 - cmap), the masked fake-unchanged pair of Demo_RSSS.py:296-300."""
    if (other is None) != (region is None):
        raise ValueError("soft_mask: give both `other` and `region` or neither")
    return _SoftMask.apply(cmap, image, other, region)


def region_loss(cmap, region, criterion):
    """Loss.py:127-141.  `criterion` must be nn.L1Loss() or nn.MSELoss() with the default 'mean' reduction — the two
    the reference uses (Demo_RSSS.py:322-327)."""
    if isinstance(criterion, nn.L1Loss):
        kind = LOSS_L1
    elif isinstance(criterion, nn.MSELoss):
        kind = LOSS_MSE
    else:
        raise NotImplementedError(f"region_loss: criterion {type(criterion).__name__} has no fused kernel (L1Loss / MSELoss)")
    if getattr(criterion, "reduction", "mean") != "mean":
        raise NotImplementedError("region_loss: only reduction='mean'")
    return _RegionLoss.apply(cmap, region, kind)


# ---- perception loss: torch passthrough (out of scope for the CUDA path) ------------------------------
class PerceptionLoss(nn.Module):
    """Loss.py:17-61 restated over torchvision's VGG16 in plain PyTorch.  `vgg_features` lets the caller supply the
    `vgg16().features` module (e.g. loaded from a local vgg16-397923af.pth); without it the pretrained weights are
    requested like the reference does, and if that fails (no network) the term is disabled."""

    FEATURE_LAYERS = [29, 22, 15, 8, 3]

    def __init__(self, feature_layer=1, perception_perBand=False, vgg_features: Optional[nn.Module] = None):
        super().__init__()
        self.enabled = True
        if vgg_features is None:
            try:
                from torchvision.models.vgg import vgg16
                vgg_features = vgg16(pretrained=True).features
            except Exception as e:  # no network / no cached weights
                warnings.warn(f"PerceptionLoss disabled (VGG16 weights unavailable: {type(e).__name__}); the perception "
                              "term is 0.  Pass vgg_features= to enable it.")
                self.enabled = False
        if self.enabled:
            self.net = vgg_features.eval()
            for p in self.net.parameters():
                p.requires_grad = False
        feature_layer = feature_layer if feature_layer > 0 else 1
        feature_layer = feature_layer if feature_layer < 6 else 5
        self.feature_layer_list = self.FEATURE_LAYERS[:feature_layer]
        self.perception_perBand = perception_perBand
        self.loss = nn.MSELoss()

    def _features_loss(self, x, y, scale):
        total = 0
        for i, layer in enumerate(self.net):
            x, y = layer(x), layer(y)
            if i in self.feature_layer_list:
                total = total + self.loss(x, y) / scale
        return total

    def forward(self, target_image, generate_image, cmask):
        if not self.enabled:
            return torch.zeros((), dtype=torch.float32, device=target_image.device)
        layer_num = len(self.feature_layer_list)
        if not self.perception_perBand:
            assert target_image.shape[1] >= 3
            m = 1 - cmask
            return self._features_loss(target_image[:, 0:3] * m, generate_image[:, 0:3] * m, layer_num)
        n_channels = target_image.shape[1]
        total = 0
        for b in range(n_channels):
            x = (target_image[:, b:b + 1] * (1 - cmask)).repeat((1, 3, 1, 1))
            y = (generate_image[:, b:b + 1] * (1 - cmask)).repeat((1, 3, 1, 1))
            total = total + self._features_loss(x, y, layer_num * n_channels)
        return total


class CNetLoss(nn.Module):
    """Loss.py:64-95 (unsupervised USSS loss).  Returns (generator_loss, l1_loss, perception_loss, ssim_loss)."""

    def __init__(self, channel=4, perception_layer=1, perception_perBand=True, vgg_features=None):
        super().__init__()
        self.loss_perception = PerceptionLoss(feature_layer=perception_layer, perception_perBand=perception_perBand,
                                              vgg_features=vgg_features)
        self.ssim = MS_SSIM(data_range=1.0, channel=channel)

    def forward(self, target_image, generate_image, cmap, generator_mask_switch=False):
        generator_loss, l1_loss, tm, gm = _MaskedRecon.apply(target_image, generate_image, cmap, LOSS_L1, True)
        if self.loss_perception.enabled:
            cmask = (torch.sign(cmap - 0.5) + 1) / 2 if generator_mask_switch else cmap
            perception_loss = self.loss_perception(target_image, generate_image, cmask)
        else:
            perception_loss = self.loss_perception(target_image, generate_image, cmap)
        ssim_loss = 1 - self.ssim(tm, gm)
        return generator_loss, l1_loss, perception_loss, ssim_loss


class CGeneratorLoss(nn.Module):
    """Loss.py:100-124 (RSSS / WSSS generator loss; all-changed samples skipped).  Returns
    (generator_loss, ssim_loss, perception_loss)."""

    def __init__(self, channel=3, perception_layer=1, perception_perBand=False, vgg_features=None):
        super().__init__()
        self.ssim = MS_SSIM(data_range=1.0, channel=channel)
        self.loss_perception = PerceptionLoss(feature_layer=perception_layer, perception_perBand=perception_perBand,
                                              vgg_features=vgg_features)

    def forward(self, target_image, generate_image, cmap):
        generator_loss, _, tm, gm = _MaskedRecon.apply(target_image, generate_image, cmap, LOSS_MSE, True)
        ssim_loss = 1 - self.ssim(tm, gm)
        perception_loss = self.loss_perception(target_image, generate_image, cmap)
        return generator_loss, ssim_loss, perception_loss
