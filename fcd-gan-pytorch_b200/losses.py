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

PerceptionLoss (VGG16, Loss.py:17-61) runs its frozen feature stack on the same conv engine as the
networks when the caller supplies the VGG16 weights (`vgg_features=`); it never downloads anything.
"""
from __future__ import annotations

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
    """out = (a*(1-region) + b*region) * (1 - mask); gradient flows to `mask` and, in the plain form (no b / region), to
    the image `a` (the perception loss masks the GENERATED image, Loss.py:43,54); b and region are data."""

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
        ctx.state = (a, b, region, mask)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b, region, mask = ctx.state
        if any(ctx.needs_input_grad[2:]) or (ctx.needs_input_grad[1] and b is not None):
            raise NotImplementedError("soft_mask: gradients flow to the mask and (plain form only) to the image")
        N, C, H, W = a.shape
        gout = gout.contiguous()
        dm = da = None
        if ctx.needs_input_grad[0]:
            dm = torch.empty((N, 1, H, W), dtype=torch.float32, device=a.device)
            _call("fcd_mask_bwd", gout.data_ptr(), a.data_ptr(), _lib.ptr(b), _lib.ptr(region), N, C, H, W, dm.data_ptr(), 0)
        if ctx.needs_input_grad[1]:          # d(a * (1 - mask)) / da = (1 - mask): the same kernel applied to the gradient
            da = torch.empty_like(a)
            _call("fcd_mask_fwd", gout.data_ptr(), None, None, mask.data_ptr(), N, C, H, W, da.data_ptr())
        return dm, da, None, None


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


# ---- perception loss: frozen VGG16 feature stack on the conv engine ------------------------------------
class PerceptionLoss(nn.Module):
    """Loss.py:17-61 on the tcgen05 conv engine.

    The reference builds `vgg16(pretrained=True).features.eval()` itself (Loss.py:25), i.e. it DOWNLOADS the ImageNet weights;
    a library on a production box must not, so here the caller supplies them: `vgg_features` is torchvision's
    `vgg16().features` module (any weights), or the path of a torchvision VGG16 state_dict (`vgg16-397923af.pth`).  Without it
    the term is DISABLED: `forward` returns 0 and the step bodies refuse a non-zero perception weight (steps._check_perception)
    — nothing is downloaded, nothing falls back to torch.

    What runs: features[0..29] = thirteen 3x3 convolutions (+bias) + ReLU and four 2x2 max-pools, with the weights frozen
    (Loss.py:26-27) — `engine.conv(frozen=True)` (forward + data gradient only), `engine.bn_act` without BatchNorm,
    `engine.maxpool2` — on ONE batch that holds the masked target images in its first half and the masked generated images
    in its second; `engine.mse_halves` is the nn.MSELoss at the selected layers (Loss.py:31-36).  `perception_perBand=True`
    (Loss.py:50-60) feeds every band as a 3-channel grey image: all bands of all samples are batched into one pass (B*C
    images; the per-band means divided by n_channels are the mean over that batch), and the three identical input channels
    are folded into the first layer's weights (sum over its input-channel axis).  Gradients reach `generate_image` and, when
    it requires one, the soft mask `cmask` — through both halves, like the reference's autograd."""

    FEATURE_LAYERS = [29, 22, 15, 8, 3]

    def __init__(self, feature_layer=1, perception_perBand=False, vgg_features=None):
        super().__init__()
        if isinstance(vgg_features, (str, bytes)) or hasattr(vgg_features, "__fspath__"):
            from torchvision.models.vgg import vgg16

            full = vgg16(weights=None)
            full.load_state_dict(torch.load(vgg_features, map_location="cpu"))
            vgg_features = full.features
        self.enabled = vgg_features is not None
        if self.enabled:
            self.net = vgg_features.eval()
            for p in self.net.parameters():
                p.requires_grad = False
            for i in range(30):
                m = self.net[i]
                ok = ((isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3) and m.padding == (1, 1) and m.stride == (1, 1))
                      or isinstance(m, nn.ReLU) or (isinstance(m, nn.MaxPool2d) and m.kernel_size in (2, (2, 2))))
                if not ok:
                    raise ValueError(f"PerceptionLoss: vgg_features[{i}] = {m} is not a VGG16 feature layer")
        feature_layer = feature_layer if feature_layer > 0 else 1
        feature_layer = feature_layer if feature_layer < 6 else 5
        self.feature_layer_list = self.FEATURE_LAYERS[:feature_layer]
        self.perception_perBand = perception_perBand

    def train(self, mode: bool = True):
        super().train(mode)
        if self.enabled:
            self.net.eval()          # Loss.py:25: the VGG stays in eval mode whatever the criterion's mode
        return self

    def _run(self, x: torch.Tensor, y: torch.Tensor, fold_input: bool) -> torch.Tensor:
        from . import engine as E

        layers = self.feature_layer_list
        last = max(layers)

        def fn(tape, inputs, need):
            a = E.stage_two_inputs(tape, inputs[0], inputs[1])
            slot = {}
            acc = torch.zeros(len(layers), dtype=torch.float64, device=tape.device)
            coefs = []
            h = a
            for i in range(last + 1):
                m = self.net[i]
                if isinstance(m, nn.Conv2d):
                    w = m.weight
                    if i == 0 and fold_input:      # three identical input channels == one channel with the summed filter
                        w = E._derived(w, "fold_in", lambda t: t.sum(1, keepdim=True).contiguous())
                    z = E.conv(tape, h, w, m.bias, 1, 1, stats=False, x_needs_grad=(i > 0 or need[0] or need[1]), frozen=True,
                               wtag="vgg")
                elif isinstance(m, nn.ReLU):
                    h = E.bn_act(tape, z, None, False, E.ACT_RELU)
                    if i in layers:
                        coefs.append(E.mse_halves(tape, h, 1.0 / len(layers), acc[len(coefs):], slot))
                else:
                    h = E.maxpool2(tape, h)
            out = acc[0] * coefs[0]          # scalar glue: weighted means of the per-layer sums
            for k in range(1, len(coefs)):
                out = out + acc[k] * coefs[k]
            out = out.to(torch.float32)
            n = inputs[0].shape[0]
            return out, slot, [E.BatchHalf(a, 0, n), E.BatchHalf(a, n, n)]

        return E.run_net(self, fn, x, y)

    def forward(self, target_image, generate_image, cmask):
        if not self.enabled:
            return torch.zeros((), dtype=torch.float32, device=target_image.device)
        for t in (target_image, generate_image, cmask):
            _check(t, "PerceptionLoss")
        B, C, H, W = target_image.shape
        if min(H, W) < 16:
            raise ValueError("PerceptionLoss: images must be at least 16x16 (four 2x2 poolings)")
        if not self.perception_perBand:
            assert C >= 3
            x = soft_mask(target_image[:, 0:3].contiguous(), cmask)
            y = soft_mask(generate_image[:, 0:3].contiguous(), cmask)
            return self._run(x, y, False)
        x = soft_mask(target_image, cmask).reshape(B * C, 1, H, W)
        y = soft_mask(generate_image, cmask).reshape(B * C, 1, H, W)
        return self._run(x, y, True)


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
            cmask = ((torch.sign(cmap.detach() - 0.5) + 1) / 2) if generator_mask_switch else cmap     # Loss.py:75,89-92
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
