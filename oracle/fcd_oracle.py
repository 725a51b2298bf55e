"""CPU oracle for the FCD-GAN hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional torch-CPU (fp32 or fp64) restatement of the reference's networks and loss stack.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it,
and only as the checker / CPU baseline; the product path (fcdgan_b200) never does.

Pinning: the reference has no tests and no golden vectors; this port is pinned against
outputs of the UNMODIFIED reference modules imported from /root/reference in the build container
(`oracle/make_golden.py` -> `tests/golden/*.pt`, checked by `tests/test_oracle_golden.py`), and, where the
reference is mounted, live against it on fresh seeds (`tests/test_oracle_vs_reference.py`).

The arithmetic itself lives in PyTorch (un-pinned by the reference, README.md:9); here torch 2.11 CPU.
Every function cites the reference lines it restates.  Parameters are plain dicts keyed by the reference's
state_dict names so the same weights drive reference, oracle and CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5  # nn.BatchNorm2d default, relied on at Module.py:27,30,156,178,181,200,204,208
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------------
# deterministic parameter sets (shared by golden generation, tests, smoke and bench)
# --------------------------------------------------------------------------------------------------
def _conv_keys(prefix: str, cout: int, cin: int, k: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [(f"{prefix}.weight", (cout, cin, k, k), "conv_w"), (f"{prefix}.bias", (cout,), "conv_b")]


def _bn_keys(prefix: str, c: int):
    return [(f"{prefix}.weight", (c,), "bn_w"), (f"{prefix}.bias", (c,), "bn_b"),
            (f"{prefix}.running_mean", (c,), "bn_rm"), (f"{prefix}.running_var", (c,), "bn_rv"),
            (f"{prefix}.num_batches_tracked", (), "bn_n")]


def _double_conv_keys(prefix: str, cin: int, cout: int, mid: int | None = None):
    mid = mid or cout
    p = f"{prefix}.double_conv"
    return (_conv_keys(f"{p}.0", mid, cin, 3) + _bn_keys(f"{p}.1", mid) + _conv_keys(f"{p}.3", cout, mid, 3) +
            _bn_keys(f"{p}.4", cout))


def generator_spec(c: int):
    """state_dict layout of Module.py:142-158 (Generator) + 174-181 (ResidualBlock)."""
    keys = _conv_keys("block1.0", 64, c, 9) + [("block1.1.weight", (1,), "prelu")]
    for b in range(2, 7):
        keys += _conv_keys(f"block{b}.conv1", 64, 64, 3) + _bn_keys(f"block{b}.bn1", 64)
        keys += [(f"block{b}.prelu.weight", (1,), "prelu")]
        keys += _conv_keys(f"block{b}.conv2", 64, 64, 3) + _bn_keys(f"block{b}.bn2", 64)
    keys += _conv_keys("block7.0", 64, 64, 3) + _bn_keys("block7.1", 64)
    keys += _conv_keys("block8", c, 64, 9)
    return keys


def segmentor_spec(c: int, n_out: int = 1, bilinear: bool = True):
    """state_dict layout of Module.py:93-111 (Segmentor) built from DoubleConv/Down/Up/OutConv (18-90)."""
    f = 2 if bilinear else 1
    keys = _double_conv_keys("inc", c, 64)
    for i, (ci, co) in enumerate([(64, 128), (128, 256), (256, 512), (512, 1024 // f)], start=1):
        keys += _double_conv_keys(f"down{i}.maxpool_conv.1", ci, co)
    for i, (ci, co) in enumerate([(2048, 1024 // f), (1024, 512 // f), (512, 256 // f), (256, 128)], start=1):
        if bilinear:
            keys += _double_conv_keys(f"up{i}.conv", ci, co, ci // 2)
        else:
            keys += [(f"up{i}.up.weight", (ci, ci // 2, 2, 2), "convT_w"), (f"up{i}.up.bias", (ci // 2,), "conv_b")]
            keys += _double_conv_keys(f"up{i}.conv", ci, co)
    keys += _conv_keys("outc.conv", n_out, 128, 1)
    return keys


def discriminator_spec(c: int):
    """state_dict layout of Module.py:192-217 (Discriminator_SRGAN_simple)."""
    keys = _conv_keys("net.0", 64, c, 3)
    for idx, (ci, co) in zip((2, 5, 8), ((64, 128), (128, 256), (256, 512))):
        keys += _conv_keys(f"net.{idx}", co, ci, 3) + _bn_keys(f"net.{idx + 1}", co)
    keys += _conv_keys("classifier.1", 1024, 512, 1) + _conv_keys("classifier.3", 1, 1024, 1)
    return keys


def make_state_dict(spec, seed: int, dtype=torch.float32) -> SD:
    """Deterministic, non-trivial parameters (BN affine/running stats are perturbed so that eval-mode and
    bias handling are exercised).  Scale follows torch's default conv init (U(+-1/sqrt(fan_in)))."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape, kind in spec:
        if kind in ("conv_w", "convT_w"):
            fan_in = shape[1] * shape[2] * shape[3] if kind == "conv_w" else shape[0] * shape[2] * shape[3] // 4
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b * math.sqrt(3.0)
        elif kind == "conv_b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "bn_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind in ("bn_b", "bn_rm"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_rv":
            t = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif kind == "bn_n":
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        elif kind == "prelu":
            t = torch.full(shape, 0.25)
        else:
            raise KeyError(kind)
        sd[name] = t.to(dtype)
    return sd


def clone_sd(sd: SD, dtype=None, requires_grad: bool = False) -> SD:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            if requires_grad and not (k.endswith("running_mean") or k.endswith("running_var")):
                t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------
def _bn(sd: SD, p: str, x: Tensor, train: bool) -> Tensor:
    """nn.BatchNorm2d (defaults): batch mean / biased var in train mode, running stats updated with the
    unbiased var and momentum 0.1; running stats in eval mode."""
    y = F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"],
                     train, BN_MOMENTUM, BN_EPS)
    if train and f"{p}.num_batches_tracked" in sd:
        sd[f"{p}.num_batches_tracked"] += 1
    return y


def _conv(sd: SD, p: str, x: Tensor, stride: int = 1, padding: int = 0) -> Tensor:
    return F.conv2d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], stride=stride, padding=padding)


def double_conv(sd: SD, p: str, x: Tensor, train: bool) -> Tensor:
    """Module.py:18-35: (conv3x3 p1 -> BN -> ReLU) x 2."""
    q = f"{p}.double_conv"
    x = F.relu(_bn(sd, f"{q}.1", _conv(sd, f"{q}.0", x, padding=1), train))
    return F.relu(_bn(sd, f"{q}.4", _conv(sd, f"{q}.3", x, padding=1), train))


def down(sd: SD, p: str, x: Tensor, train: bool) -> Tensor:
    """Module.py:38-49: MaxPool2d(2) -> DoubleConv."""
    return double_conv(sd, f"{p}.maxpool_conv.1", F.max_pool2d(x, 2), train)


def up(sd: SD, p: str, x1: Tensor, x2: Tensor, bilinear: bool, train: bool) -> Tensor:
    """Module.py:52-79: upsample x1 (bilinear align_corners=True, or ConvTranspose2d k2 s2), zero-pad it to
    x2's size with the extra row/column on the right/bottom, cat([x2, x1]), DoubleConv."""
    if bilinear:
        x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        x1 = F.conv_transpose2d(x1, sd[f"{p}.up.weight"], sd[f"{p}.up.bias"], stride=2)
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return double_conv(sd, f"{p}.conv", torch.cat([x2, x1], dim=1), train)


def segmentor(sd: SD, x1: Tensor, x2: Tensor, bilinear: bool = True, train: bool = True) -> Tensor:
    """Module.py:113-140: siamese shared-weight encoder called once per temporal image (BN statistics are
    per call), per-level concat of the two branches, U-Net decoder, 1x1 conv + sigmoid (82-90)."""
    a = [double_conv(sd, "inc", x1, train)]
    b = [double_conv(sd, "inc", x2, train)]
    for i in range(1, 5):
        a.append(down(sd, f"down{i}", a[-1], train))
        b.append(down(sd, f"down{i}", b[-1], train))
    cat = [torch.cat([u, v], dim=1) for u, v in zip(a, b)]
    x = up(sd, "up1", cat[4], cat[3], bilinear, train)
    x = up(sd, "up2", x, cat[2], bilinear, train)
    x = up(sd, "up3", x, cat[1], bilinear, train)
    x = up(sd, "up4", x, cat[0], bilinear, train)
    return torch.sigmoid(_conv(sd, "outc.conv", x))


def residual_block(sd: SD, p: str, x: Tensor, train: bool) -> Tensor:
    """Module.py:174-190: conv-BN-PReLU-conv-BN + identity."""
    r = _bn(sd, f"{p}.bn1", _conv(sd, f"{p}.conv1", x, padding=1), train)
    r = F.prelu(r, sd[f"{p}.prelu.weight"])
    r = _bn(sd, f"{p}.bn2", _conv(sd, f"{p}.conv2", r, padding=1), train)
    return x + r


def generator(sd: SD, x: Tensor, train: bool = True) -> Tensor:
    """Module.py:160-172: conv9x9+PReLU, 5 residual blocks, conv3x3+BN, conv9x9 on (block1 + x); linear out."""
    b1 = F.prelu(_conv(sd, "block1.0", x, padding=4), sd["block1.1.weight"])
    h = b1
    for i in range(2, 7):
        h = residual_block(sd, f"block{i}", h, train)
    h = _bn(sd, "block7.1", _conv(sd, "block7.0", h, padding=1), train)
    return _conv(sd, "block8", b1 + h, padding=4)


def _disc_net(sd: SD, x: Tensor, train: bool) -> Tensor:
    """Module.py:195-210: 4 stride-2 3x3 convs; LeakyReLU(0.2); BN on layers 2-4."""
    x = F.leaky_relu(_conv(sd, "net.0", x, stride=2, padding=1), 0.2)
    for idx in (2, 5, 8):
        x = F.leaky_relu(_bn(sd, f"net.{idx + 1}", _conv(sd, f"net.{idx}", x, stride=2, padding=1), train), 0.2)
    return x


def discriminator(sd: SD, x: Tensor, y: Tensor, train: bool = True) -> Tensor:
    """Module.py:219-223: shared net on x and y (two calls), classifier(GAP(fx - fy)), sigmoid -> (B,)."""
    fx = _disc_net(sd, x, train)
    fy = _disc_net(sd, y, train)
    h = F.adaptive_avg_pool2d(fx - fy, 1)
    h = F.leaky_relu(_conv(sd, "classifier.1", h), 0.2)
    h = _conv(sd, "classifier.3", h)
    return torch.sigmoid(h.view(x.shape[0]))


# --------------------------------------------------------------------------------------------------
# SSIM / MS-SSIM  (ssim.py, vendored pytorch-msssim)
# --------------------------------------------------------------------------------------------------
MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)  # ssim.py:199


def gauss_window(size: int = 11, sigma: float = 1.5) -> Tensor:
    """ssim.py:9-23: normalised 1-D Gaussian, built in fp32."""
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _blur(x: Tensor, win: Tensor) -> Tensor:
    """ssim.py:26-52: separable depthwise 'valid' blur; a dimension smaller than the window is skipped."""
    C = x.shape[1]
    k = win.numel()
    w = win.to(x.dtype)
    if x.shape[2] >= k:
        x = F.conv2d(x, w.view(1, 1, k, 1).expand(C, 1, k, 1), groups=C)
    if x.shape[3] >= k:
        x = F.conv2d(x, w.view(1, 1, 1, k).expand(C, 1, 1, k), groups=C)
    return x


def ssim_core(X: Tensor, Y: Tensor, data_range: float, win: Tensor, K=(0.01, 0.03)) -> Tuple[Tensor, Tensor]:
    """ssim.py:55-92: returns (ssim_per_channel, cs) each (B, C)."""
    C1 = (K[0] * data_range) ** 2
    C2 = (K[1] * data_range) ** 2
    mu1, mu2 = _blur(X, win), _blur(Y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur(X * X, win) - mu1_sq
    s2 = _blur(Y * Y, win) - mu2_sq
    s12 = _blur(X * Y, win) - mu12
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ssim(X: Tensor, Y: Tensor, data_range: float = 255, size_average: bool = True, win_size: int = 11,
         win_sigma: float = 1.5, K=(0.01, 0.03), nonnegative_ssim: bool = False) -> Tensor:
    """ssim.py:95-150 (4-d inputs)."""
    if X.shape != Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    s, _ = ssim_core(X, Y, data_range, gauss_window(win_size, win_sigma), K)
    if nonnegative_ssim:
        s = torch.relu(s)
    return s.mean() if size_average else s.mean(1)


def ms_ssim(X: Tensor, Y: Tensor, data_range: float = 255, size_average: bool = True, win_size: int = 11,
            win_sigma: float = 1.5, weights=None, K=(0.01, 0.03)) -> Tensor:
    """ssim.py:153-225: 5 levels; relu(cs)^w on levels 0-3, relu(ssim)^w on level 4; avg_pool2d(2, padding=s%2)
    between levels; product over levels; mean over (B, C)."""
    if X.shape != Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    assert min(X.shape[-2:]) > (win_size - 1) * 16, \
        "Image size should be larger than %d due to the 4 downsamplings in ms-ssim" % ((win_size - 1) * 16)
    w = torch.tensor(list(weights) if weights is not None else MS_WEIGHTS, dtype=torch.float32).to(X.dtype)
    win = gauss_window(win_size, win_sigma)
    vals = []
    L = w.numel()
    for i in range(L):
        s, cs = ssim_core(X, Y, data_range, win, K)
        if i < L - 1:
            vals.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X = F.avg_pool2d(X, kernel_size=2, padding=pad)
            Y = F.avg_pool2d(Y, kernel_size=2, padding=pad)
    vals.append(torch.relu(s))
    stack = torch.stack(vals, dim=0)
    out = torch.prod(stack ** w.view(-1, 1, 1), dim=0)
    return out.mean() if size_average else out.mean(1)


# --------------------------------------------------------------------------------------------------
# losses (Loss.py) — the VGG perception term (Loss.py:17-61) is out of scope
# --------------------------------------------------------------------------------------------------
def masked_recon_loss(t: Tensor, g: Tensor, cmap: Tensor, kind: str, skip_empty: bool) -> Tensor:
    """Loss.py:76-84 (L1, CNetLoss) / Loss.py:109-119 (MSE, CGeneratorLoss; samples with sum(1-cmap)==0 skipped,
    divisor stays B):  (1/B) sum_i mean_{c,p}(crit(t*m, g*m)) * P / sum_p m_i,  m = 1 - cmap."""
    B, C, H, W = t.shape
    m = 1 - cmap
    num_wnc = m.sum(dim=(1, 2, 3))
    tm, gm = t * m, g * m
    total = 0
    for i in range(B):
        if skip_empty and num_wnc[i] == 0:
            continue
        d = tm[i] - gm[i]
        crit = d.abs().mean() if kind == "l1" else (d * d).mean()
        total = total + crit * (H * W) / num_wnc[i]
    return total / B


def cnet_loss(t: Tensor, g: Tensor, cmap: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Loss.py:73-95 without the perception term: (generator_loss[L1], l1_loss = mean|cmap|, ssim_loss)."""
    gen = masked_recon_loss(t, g, cmap, "l1", skip_empty=False)
    l1 = cmap.abs().mean()
    m = 1 - cmap
    ssim_loss = 1 - ms_ssim(t * m, g * m, data_range=1.0)
    return gen, l1, ssim_loss


def cgenerator_loss(t: Tensor, g: Tensor, cmap: Tensor) -> Tuple[Tensor, Tensor]:
    """Loss.py:108-124 without the perception term: (generator_loss[MSE, empty samples skipped], ssim_loss)."""
    gen = masked_recon_loss(t, g, cmap, "mse", skip_empty=True)
    m = 1 - cmap
    ssim_loss = 1 - ms_ssim(t * m, g * m, data_range=1.0)
    return gen, ssim_loss


# --------------------------------------------------------------------------------------------------
# perception loss (Loss.py:17-61) — SURVEY.md §8(f) N1
# --------------------------------------------------------------------------------------------------
VGG16_FEATURES_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]   # torchvision cfg "D"
VGG_FEATURE_LAYERS = [29, 22, 15, 8, 3]      # Loss.py:30


def vgg16_features(seed: int = 1234):
    """torchvision's `vgg16(weights=None).features` under a fixed seed, frozen, eval — the stand-in for the ImageNet weights
    the reference downloads at Loss.py:25 (no network here).  oracle/ref_import.py rebinds `Loss.vgg16` to the same
    construction, so reference, oracle and CUDA path see identical weights."""
    import torchvision

    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = torchvision.models.vgg16(weights=None).features.eval()
    torch.random.set_rng_state(g)
    for p in net.parameters():
        p.requires_grad = False
    return net


def vgg_feature_mse(vgg_sd: SD, x: Tensor, y: Tensor, layers: List[int], scale: float) -> Tensor:
    """Loss.py:44-48 / 56-59: run x and y through features[0..max(layers)], summing MSE(x, y) / scale at `layers`.
    `vgg_sd`: state_dict of vgg16.features (keys "0.weight", "0.bias", "2.weight", ...)."""
    total = 0
    i = 0
    for v in VGG16_FEATURES_CFG:
        if v == "M":
            x, y = F.max_pool2d(x, 2), F.max_pool2d(y, 2)
            i += 1
        else:
            w, b = vgg_sd[f"{i}.weight"].to(x.dtype), vgg_sd[f"{i}.bias"].to(x.dtype)
            x, y = F.conv2d(x, w, b, padding=1), F.conv2d(y, w, b, padding=1)
            i += 1                                   # the ReLU layer
            x, y = F.relu(x), F.relu(y)
            if i in layers:
                total = total + F.mse_loss(x, y) / scale
            i += 1
        if i > max(layers):
            break
    return total


def perception_loss(vgg_sd: SD, target: Tensor, generate: Tensor, cmask: Tensor, feature_layer: int = 1,
                    per_band: bool = False) -> Tensor:
    """PerceptionLoss.forward, Loss.py:38-61."""
    feature_layer = min(max(feature_layer, 1), 5)                      # Loss.py:32-33
    layers = VGG_FEATURE_LAYERS[:feature_layer]
    n = len(layers)
    if not per_band:                                                   # Loss.py:40-48
        m = 1 - cmask.repeat((1, 3, 1, 1))
        return vgg_feature_mse(vgg_sd, target[:, 0:3] * m, generate[:, 0:3] * m, layers, n)
    C = target.shape[1]                                                # Loss.py:50-60
    total = 0
    for b in range(C):
        x = (target[:, b].unsqueeze(1) * (1 - cmask)).repeat((1, 3, 1, 1))
        y = (generate[:, b].unsqueeze(1) * (1 - cmask)).repeat((1, 3, 1, 1))
        total = total + vgg_feature_mse(vgg_sd, x, y, layers, n * C)
    return total


def region_loss(cmap: Tensor, region: Tensor, kind: str) -> Tensor:
    """Loss.py:127-141: (1/B) sum_{i: sum(region_i) > 0} mean_p(crit(cmap_i*region_i, 0)) * P / sum(region_i)."""
    B, _, H, W = cmap.shape
    num = region.sum(dim=(1, 2, 3))
    v = cmap * region
    total = 0
    for i in range(B):
        if num[i] == 0:
            continue
        crit = v[i].abs().mean() if kind == "l1" else (v[i] * v[i]).mean()
        total = total + crit * (H * W) / num[i]
    return total / B


# --------------------------------------------------------------------------------------------------
# step bodies (the callers of the path; Demo_USSS.py:305-341, Demo_RSSS.py:270-332) with perception weight 0
# --------------------------------------------------------------------------------------------------
def usss_joint_losses(sdG: SD, sdS: SD, x: Tensor, y: Tensor, ssim_weight: float, l1_weight: float,
                      bilinear: bool = True):
    """Forward of the USSS joint iteration (Demo_USSS.py:320-336): returns (Loss, NetLoss, cmap, parts)."""
    y_fake = generator(sdG, x, train=True)
    cmap = segmentor(sdS, x, y, bilinear=bilinear, train=True)
    gen, l1, ss = cnet_loss(y, y_fake, cmap)
    loss = gen + ssim_weight * ss
    net_loss = gen + l1_weight * l1 + ssim_weight * ss
    return loss, net_loss, cmap, (gen, l1, ss)


def rsss_d_loss(sdS: SD, sdD: SD, x: Tensor, y: Tensor, region: Tensor, bilinear: bool = True):
    """Demo_RSSS.py:285-304 (discriminator_continuous=True): returns (d_loss, cmap, x_mask, y_mask)."""
    cmap = segmentor(sdS, x, y, bilinear=bilinear, train=True)
    C = x.shape[1]
    m = 1 - cmap.repeat(1, C, 1, 1)
    x_mask, y_mask = x * m, y * m
    c_out = discriminator(sdD, x_mask, y_mask, train=True)
    y_unc = y * (1 - region) + x * region
    nc_out = discriminator(sdD, x * m, y_unc * m, train=True)
    return 1 + nc_out.mean() - c_out.mean(), cmap, x_mask, y_mask


def rsss_s_loss(sdG: SD, sdD: SD, x: Tensor, y: Tensor, region: Tensor, cmap: Tensor, x_mask: Tensor, y_mask: Tensor,
                d_weight=1.0, l1_weight=0.02, g_weight=0.5, r_weight=2.0, ssim_weight=0.0):
    """Demo_RSSS.py:317-328 with perception weight 0; G runs in eval mode (Demo_RSSS.py:240)."""
    c_out = discriminator(sdD, x_mask, y_mask, train=True)
    y_fake = generator(sdG, x, train=False)
    gen, ss = cgenerator_loss(y, y_fake, cmap)
    g_loss = gen + ssim_weight * ss
    l1 = region_loss(cmap, region, "l1")
    r = region_loss(cmap, 1 - region, "mse")
    return d_weight * c_out.mean() + l1_weight * l1 + g_weight * g_loss + r_weight * r


def wsss_d_loss(sdS: SD, sdD: SD, x: Tensor, y: Tensor, x_nc: Tensor, y_nc: Tensor, bilinear: bool = True):
    """Demo_WSSS.py:247-283 (discriminator_continuous=True): the UNCHANGED pair is masked with the CHANGED pair's map.
    Returns (d_loss, cmap, ncmap, x_mask, y_mask, c_out, nc_out)."""
    cmap = segmentor(sdS, x, y, bilinear=bilinear, train=True)
    C = x.shape[1]
    m = 1 - cmap.repeat(1, C, 1, 1)
    x_mask, y_mask = x * m, y * m
    c_out = discriminator(sdD, x_mask, y_mask, train=True)
    ncmap = segmentor(sdS, x_nc, y_nc, bilinear=bilinear, train=True)
    nc_out = discriminator(sdD, x_nc * m, y_nc * m, train=True)
    return 1 + nc_out.mean() - c_out.mean(), cmap, ncmap, x_mask, y_mask, c_out, nc_out


def wsss_s_loss(sdG: SD, sdD: SD, x: Tensor, y: Tensor, cmap: Tensor, ncmap: Tensor, x_mask: Tensor, y_mask: Tensor,
                d_weight=1.0, l1_weight=1.6, g_weight=0.2, nc_weight=1.5, ssim_weight=0.0):
    """Demo_WSSS.py:295-319 with perception weight 0; G runs in eval mode (Demo_WSSS.py:207)."""
    nc_loss = (ncmap ** 2).mean()
    c_out = discriminator(sdD, x_mask, y_mask, train=True)
    y_fake = generator(sdG, x, train=False)
    gen, ss = cgenerator_loss(y, y_fake, cmap)
    g_loss = gen + ssim_weight * ss
    l1 = cmap.abs().mean()
    return d_weight * c_out.mean() + l1_weight * l1 + g_weight * g_loss + nc_weight * nc_loss
