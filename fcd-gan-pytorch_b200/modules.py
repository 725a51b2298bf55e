"""Drop-in replacements for the networks of the reference's Module.py (same class names, constructor
arguments, call signatures and state_dict keys — SURVEY.md §8(b)), executed by hand-written sm_100a kernels.

The torch.nn layers created here (nn.Conv2d, nn.BatchNorm2d, nn.PReLU ...) are PARAMETER HOLDERS ONLY: they
give the modules the reference's parameter names / shapes / default initialisation, `.to()`, `.train()`,
`.state_dict()` and optimizer plumbing.  Their forward methods are never called; all arithmetic goes through
`engine` (libfcd_b200.so).  A reference `.pkl` state_dict loads unchanged and vice versa.

Reference: Module.py:18-223.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine as E


def _bn(m: nn.BatchNorm2d) -> E.BN:
    return E.BN(m.weight, m.bias, m.running_mean, m.running_var, m.num_batches_tracked)


def _check_input(x: torch.Tensor, channels: int, what: str):
    if x.dim() != 4 or x.shape[1] != channels:
        raise ValueError(f"{what}: expected (B, {channels}, H, W), got {tuple(x.shape)}")


def _check_bn_width(c: int, what: str):
    """Widths of BatchNorm-ed tensors the reduction kernels take: the channel count, padded to a multiple of 64, must be a
    power of two <= 2048 (fcd_bn_stats / fcd_bn_act_bwd_reduce walk 256-thread blocks over Cp / 8 channel groups).  Every
    width of the reference's networks qualifies (64 ... 1024, Module.py:101-111, 195-210)."""
    cp = E.pad_ch(c)
    if c < 1 or cp > 2048 or cp & (cp - 1):
        raise ValueError(f"{what}: unsupported channel count {c} (padded to {cp}): BatchNorm-ed tensors need a padded channel "
                         "count that is a power of two between 64 and 2048")


# ------------------------------------------------------------------------------------------------
# engine-level building blocks (operate on internal NHWC activations)
# ------------------------------------------------------------------------------------------------
def _double_conv(tape, dc: "DoubleConv", x: E.Act, training: bool, out=None, x_needs_grad=True) -> E.Act:
    """(conv3x3 p1 -> BN -> ReLU) x 2, Module.py:18-35."""
    seq = dc.double_conv
    if isinstance(x, E.PackedAct):      # <= 16-band input without a gradient: 4-pixel channel-packed first conv
        z = E.conv_small_in(tape, x, seq[0].weight, seq[0].bias, 1, stats=training)
    else:
        z = E.conv(tape, x, seq[0].weight, seq[0].bias, 1, 1, stats=training, x_needs_grad=x_needs_grad)
    m = E.bn_act(tape, z, _bn(seq[1]), training, E.ACT_RELU)
    z = E.conv(tape, m, seq[3].weight, seq[3].bias, 1, 1, stats=training)
    return E.bn_act(tape, z, _bn(seq[4]), training, E.ACT_RELU, out=out)


def _residual_block(tape, rb: "ResidualBlock", x: E.Act, training: bool) -> E.Act:
    """conv-BN-PReLU-conv-BN + identity, Module.py:183-190."""
    z = E.conv(tape, x, rb.conv1.weight, rb.conv1.bias, 1, 1, stats=training)
    a = E.bn_act(tape, z, _bn(rb.bn1), training, E.ACT_PRELU, slope=rb.prelu.weight)
    z = E.conv(tape, a, rb.conv2.weight, rb.conv2.bias, 1, 1, stats=training)
    return E.bn_act(tape, z, _bn(rb.bn2), training, E.ACT_NONE, residual=x)


# ------------------------------------------------------------------------------------------------
class DoubleConv(nn.Module):
    """(convolution => [BN] => ReLU) * 2 — Module.py:18-35."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        _check_bn_width(mid_channels, "DoubleConv(mid_channels)")
        _check_bn_width(out_channels, "DoubleConv(out_channels)")
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        _check_input(x, self.double_conv[0].in_channels, "DoubleConv")

        def fn(tape, inputs, need):
            a = E.stage_input(tape, inputs[0], need[0])
            o = _double_conv(tape, self, a, self.training, x_needs_grad=need[0])
            slot = {}
            return E.act_to_nchw(tape, o, slot), slot, [a]

        return E.run_net(self, fn, x)


class Down(nn.Module):
    """Downscaling with maxpool then double conv — Module.py:38-49."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        _check_input(x, self.maxpool_conv[1].double_conv[0].in_channels, "Down")

        def fn(tape, inputs, need):
            a = E.stage_input(tape, inputs[0], need[0])
            o = _double_conv(tape, self.maxpool_conv[1], E.maxpool2(tape, a), self.training)
            slot = {}
            return E.act_to_nchw(tape, o, slot), slot, [a]

        return E.run_net(self, fn, x)


class Up(nn.Module):
    """Upscaling then double conv — Module.py:52-79.  forward(x1, x2): x1 is upsampled, zero padded to x2's
    size and concatenated AFTER x2."""

    def __init__(self, in_channels, out_channels, bilinear=False):
        super().__init__()
        self.bilinear = bool(bilinear)
        if in_channels % 128:
            raise ValueError(f"Up: unsupported channel count {in_channels}: the skip and the upsampled tensor (in_channels / 2 "
                             "each) share one NHWC concatenation buffer in 64-channel slots, so in_channels must be a multiple "
                             "of 128 (the reference uses 256 ... 2048, Module.py:107-110)")
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels)

    def _run(self, tape, x1: E.Act, cat: E.Act, up_slot: E.Act, training: bool) -> E.Act:
        if self.bilinear:
            E.upsample2x_into(tape, x1, up_slot)
        else:
            E.conv_transpose2x2_into(tape, x1, self.up.weight, self.up.bias, up_slot)
        return _double_conv(tape, self.conv, cat, training)

    def forward(self, x1, x2):
        c1 = x1.shape[1] if self.bilinear else x1.shape[1] // 2
        cin = self.conv.double_conv[0].in_channels
        if x2.shape[1] + c1 != cin:
            raise ValueError(f"Up: channels {x2.shape[1]} + {c1} do not match DoubleConv input {cin}")

        def fn(tape, inputs, need):
            a1 = E.stage_input(tape, inputs[0], need[0])
            N, C2, H, W = inputs[1].shape
            assert C2 % 8 == 0 and c1 % 8 == 0, "Up: channel counts must be multiples of 8"
            cat = tape.new_act(N, H, W, C2 + c1, C2 + c1)
            skip = tape.track(cat.slice(0, C2))
            E.stage_input_into(tape, inputs[1], skip)
            up_slot = tape.track(cat.slice(C2, c1))
            o = self._run(tape, a1, cat, up_slot, self.training)
            slot = {}
            return E.act_to_nchw(tape, o, slot), slot, [a1, skip]

        return E.run_net(self, fn, x1, x2)


class OutConv(nn.Module):
    """1x1 convolution + sigmoid — Module.py:82-90."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        if in_channels > 128 or out_channels > 4:
            raise ValueError(f"OutConv({in_channels}, {out_channels}): the fused 1x1 + sigmoid kernel takes at most 128 input and 4 "
                             "output channels (the reference uses 128 -> 1, Module.py:111)")
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        _check_input(x, self.conv.in_channels, "OutConv")

        def fn(tape, inputs, need):
            a = E.stage_input(tape, inputs[0], need[0])
            slot = {}
            return E.outconv_sigmoid(tape, a, self.conv.weight, self.conv.bias, slot), slot, [a]

        return E.run_net(self, fn, x)


class Segmentor(nn.Module):
    """Siamese U-Net producing the change-density map — Module.py:93-140.

    forward(x1, x2) -> (B, n_outchannels, H, W) in [0, 1].  The shared-weight encoder is run once per temporal
    image (BatchNorm statistics and running-stat updates are per call); each level's two
    branch outputs and the decoder's upsampled tensor are written straight into one NHWC concatenation buffer,
    so torch.cat (Module.py:116-132, 78) costs no copy."""

    def __init__(self, n_channels, n_outchannels=1, bilinear=False):
        super().__init__()
        self.n_channels = n_channels
        self.n_outchannels = n_outchannels
        self.bilinear = bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        factor = 2 if bilinear else 1
        self.down4 = Down(512, 1024 // factor)
        self.up1 = Up(2048, 1024 // factor, bilinear)
        self.up2 = Up(1024, 512 // factor, bilinear)
        self.up3 = Up(512, 256 // factor, bilinear)
        self.up4 = Up(256, 128, bilinear)
        self.outc = OutConv(128, n_outchannels)

    def forward(self, x1, x2):
        _check_input(x1, self.n_channels, "Segmentor")
        _check_input(x2, self.n_channels, "Segmentor")
        if x1.shape != x2.shape:
            raise ValueError("Segmentor: x1 and x2 must have the same shape")
        if min(x1.shape[2], x1.shape[3]) < 16:
            raise ValueError("Segmentor: tiles must be at least 16x16 (four 2x2 poolings)")

        def fn(tape, inputs, need):
            training = self.training
            N, _, H, W = inputs[0].shape
            packed = self.n_channels <= 16 and not (need[0] or need[1])
            if packed:
                a, b = E.PackedAct(inputs[0]), E.PackedAct(inputs[1])
            else:
                a = E.stage_input(tape, inputs[0], need[0])
                b = E.stage_input(tape, inputs[1], need[1])
            enc = [self.inc, self.down1.maxpool_conv[1], self.down2.maxpool_conv[1], self.down3.maxpool_conv[1],
                   self.down4.maxpool_conv[1]]
            ups = [self.up4, self.up3, self.up2, self.up1]       # ups[l] consumes the level-l skip
            ch = [m.double_conv[3].out_channels for m in enc]
            # channels arriving from below at level l (after bilinear upsampling or the transposed conv)
            below = [None] * 4
            for l in range(4):
                src = 2 * ch[4] if l == 3 else ups[l + 1].conv.double_conv[3].out_channels
                below[l] = src if self.bilinear else src // 2
            cats, A, B = [], [], []
            h, w = H, W
            for l in range(5):
                if l > 0:
                    h, w = h // 2, w // 2
                tot = 2 * ch[l] + (below[l] if l < 4 else 0)
                cat = tape.new_act(N, h, w, tot, tot, name=f"cat{l}")
                cats.append(cat)
                A.append(tape.track(cat.slice(0, ch[l])))
                B.append(tape.track(cat.slice(ch[l], ch[l])))
            for l in range(5):
                if l == 0:
                    _double_conv(tape, enc[0], a, training, out=A[0], x_needs_grad=need[0])
                    _double_conv(tape, enc[0], b, training, out=B[0], x_needs_grad=need[1])
                else:
                    _double_conv(tape, enc[l], E.maxpool2(tape, A[l - 1]), training, out=A[l])
                    _double_conv(tape, enc[l], E.maxpool2(tape, B[l - 1]), training, out=B[l])
            x = cats[4]
            for l in (3, 2, 1, 0):
                up_slot = tape.track(cats[l].slice(2 * ch[l], below[l]))
                x = ups[l]._run(tape, x, cats[l], up_slot, training)
            slot = {}
            out = E.outconv_sigmoid(tape, x, self.outc.conv.weight, self.outc.conv.bias, slot)
            return out, slot, ([None, None] if packed else [a, b])

        return E.run_net(self, fn, x1, x2)


class ResidualBlock(nn.Module):
    """Module.py:174-190."""

    def __init__(self, channels):
        super().__init__()
        _check_bn_width(channels, "ResidualBlock")
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.prelu = nn.PReLU()
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)

    def forward(self, x):
        _check_input(x, self.conv1.in_channels, "ResidualBlock")

        def fn(tape, inputs, need):
            a = E.stage_input(tape, inputs[0], need[0])
            o = _residual_block(tape, self, a, self.training)
            slot = {}
            return E.act_to_nchw(tape, o, slot), slot, [a]

        return E.run_net(self, fn, x)


class Generator(nn.Module):
    """SRGAN-style ResNet generator (no resampling, linear output) — Module.py:142-172."""

    def __init__(self, n_channels):
        super().__init__()
        self.n_channels = n_channels
        self.block1 = nn.Sequential(nn.Conv2d(n_channels, 64, kernel_size=9, padding=4), nn.PReLU())
        self.block2 = ResidualBlock(64)
        self.block3 = ResidualBlock(64)
        self.block4 = ResidualBlock(64)
        self.block5 = ResidualBlock(64)
        self.block6 = ResidualBlock(64)
        self.block7 = nn.Sequential(nn.Conv2d(64, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64))
        self.block8 = nn.Conv2d(64, n_channels, kernel_size=9, padding=4)

    def forward(self, x):
        _check_input(x, self.n_channels, "Generator")

        def fn(tape, inputs, need):
            training = self.training
            W = inputs[0].shape[3]
            if self.n_channels <= 16 and not need[0]:
                # few-band head on the channel-packed input: one K = pad64(9 C) tap per filter row (13 bands: 117 -> 128) instead of 9
                a = None
                xp = E.PackedAct(inputs[0], P=E.row_pack_pixels(self.n_channels, 9))
                z = E.conv_small_in(tape, xp, self.block1[0].weight, self.block1[0].bias, 4, stats=False)
            else:
                a = E.stage_input(tape, inputs[0], need[0])
                z = E.conv(tape, a, self.block1[0].weight, self.block1[0].bias, 1, 4, stats=False, x_needs_grad=need[0])
            b1 = E.bn_act(tape, z, None, training, E.ACT_PRELU, slope=self.block1[1].weight)
            h = b1
            for blk in (self.block2, self.block3, self.block4, self.block5, self.block6):
                h = _residual_block(tape, blk, h, training)
            z = E.conv(tape, h, self.block7[0].weight, self.block7[0].bias, 1, 1, stats=training)
            s = E.bn_act(tape, z, _bn(self.block7[1]), training, E.ACT_NONE, residual=b1)   # block1 + block7
            if self.n_channels <= 16 and W % 4 == 0:
                z = E.conv_small_out(tape, s, self.block8.weight, self.block8.bias, 4)     # 4 output pixels per MMA row
            else:
                z = E.conv(tape, s, self.block8.weight, self.block8.bias, 1, 4, stats=False)
            slot = {}
            return E.z_to_nchw(tape, z, slot), slot, [a]

        return E.run_net(self, fn, x)


class Discriminator_SRGAN_simple(nn.Module):
    """Siamese global discriminator — Module.py:192-223.  forward(x, y) -> (B,) sigmoid scores."""

    def __init__(self, n_channels=3):
        super().__init__()
        self.n_channels = n_channels
        self.net = nn.Sequential(
            nn.Conv2d(n_channels, 64, kernel_size=3, stride=2, padding=1),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1),
            nn.BatchNorm2d(128),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(128, 256, kernel_size=3, stride=2, padding=1),
            nn.BatchNorm2d(256),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(256, 512, kernel_size=3, stride=2, padding=1),
            nn.BatchNorm2d(512),
            nn.LeakyReLU(0.2, inplace=True),
        )
        self.classifier = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(512, 1024, kernel_size=1),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(1024, 1, kernel_size=1),
        )
        self.sigmoid = nn.Sigmoid()

    def _features(self, tape, a, training: bool, need_in: bool, groups: int = 1) -> E.Act:
        """The convolution stack of one branch (Module.py:196-210) — or, with `groups` = 2, of both branches as one batch
        (x in the first half, y in the second): the convolutions, their data and weight gradients run once over 2B images,
        BatchNorm keeps separate statistics per branch, in the reference's call order (x, then y)."""
        net = self.net
        if isinstance(a, E.Act):
            z = E.conv(tape, a, net[0].weight, net[0].bias, 2, 1, stats=False, x_needs_grad=need_in)
        else:       # NCHW data tensor(s): receptive-field-packed first layer (no input gradient needed)
            z = E.conv_im2col_s2(tape, a, net[0].weight, net[0].bias)
        h = E.bn_act(tape, z, None, training, E.ACT_LEAKY, slope_const=0.2)
        for ci, bi in ((2, 3), (5, 6), (8, 9)):
            z = E.conv(tape, h, net[ci].weight, net[ci].bias, 2, 1, stats=training and groups == 1)
            h = E.bn_act(tape, z, _bn(net[bi]), training, E.ACT_LEAKY, slope_const=0.2, groups=groups)
        return h

    def forward(self, x, y):
        _check_input(x, self.n_channels, "Discriminator_SRGAN_simple")
        _check_input(y, self.n_channels, "Discriminator_SRGAN_simple")
        if x.shape != y.shape:
            raise ValueError("Discriminator_SRGAN_simple: x and y must have the same shape")

        def fn(tape, inputs, need):
            training = self.training
            # an input that needs no gradient (data, or masks built from a detached map) skips the 64-channel zero-padded
            # staging: its first layer runs on receptive-field-packed rows (engine.conv_im2col_s2)
            packed = [E.im2col_enabled() and not need[i] and 9 * self.n_channels <= 256 for i in range(2)]
            c1, c3 = self.classifier[1], self.classifier[3]
            slot = {}
            B = inputs[0].shape[0]
            if E.batch_branches_enabled() and packed[0] == packed[1] and need[0] == need[1]:
                # both siamese branches as ONE batch of 2B images (x first): every convolution / gradient kernel launches once
                # instead of twice, on twice the pixels (the small late layers fill whole waves of CTAs)
                a2 = (inputs[0], inputs[1]) if packed[0] else E.stage_two_inputs(tape, inputs[0], inputs[1])
                h = self._features(tape, a2, training, need[0], groups=2)
                fx, fy = tape.track(h.batch_view(0, B)), tape.track(h.batch_view(B, B))
                tape.push(lambda tape: h.mark_ready())      # runs right after the head's backward has filled both halves
                out = E.disc_head(tape, fx, fy, c1.weight, c1.bias, c3.weight, c3.bias, slot)
                acts = [None, None] if packed[0] else [E.BatchHalf(a2, 0, B), E.BatchHalf(a2, B, B)]
                return out, slot, acts
            a = inputs[0] if packed[0] else E.stage_input(tape, inputs[0], need[0])
            b = inputs[1] if packed[1] else E.stage_input(tape, inputs[1], need[1])
            fx = self._features(tape, a, training, need[0])
            fy = self._features(tape, b, training, need[1])
            out = E.disc_head(tape, fx, fy, c1.weight, c1.bias, c3.weight, c3.bias, slot)
            return out, slot, [None if packed[0] else a, None if packed[1] else b]

        return E.run_net(self, fn, x, y)
