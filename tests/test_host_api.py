"""CPU: the host-side mirror of the reference interface — class names, constructor signatures, state_dict keys and
shapes (checked against the oracle's restatement of the reference layouts AND, where /root/reference is mounted,
against the real reference modules), error behaviour without a GPU."""
import inspect

import pytest
import torch

import fcdgan_b200 as fb
from fcdgan_b200._lib import FcdError
from oracle import fcd_oracle as O
from oracle import ref_import


def _shapes(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}


@pytest.mark.parametrize("C", [3, 4, 13])
def test_state_dict_layouts_match_reference_spec(C):
    assert _shapes(fb.Generator(C).state_dict()) == _shapes(O.make_state_dict(O.generator_spec(C), 0))
    assert _shapes(fb.Discriminator_SRGAN_simple(C).state_dict()) == _shapes(O.make_state_dict(O.discriminator_spec(C), 0))
    for bil in (True, False):
        assert _shapes(fb.Segmentor(C, 1, bil).state_dict()) == _shapes(O.make_state_dict(O.segmentor_spec(C, 1, bil), 0))
    assert len(fb.Generator(13).state_dict()) == 87 and len(fb.Discriminator_SRGAN_simple(13).state_dict()) == 27
    assert len(fb.Segmentor(13, 1, True).state_dict()) == 128
    assert sum(p.numel() for p in fb.Generator(13).parameters()) == 542483
    assert sum(p.numel() for p in fb.Segmentor(13, 1, True).parameters()) == 40833729
    assert sum(p.numel() for p in fb.Discriminator_SRGAN_simple(13).parameters()) == 2084865


@pytest.mark.skipif(not ref_import.available(), reason="reference not mounted (build container only)")
def test_state_dicts_interchange_with_the_real_reference():
    M, L, S = ref_import.load()
    for ours, ref in ((fb.Generator(4), M.Generator(4)), (fb.Segmentor(4, 1, True), M.Segmentor(4, 1, True)),
                      (fb.Segmentor(4, 1, False), M.Segmentor(4, 1, False)),
                      (fb.Discriminator_SRGAN_simple(4), M.Discriminator_SRGAN_simple(4))):
        assert _shapes(ours.state_dict()) == _shapes(ref.state_dict())
        ours.load_state_dict(ref.state_dict())          # reference .pkl -> ours
        ref.load_state_dict(ours.state_dict())          # ours -> reference
    # constructor / forward signatures
    for name in ("Generator", "Segmentor", "Discriminator_SRGAN_simple", "DoubleConv", "Down", "Up", "OutConv", "ResidualBlock"):
        a = inspect.signature(getattr(M, name).__init__)
        b = inspect.signature(getattr(fb, name).__init__)
        assert list(a.parameters) == list(b.parameters), name
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()], name
        fa = list(inspect.signature(getattr(M, name).forward).parameters)
        fbb = list(inspect.signature(getattr(fb, name).forward).parameters)
        assert fa == fbb, name
    for name in ("SSIM", "MS_SSIM"):
        a, b = inspect.signature(getattr(S, name).__init__), inspect.signature(getattr(fb, name).__init__)
        assert list(a.parameters) == list(b.parameters)
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()]
    assert list(inspect.signature(S.ms_ssim).parameters) == list(inspect.signature(fb.ms_ssim).parameters)
    assert list(inspect.signature(S.ssim).parameters) == list(inspect.signature(fb.ssim).parameters)
    assert list(inspect.signature(L.region_loss).parameters) == list(inspect.signature(fb.region_loss).parameters)
    for name in ("CNetLoss", "CGeneratorLoss"):
        a = list(inspect.signature(getattr(L, name).forward).parameters)
        b = list(inspect.signature(getattr(fb, name).forward).parameters)
        assert a == b, name
    # the Gaussian window and default MS-SSIM weights are the reference's
    import sys
    assert torch.equal(S._fspecial_gauss_1d(11, 1.5), sys.modules["fcdgan_b200.ssim"]._fspecial_gauss_1d(11, 1.5))


def test_gauss_window_matches_oracle():
    import sys
    ssim_mod = sys.modules["fcdgan_b200.ssim"]      # (the package attribute `ssim` is the function, like the reference's)
    assert torch.equal(ssim_mod._fspecial_gauss_1d(11, 1.5).flatten(), O.gauss_window(11, 1.5))
    m = fb.MS_SSIM(data_range=1.0, channel=13)
    assert tuple(m.win.shape) == (13, 1, 1, 11) and m.data_range == 1.0 and m.weights is None


def test_no_cpu_fallback_and_input_validation():
    g = fb.Generator(3)
    with pytest.raises(FcdError):
        g(torch.rand(1, 3, 16, 16))                     # CPU tensor: loud failure, no fallback
    with pytest.raises(ValueError):
        g(torch.rand(1, 4, 16, 16))                     # wrong band count
    s = fb.Segmentor(3, 1, True)
    with pytest.raises(ValueError):
        s(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 16))
    with pytest.raises(ValueError):
        s(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))
    with pytest.raises(ValueError):
        fb.ms_ssim(torch.rand(1, 3, 200, 200), torch.rand(1, 2, 200, 200))
    with pytest.raises(AssertionError):
        fb.ms_ssim(torch.rand(1, 3, 160, 160), torch.rand(1, 3, 160, 160))
    with pytest.raises(ValueError):
        fb.ssim(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32), win_size=8)
    with pytest.raises(NotImplementedError):
        fb.region_loss(torch.rand(1, 1, 8, 8), torch.rand(1, 1, 8, 8), torch.nn.BCELoss())
    with pytest.raises(ValueError):
        fb.set_precision("fp8")


def test_train_eval_and_optimizer_plumbing():
    g = fb.Generator(4)
    assert g.training and not g.eval().training and g.train().training
    opt = torch.optim.Adam(g.parameters(), lr=2e-4, betas=(0.9, 0.99))   # Demo_USSS.py:121
    assert len(opt.param_groups[0]["params"]) == len(list(g.parameters()))
    assert float(g.block1[1].weight) == 0.25                             # nn.PReLU() default slope, Module.py:147


def test_raster_tile_grid_matches_the_oracle_geometry():
    """fcdgan_b200.raster.TileGrid (host side of the raster kernels) against the numpy restatement of GDALDataset's geometry
    (data_utils.py:57-63, 151-176), which tests/test_raster_oracle.py pins to the unmodified reference."""
    import numpy as np

    from fcdgan_b200.raster import TileGrid
    from oracle import raster_oracle as RO

    for xs, ys, patch, pad in ((256, 256, (220, 220), (10, 10)), (463, 431, (220, 220), (10, 10)), (150, 131, (64, 48), (6, 4)),
                               (100, 100, (30, 30), (10, 10)), (64, 64, (64, 64), (0, 0)), (65, 33, (32, 32), (1, 3))):
        g, o = TileGrid(xs, ys, patch, pad), RO.tile_grid(xs, ys, patch, pad)
        assert g.patch_count() == RO.patch_count(o) and len(g) == g.patch_count()[0] * g.patch_count()[1]
        try:
            gather, crop = g.gather_geom(), g.crop_geom()
        except ValueError:
            # a tile whose start equals the padding reads from 0 but is written at offset `pad` (the reference's `> 0` rule);
            # when that overflows the patch the reference raises a numpy broadcast error, TileGrid a ValueError
            assert any(RO.slice_assign(o, *RO.item_xy(o, i))[2][0] + RO.slice_assign(o, *RO.item_xy(o, i))[1][2] > patch[0] or
                       RO.slice_assign(o, *RO.item_xy(o, i))[2][1] + RO.slice_assign(o, *RO.item_xy(o, i))[1][3] > patch[1]
                       for i in range(len(g)))
            continue
        for i in range(len(g)):
            sl, rd, wr = RO.slice_assign(o, *RO.item_xy(o, i))
            assert g.slice_assign(*g.item_xy(i)) == (sl, rd, wr)
            assert gather[i].tolist() == [rd[0], rd[1], rd[2], rd[3], wr[0], wr[1]]
            assert crop[i].tolist() == [pad[0], pad[1], sl[0], sl[1], sl[2], sl[3]]
        # the centre crops tile the raster exactly once
        cover = np.zeros((ys, xs), dtype=np.int32)
        for i in range(len(g)):
            _, _, x0, y0, w, h = crop[i]
            cover[y0:y0 + h, x0:x0 + w] += 1
        assert (cover == 1).all()
    with pytest.raises(ValueError):
        TileGrid(100, 100, (20, 20), (10, 10))


def test_usss_step_lean_equals_faithful_on_cpu_stand_ins():
    """Host logic of steps.usss_gen: the faithful body back-propagates Loss (retain_graph) and then NetLoss = Loss + w*l1, which
    hands G exactly 2 x dLoss and S only dNetLoss; the lean body does one sweep of NetLoss and doubles G's gradients.  Checked
    here with plain torch stand-ins for the networks and the criterion (the step bodies only use the call surface), including the
    order of the exchange points they yield."""
    import torch
    import torch.nn as nn

    from fcdgan_b200 import steps as S

    class Crit(nn.Module):
        loss_perception = type("P", (), {"enabled": False})()

        def forward(self, t, g, cmap):
            m = 1 - cmap
            gen = ((t - g).abs() * m).mean()
            return gen, cmap.abs().mean(), torch.zeros(()), ((t * m - g * m) ** 2).mean()

    class Seg(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(6, 1, 3, padding=1)

        def forward(self, x, y):
            return torch.sigmoid(self.c(torch.cat([x, y], 1)))

    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 3, 8, 8, generator=g)
    seen = {}
    for lean in (False, True):
        torch.manual_seed(1)
        netG, netS = nn.Conv2d(3, 3, 3, padding=1), Seg()
        order = []

        def hook(net, wait):
            order.append((net is netG, wait))
            seen[(lean, net is netG)] = [p.grad.clone() for p in net.parameters()]

        out = S.usss_step(netG, netS, x, y, Crit(), ssim_weight=0.3, l1_weight=0.65, on_grads=hook, lean=lean)
        assert order == [(True, False), (False, True)]            # G's bucket may stay in flight, S's is waited for
        assert set(out) >= {"generator_loss", "l1_loss", "ssim_loss", "Loss", "NetLoss", "cmap"}
    for is_g in (True, False):
        for a, b in zip(seen[(True, is_g)], seen[(False, is_g)]):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-8)
    with __import__("pytest").raises(ValueError, match="perception_weight"):
        S.usss_step(nn.Conv2d(3, 3, 1), Seg(), x, y, Crit(), perception_weight=0.4)


def test_any_optimizer_step_invalidates_the_packed_weight_cache():
    """Fused optimizers do not move Tensor._version, so the engine's cache stamp also carries an epoch that torch's global
    optimizer post-step hook bumps (engine._after_any_optimizer_step)."""
    import torch
    from fcdgan_b200 import engine as E
    m = torch.nn.Linear(4, 4)
    for kw in ({}, {"fused": True}):
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, **kw)
        m(torch.randn(2, 4)).sum().backward()
        e0 = E._weight_epoch
        opt.step()
        assert E._weight_epoch > e0


def test_tape_zero_pool_hands_out_disjoint_aligned_zeroed_slices():
    """engine.Tape.zeros64: accumulators carved from one zero-filled pool per block — never handed out twice (a replayed tape gets
    fresh zeros), 16-byte aligned, a new block when one is exhausted or a request is larger than a block."""
    import torch
    from fcdgan_b200 import engine as E
    tape = E.Tape(torch.device("cpu"), True)
    a = tape.zeros64(2, 72)
    b = tape.zeros64(3)                      # odd count: the next slice must still start on a 16-byte boundary
    c = tape.zeros64(2, 2, 64)
    for t in (a, b, c):
        assert t.dtype == torch.float64 and t.is_contiguous() and float(t.abs().sum()) == 0.0 and t.data_ptr() % 16 == 0
    a.fill_(1.0); b.fill_(2.0); c.fill_(3.0)
    assert float(a.sum()) == 144.0 and float(b.sum()) == 6.0 and float(c.sum()) == 3.0 * 256      # no overlap
    first = tape._arena
    big = tape.zeros64(E.Tape.ARENA + 2)     # larger than a block: its own block
    assert tape._arena is not first and float(big.abs().sum()) == 0.0 and big.numel() == E.Tape.ARENA + 2
    d = tape.zeros64(8)
    assert float(d.abs().sum()) == 0.0 and float(a.sum()) == 144.0                                 # earlier slices untouched
    E._cfg["zero_pool"] = False
    try:
        e = tape.zeros64(2, 8)
        assert e.shape == (2, 8) and float(e.abs().sum()) == 0.0
    finally:
        E._cfg["zero_pool"] = True


def test_act_batch_view_shares_data_and_gradient_with_its_parent():
    """engine.Act.batch_view: the images of one siamese branch inside the 2B batch the Discriminator's convolutions run on —
    data and gradient are slices of the parent's buffers (no copies), readiness follows the parent."""
    import torch
    from fcdgan_b200 import engine as E
    h = E.Act.empty(4, 3, 5, 64, torch.device("cpu"))
    fx, fy = h.batch_view(0, 2), h.batch_view(2, 2)
    assert fx.hi.data_ptr() == h.hi.data_ptr() and fy.hi.data_ptr() == h.hi[2:].data_ptr()
    assert (fx.N, fx.H, fx.W, fx.Cp, fx.ld) == (2, 3, 5, 64, 64)
    fy.grad.fill_(7.0)
    fx.grad.fill_(1.0)
    assert float(h.grad[:2].mean()) == 1.0 and float(h.grad[2:].mean()) == 7.0
    assert not fx.ready
    h.mark_ready()
    assert fx.ready and fy.ready
    h.reset_grad(); fx.reset_grad(); fy.reset_grad()
    assert fx.grad.data_ptr() == h.grad.data_ptr()          # a new replay: the views follow the parent's NEW gradient buffer
